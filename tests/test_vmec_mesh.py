"""grid_kind = 3 (VMEC, symmetry-flux coordinates): host mesh builder checks and device-algorithm parity
(host compile) in flux coordinates, where the handover applies the theta and phi periodic shifts."""
from pathlib import Path

import numpy as np
import pytest

import workloads
from host_mirror_binding import HostMirror
from oracle_binding import OracleMesh

NC = Path(__file__).resolve().parent.parent / "data" / "equilibria" / "netcdf_file_for_test.nc"


@pytest.fixture(scope="module")
def vmec_mesh(product_lib):
    from gorilla_b200 import build_mesh
    grid, settings = workloads.vmec_qi(NC, 16, 10, 12)
    return build_mesh(grid, settings), grid, settings


def test_netcdf_reader_matches_scipy(vmec_mesh):
    """Toroidal flux and field strength derived from our NetCDF-3 reader agree with scipy's reader."""
    from scipy.io import netcdf_file
    mesh, grid, _ = vmec_mesh
    f = netcdf_file(str(NC), "r", mmap=False)
    phi_edge = float(f.variables["phi"].data[-1])
    nfp = int(f.variables["nfp"].data)
    assert mesh.scalars["n_field_periods"] == nfp == 5
    torflux = phi_edge / (2 * 3.14159265358979) * 1e8
    # A_theta = torflux * s at the first vertex of every tetrahedron (Atheta1, offset 25)
    np.testing.assert_allclose(mesh.tetra_physics[:, 25], torflux * mesh.tetra_physics[:, 0], rtol=1e-13)
    b0 = float(f.variables["b0"].data) if "b0" in f.variables else 4.9
    assert 0.7 * b0 * 1e4 < np.median(mesh.tetra_physics[:, 24]) < 1.4 * b0 * 1e4


def test_topology_consistency(vmec_mesh):
    """The reference's own mesh consistency check (circular_mesh.f90:492-510) + periodic flag bookkeeping."""
    mesh, grid, _ = vmec_mesh
    tg, tp = mesh.tetra_grid, mesh.tetra_physics
    n1, n2, n3 = grid.n1, grid.n2, grid.n3
    assert mesh.ntetr == 6 * n1 * n2 * n3
    nb, nf, pp, pt = tg[:, 4:8], tg[:, 8:12], tg[:, 12:16], tg[:, 16:20]
    has = nb > 0
    assert (~has).sum() == 2 * (2 * n3) * n2           # inner (s_min) and outer (s = 1) boundary triangles
    ti, fi = np.nonzero(has)
    assert (nb[nb[ti, fi] - 1, nf[ti, fi] - 1] == ti + 1).all()
    assert (pp[ti, fi] == -pp[nb[ti, fi] - 1, nf[ti, fi] - 1]).all()
    assert (pt[ti, fi] == -pt[nb[ti, fi] - 1, nf[ti, fi] - 1]).all()
    tps = mesh.ntetr // n2
    assert (pp == -1).sum() == tps // 3 and (pp == 1).sum() == tps // 3
    assert (pt == -1).sum() == 2 * n1 * n2 and (pt == 1).sum() == 2 * n1 * n2
    # toroidal neighbours are +- one slice
    t0 = np.arange(0, tps, 3)
    assert (nb[t0 + tps, 3] == t0 + 2 + 1).all()
    # no overlapping tetrahedra, positive volumes, sign of sqrt(g) negative for this equilibrium
    assert not ((nf == -1) & has).any()
    assert (tp[:, 3] > 0).all() and mesh.scalars["sign_sqg"] == -1
    assert np.isfinite(tp).all()
    # cells tile the annulus s in [s_min, 1] x theta x one field period
    np.testing.assert_allclose(tp[:, 3].sum() / 6.0, 0.9 * 2 * np.pi * 2 * np.pi / 5, rtol=1e-9)


def test_field_is_divergence_free_and_consistent(vmec_mesh):
    """curl A of the linearised field reproduces sqrt(g) B^k: B^phi sqrt(g) = dA_theta/ds = torflux exactly,
    B^theta sqrt(g) = -dA_phi/ds = torflux*iota; |B|^2 = B^k B_k."""
    mesh, _, _ = vmec_mesh
    tp = mesh.tetra_physics
    torflux = tp[0, 25] / tp[0, 0]
    curlA = tp[:, 21:24]
    np.testing.assert_allclose(curlA[:, 2], torflux, rtol=1e-9)       # dA_theta/ds
    iota = curlA[:, 1] / curlA[:, 2]                                  # -dA_phi/ds / torflux
    assert 0.85 < np.median(iota) < 1.05 and (iota > 0.8).all() and (iota < 1.1).all()
    np.testing.assert_allclose(curlA[:, 0], 0.0, atol=1e-6 * abs(torflux))
    # h is a unit vector: h_k B^k / |B| = 1  with B^k = curlA^k / sqrt(g)
    sqg = tp[:, 39]
    hB = (tp[:, 28] * curlA[:, 1] + tp[:, 29] * curlA[:, 2]) / sqg / tp[:, 24]
    np.testing.assert_allclose(hB, 1.0, rtol=2e-2)   # linearisation error of a coarse 16x10x12 grid


@pytest.mark.parametrize("K", [2, 3, 4])
def test_flux_coordinate_parity_host_mirror_vs_oracle(vmec_mesh, K):
    mesh, _, settings = vmec_mesh
    settings = type(settings)(**{**settings.__dict__, "poly_order": K})
    om, hm = OracleMesh(mesh, settings), HostMirror(mesh, settings)
    n = 150
    xa, va, wa = workloads.particles_vmec_alpha(n, 3)
    xa[::7, 1] += 2 * np.pi       # exercises boole_periodic_relocation = .true. (modulo)
    xa[1::7, 2] -= 2 * np.pi / 5
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    sa, sb = workloads.fresh_state(n), workloads.fresh_state(n)
    for _ in range(2):
        ra = om.orbit_timestep_trace(xa, va, wa, 3e-5, *sa, 512)
        rb = hm.orbit_timestep(xb, vb, wb, 3e-5, *sb, 512)
        for a, b in ((ra["trace_tetr"], rb["trace_tetr"]), (ra["trace_face"], rb["trace_face"]), (xa, xb), (va, vb),
                     (wa, wb), (sa[1], sb[1]), (sa[2], sb[2]), (ra["t_remain"], rb["t_remain"])):
            assert np.array_equal(a, b)
    assert ra["n_pushes"].sum() > 10000, ra["n_pushes"].sum()
    # periodic coordinates stay inside the fundamental domain (up to the cell the particle is in)
    assert xa[:, 1].min() > -0.6 and xa[:, 1].max() < 2 * np.pi + 0.6
    assert xa[:, 2].min() > -0.2 and xa[:, 2].max() < 2 * np.pi / 5 + 0.2


def test_energy_and_moment_conservation_in_the_stellarator(vmec_mesh):
    mesh, _, settings = vmec_mesh
    settings = type(settings)(**{**settings.__dict__, "poly_order": 4})
    om = OracleMesh(mesh, settings)
    n = 100
    x, vpar, vperp = workloads.particles_vmec_alpha(n, 9)
    st = workloads.fresh_state(n)
    om.orbit_timestep_batch(x, vpar, vperp, 0.0, *st)
    e0, _, mu0 = om.invariants(x, vpar, vperp, st[1])
    om.orbit_timestep_batch(x, vpar, vperp, 2e-5, *st, nthreads=4)
    e1, _, mu1 = om.invariants(x, vpar, vperp, st[1])
    ok = st[1] > 0
    assert ok.sum() > 90
    assert np.abs(e1 / e0 - 1)[ok].max() < 1e-6 and np.abs(mu1 / mu0 - 1)[ok].max() < 1e-13
