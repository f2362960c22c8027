"""CUDA kernels (through the C ABI, host buffers) against the CPU oracle on identical seeded inputs.

Bar (BASELINE.json north_star): visited-tetra sequence bit-exact, positions/velocities within 1e-10 relative.
The strict (--fmad=false) build does better: every output is IDENTICAL to the oracle's, which is what is
asserted; the 1e-10 bound is asserted separately so that a relaxed build would still be checked against it."""
import numpy as np
import pytest

import workloads
from oracle_binding import OracleMesh

pytestmark = pytest.mark.gpu


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def rel_close(a, b, tol=1e-10):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return bool(np.all(np.abs(a - b) <= tol * np.maximum(1e-300, np.maximum(np.abs(a), np.abs(b)))))


def _gorilla(mesh, settings):
    from gorilla_b200 import Gorilla
    return Gorilla(mesh, settings)


def run_pair(mesh, settings, n, seed, t_step, cap, nsteps=1, force_full=False, **pk):
    om = OracleMesh(mesh, settings)
    g = _gorilla(mesh, settings)
    g._debug_force_full(force_full)
    xa, va, wa = workloads.particles_cyl(n, seed, **pk)
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    ia, ta, fa = workloads.fresh_state(n)
    ib, tb, fb = workloads.fresh_state(n)
    out = None
    for _ in range(nsteps):
        ra = om.orbit_timestep_trace(xa, va, wa, t_step, ia, ta, fa, cap)
        tro, npu = np.zeros(n), np.zeros(n, np.int64)
        tt, tf = g.orbit_timestep_gorilla(xb, vb, wb, t_step, ib, tb, fb, t_remain_out=tro, n_pushes=npu, trace_cap=cap)
        c = g.counters()
        assert same(ra["trace_tetr"], tt), "visited tetra sequence differs"
        assert same(ra["trace_face"], tf)
        assert same(ra["n_pushes"], npu) and c.n_pushes == int(ra["n_pushes"].sum())
        assert rel_close(xa, xb) and rel_close(va, vb) and rel_close(wa, wb)
        assert same(xa, xb) and same(va, vb) and same(wa, wb), "strict build must be bit-identical"
        assert same(ta, tb) and same(fa, fb) and same(ia, ib)
        assert same(ra["t_remain"], tro)
        assert tuple(int(v) for v in ra["fallback"]) == c.n_fallback
        assert c.n_lost == int((ta == -1).sum())
        out = (ra, ta, c)
    g.close()
    return out


@pytest.mark.parametrize("K", [1, 2, 3, 4])
def test_tetra_sequence_and_state_bit_exact(small_mesh, cuda_device, K):
    mesh, _, settings = small_mesh
    settings = type(settings)(**{**settings.__dict__, "poly_order": K})
    ra, ind, c = run_pair(mesh, settings, 1000, 3, 2e-5, 128)
    assert c.n_pushes > 40000 and c.kernel_ms > 0.0


@pytest.mark.parametrize("K", [2, 4])
def test_complete_ladder_path_bit_exact(small_mesh, cuda_device, K):
    mesh, _, settings = small_mesh
    settings = type(settings)(**{**settings.__dict__, "poly_order": K})
    run_pair(mesh, settings, 300, 31, 1e-5, 64, force_full=True)


@pytest.mark.parametrize("K", [2, 3, 4])
def test_with_electrostatic_potential(small_mesh_phi, cuda_device, K):
    mesh, _, settings = small_mesh_phi
    settings = type(settings)(**{**settings.__dict__, "poly_order": K})
    run_pair(mesh, settings, 500, 5, 2e-5, 128)


@pytest.mark.parametrize("K", [2, 4])
def test_backward_in_time_and_repeated_calls(small_mesh, cuda_device, K):
    mesh, _, settings = small_mesh
    settings = type(settings)(**{**settings.__dict__, "poly_order": K})
    run_pair(mesh, settings, 400, 8, -1e-5, 64, nsteps=3)


@pytest.mark.parametrize("K", [2, 4])
def test_losses_through_the_domain_boundary(small_mesh, cuda_device, K):
    mesh, _, settings = small_mesh
    settings = type(settings)(**{**settings.__dict__, "poly_order": K})
    ra, ind, c = run_pair(mesh, settings, 600, 13, 2e-4, 64, rmin_frac=0.9, rmax_frac=0.99, energy_ev=3.0e4)
    assert c.n_lost > 30 and c.n_finished > 30 and c.n_lost + c.n_finished == 600


def test_find_tetra_matches_oracle_including_face_starts(small_mesh, cuda_device):
    mesh, grid, settings = small_mesh
    om, g = OracleMesh(mesh, settings), _gorilla(mesh, settings)
    n = 3000
    xa, va, wa = workloads.particles_cyl(n, 17, rmin_frac=0.05, rmax_frac=0.95)
    hr, hz, hphi = 100.0 / grid.n1, 100.0 / grid.n3, 2 * np.pi / grid.n2
    xa[0::5, 0] = 120.0 + hr * np.round((xa[0::5, 0] - 120.0) / hr)
    xa[1::5, 2] = -50.0 + hz * np.round((xa[1::5, 2] + 50.0) / hz)
    xa[2::5, 1] = hphi * np.floor(xa[2::5, 1] / hphi)
    xb = xa.copy()
    ta, fa = om.find_tetra(xa, va, wa)
    tb, fb = g.find_tetra(xb, va, wa)
    assert same(ta, tb) and same(fa, fb) and same(xa, xb)
    assert (ta > 0).all() and (fa > 0).sum() > 100
    g.close()


def test_invariants_match_oracle(small_mesh_phi, cuda_device):
    mesh, _, settings = small_mesh_phi
    om, g = OracleMesh(mesh, settings), _gorilla(mesh, settings)
    n = 500
    x, vpar, vperp = workloads.particles_cyl(n, 23)
    ind, _ = g.find_tetra(x, vpar, vperp)
    ind[::50] = -1
    e, p, mu = g.invariants(x, vpar, vperp, ind)
    eo, po, muo = om.invariants(x, vpar, vperp, ind)
    assert same(e, eo) and same(p, po) and same(mu, muo)
    assert np.isnan(e[::50]).all()
    g.close()


def test_edge_cases_empty_scalar_zero_step(small_mesh, cuda_device):
    mesh, _, settings = small_mesh
    g = _gorilla(mesh, settings)
    # n = 0
    z = np.zeros((0, 3))
    g.orbit_timestep_gorilla(z, np.zeros(0), np.zeros(0), 1e-5, *workloads.fresh_state(0))
    # n = 1 is the reference's scalar call
    om = OracleMesh(mesh, settings)
    xa, va, wa = workloads.particles_cyl(1, 77)
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    sa, sb = workloads.fresh_state(1), workloads.fresh_state(1)
    om.orbit_timestep_batch(xa, va, wa, 3e-5, *sa)
    g.orbit_timestep_gorilla(xb, vb, wb, 3e-5, *sb)
    assert same(xa, xb) and same(va, vb) and same(wa, wb) and same(sa[1], sb[1]) and sb[0][0] == 1
    # t_step = 0 only localises; particles outside the mesh are reported in-band
    x, vpar, vperp = workloads.particles_cyl(64, 5)
    st = workloads.fresh_state(64)
    x0 = x.copy()
    g.orbit_timestep_gorilla(x, vpar, vperp, 0.0, *st)
    assert (st[0] == 1).all() and (st[1] > 0).all() and same(x, x0) and g.counters().n_pushes == 0
    x[:, 0] = 500.0  # outside [Rmin, Rmax]
    st = workloads.fresh_state(64)
    g.orbit_timestep_gorilla(x, vpar, vperp, 1e-5, *st)
    assert (st[0] == 0).all() and (st[1] == -1).all() and (st[2] == -1).all()
    # phi outside [0, 2pi] with boole_periodic_relocation = .false. is a domain error (reference: stop)
    from gorilla_b200 import GorillaError
    x, vpar, vperp = workloads.particles_cyl(8, 5)
    x[3, 1] = 7.0
    with pytest.raises(GorillaError) as ei:
        g.orbit_timestep_gorilla(x, vpar, vperp, 1e-5, *workloads.fresh_state(8))
    assert ei.value.code == 4
    g.close()


def test_device_resident_api_and_sorting(small_mesh, cuda_device):
    """orbit_timestep_gorilla_dev on torch tensors == host-buffer API; result independent of particle order
    (lane refill + sort by tetra index only change the schedule, never a particle's orbit)."""
    import torch
    mesh, _, settings = small_mesh
    settings = type(settings)(**{**settings.__dict__, "poly_order": 4})
    g = _gorilla(mesh, settings)
    n = 5000
    x, vpar, vperp = workloads.particles_cyl(n, 41)
    st = workloads.fresh_state(n)
    xh, vh, wh = x.copy(), vpar.copy(), vperp.copy()
    sth = tuple(a.copy() for a in st)
    g.orbit_timestep_gorilla(xh, vh, wh, 2e-5, *sth)
    dev = cuda_device
    t = lambda a: torch.from_numpy(a.copy()).to(dev)  # noqa: E731
    xd, vd, wd, bi, it, ifc = t(x), t(vpar), t(vperp), t(st[0]), t(st[1]), t(st[2])
    g.orbit_timestep_gorilla_dev(xd, vd, wd, 0.0, bi, it, ifc)  # localise
    perm = torch.empty(n, dtype=torch.int64, device=dev)
    g.sort_permutation_dev(it, perm)
    torch.cuda.synchronize()
    assert torch.equal(torch.sort(perm).values, torch.arange(n, device=dev))
    keys = it[perm]
    assert bool((keys[1:] >= keys[:-1]).all())
    xd, vd, wd, bi, it, ifc = (a[perm].contiguous() for a in (xd, vd, wd, bi, it, ifc))
    npu = torch.zeros(n, dtype=torch.int64, device=dev)
    g.orbit_timestep_gorilla_dev(xd, vd, wd, 2e-5, bi, it, ifc, n_pushes=npu)
    torch.cuda.synchronize()
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(n, device=dev)
    assert same(xd[inv].cpu().numpy(), xh) and same(vd[inv].cpu().numpy(), vh) and same(it[inv].cpu().numpy(), sth[1])
    assert int(npu.sum()) == g.counters().n_pushes
    g.close()


def test_large_batch_properties(small_mesh, cuda_device):
    """Size-independent properties at a batch far beyond what the oracle is run on: magnetic moment carried
    exactly, energy drift at round-off (order 4), forward+backward step returns to the start."""
    mesh, _, settings = small_mesh
    settings = type(settings)(**{**settings.__dict__, "poly_order": 4})
    g = _gorilla(mesh, settings)
    n = 200_000
    x, vpar, vperp = workloads.particles_cyl(n, 99)
    st = workloads.fresh_state(n)
    g.orbit_timestep_gorilla(x, vpar, vperp, 0.0, *st)
    x0, v0 = x.copy(), vpar.copy()
    e0, p0, mu0 = g.invariants(x, vpar, vperp, st[1])
    g.orbit_timestep_gorilla(x, vpar, vperp, 2e-5, *st)
    c = g.counters()
    assert c.n_lost == 0 and c.n_finished == n and c.n_pushes > 5 * n
    e1, p1, mu1 = g.invariants(x, vpar, vperp, st[1])
    assert np.abs(mu1 / mu0 - 1).max() < 1e-13
    assert np.abs(e1 / e0 - 1).max() < 1e-11
    assert np.abs(p1 / p0 - 1).max() < 1e-8
    g.orbit_timestep_gorilla(x, vpar, vperp, -2e-5, *st)
    # order 4 is time-reversible up to truncation error: nearly every particle retraces its orbit
    back = np.abs(x - x0).max(axis=1)
    assert np.median(back) < 1e-9 and (back < 1e-6).mean() > 0.99
    assert (np.abs(vpar - v0) < 1e-6 * np.abs(v0).max()).mean() > 0.99
    g.close()


@pytest.mark.parametrize("K", [2, 4])
def test_vmec_flux_coordinates_bit_exact(cuda_device, product_lib, K):
    """BASELINE config 3 geometry (QI stellarator, symmetry-flux coordinates, 3.5 MeV alphas) on a reduced grid:
    theta/phi periodic handover, negative sqrt(g) (sign_sqg = -1), periodic relocation of start points."""
    from pathlib import Path
    from gorilla_b200 import build_mesh
    nc = Path(__file__).resolve().parent.parent / "data" / "equilibria" / "netcdf_file_for_test.nc"
    grid, settings = workloads.vmec_qi(nc, 16, 10, 12, poly_order=K)
    mesh = build_mesh(grid, settings)
    om, g = OracleMesh(mesh, settings), _gorilla(mesh, settings)
    n = 600
    xa, va, wa = workloads.particles_vmec_alpha(n, 3)
    xa[::7, 1] += 2 * np.pi
    xa[1::7, 2] -= 2 * np.pi / 5
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    sa, sb = workloads.fresh_state(n), workloads.fresh_state(n)
    for _ in range(2):
        ra = om.orbit_timestep_trace(xa, va, wa, 3e-5, *sa, 256)
        tt, tf = g.orbit_timestep_gorilla(xb, vb, wb, 3e-5, *sb, trace_cap=256)
        assert same(ra["trace_tetr"], tt) and same(ra["trace_face"], tf)
        assert same(xa, xb) and same(va, vb) and same(wa, wb) and same(sa[1], sb[1]) and same(sa[2], sb[2])
    assert g.counters().n_pushes > 10000
    g.close()


@pytest.mark.parametrize("K", [3, 4])
def test_rebinned_solver_kernel_bit_exact(small_mesh, small_mesh_phi, cuda_device, K):
    """orbit_kernel_g with the root solves of a group re-binned by solver mode between iterations (solver state in shared
    memory, counting sort, dense hand-out): the arithmetic of an iteration is unchanged, so every result is identical to the
    oracle's -- incl. lanes without a solve, pushes that fall back to the complete ladder, losses and queue refills."""
    for mesh, _, settings in (small_mesh, small_mesh_phi):
        st = type(settings)(**{**settings.__dict__, "poly_order": K})
        om, g = OracleMesh(mesh, st), _gorilla(mesh, st)
        g._debug_use_group(2)
        n = 3000     # more particles than one CTA holds, not a multiple of the group size
        xa, va, wa = workloads.particles_cyl(n, 41, rmax_frac=0.97, energy_ev=2.0e4)
        xb, vb, wb = xa.copy(), va.copy(), wa.copy()
        sa, sb = workloads.fresh_state(n), workloads.fresh_state(n)
        for t_step in (1.5e-5, -1.0e-5):
            ra = om.orbit_timestep_trace(xa, va, wa, t_step, *sa, 64)
            tro, npu = np.zeros(n), np.zeros(n, np.int64)
            tt, tf = g.orbit_timestep_gorilla(xb, vb, wb, t_step, *sb, t_remain_out=tro, n_pushes=npu, trace_cap=64)
            c = g.counters()
            assert same(ra["trace_tetr"], tt) and same(ra["trace_face"], tf)
            assert same(xa, xb) and same(va, vb) and same(wa, wb) and same(ra["t_remain"], tro)
            assert same(sa[1], sb[1]) and same(sa[2], sb[2]) and same(ra["n_pushes"], npu)
            assert tuple(int(v) for v in ra["fallback"]) == c.n_fallback
        assert (sa[1] == -1).sum() > 0
        g.close()
