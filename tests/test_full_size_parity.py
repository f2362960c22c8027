"""CUDA path (through the C ABI) against the CPU oracle at the sizes BASELINE.json names:
the meshes are the full-size ones (VMEC 100x40x40, EFIT symmetry-flux 100x40x40, WEST / SOLEDGE3X-EIRENE n2 = 60 with
4.24 M tetrahedra), >= 500 particles each, the visited (tetrahedron, face) sequence of the first 10^4 crossings of every
particle plus the complete final state, several successive orbit_timestep_gorilla calls.  Bar: bit-identical (strict
build); north_star's 1e-10 bound is asserted separately."""
from pathlib import Path

import numpy as np
import pytest

import workloads
from gorilla_b200 import Gorilla, build_mesh
from oracle_binding import OracleMesh

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
DATA = ROOT / "data" / "equilibria"
CAP = 10_000


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def rel_close(a, b, tol=1e-10):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return bool(np.all(np.abs(a - b) <= tol * np.maximum(1e-300, np.maximum(np.abs(a), np.abs(b)))))


def _with(settings, **kw):
    return type(settings)(**{**settings.__dict__, **kw})


def cross_check(mesh, settings, particles, t_step, n_calls, n_traced, gather=None):
    """n_calls successive calls on all particles; the first call records the trace of the first n_traced particles (oracle:
    one particle at a time), the other particles and calls go through the oracle's batched entry point (OpenMP)."""
    x, vpar, vperp = particles
    n = x.shape[0]
    om, g = OracleMesh(mesh, settings), Gorilla(mesh, settings)
    if gather is not None:
        g.set_gather(gather)
        assert g.get_gather() == gather
    xa, va, wa = x.copy(), vpar.copy(), vperp.copy()
    xb, vb, wb = x.copy(), vpar.copy(), vperp.copy()
    ia, ta, fa = workloads.fresh_state(n)
    ib, tb, fb = workloads.fresh_state(n)
    pushes = 0
    for call in range(n_calls):
        tro_a, npu_a = np.zeros(n), np.zeros(n, np.int64)
        tro_b, npu_b = np.zeros(n), np.zeros(n, np.int64)
        if call == 0:
            k = n_traced
            r = om.orbit_timestep_trace(xa[:k], va[:k], wa[:k], t_step, ia[:k], ta[:k], fa[:k], CAP)
            npu_a[:k], tro_a[:k] = r["n_pushes"], r["t_remain"]
            if k < n:
                om.orbit_timestep_batch(xa[k:], va[k:], wa[k:], t_step, ia[k:], ta[k:], fa[k:], t_remain_out=tro_a[k:],
                                        n_pushes=npu_a[k:])
            tt, tf = g.orbit_timestep_gorilla(xb[:k], vb[:k], wb[:k], t_step, ib[:k], tb[:k], fb[:k], t_remain_out=tro_b[:k],
                                              n_pushes=npu_b[:k], trace_cap=CAP)
            assert same(r["trace_tetr"], tt), "visited tetrahedron sequence differs"
            assert same(r["trace_face"], tf), "visited face sequence differs"
            assert int(npu_a[:k].max()) > 0
            if k < n:
                g.orbit_timestep_gorilla(xb[k:], vb[k:], wb[k:], t_step, ib[k:], tb[k:], fb[k:], t_remain_out=tro_b[k:],
                                         n_pushes=npu_b[k:])
        else:
            om.orbit_timestep_batch(xa, va, wa, t_step, ia, ta, fa, t_remain_out=tro_a, n_pushes=npu_a)
            g.orbit_timestep_gorilla(xb, vb, wb, t_step, ib, tb, fb, t_remain_out=tro_b, n_pushes=npu_b)
        assert rel_close(xa, xb) and rel_close(va, vb) and rel_close(wa, wb), "north_star: 1e-10 relative"
        assert same(xa, xb) and same(va, vb) and same(wa, wb), f"call {call}: state not bit-identical"
        assert same(ta, tb) and same(fa, fb) and same(ia, ib) and same(npu_a, npu_b) and same(tro_a, tro_b)
        pushes += int(npu_a.sum())
    e, p, mu = g.invariants(xb, vb, wb, tb)
    g.close()
    return dict(pushes=pushes, lost=int((ta == -1).sum()), x=xb, vpar=vb, vperp=wb, ind=tb, energy=e, p_phi=p, perpinv=mu)


@pytest.fixture(scope="module")
def vmec_full(product_lib):
    grid, settings = workloads.vmec_qi(DATA / "netcdf_file_for_test.nc")
    return build_mesh(grid, settings), settings


@pytest.fixture(scope="module")
def efit_flux_full(product_lib):
    grid, settings = workloads.efit_flux(DATA)
    return build_mesh(grid, settings), settings


@pytest.mark.parametrize("K", [2, 4])
def test_config3_vmec_alphas(vmec_full, cuda_device, K):
    """BASELINE config 3 / 5: QI stellarator, 100x40x40 = 960 000 tetrahedra, 3.5 MeV alphas at s = 0.5, steps of 1e-4 s."""
    mesh, settings = vmec_full
    assert mesh.ntetr == 960_000
    r = cross_check(mesh, _with(settings, poly_order=K), workloads.particles_vmec_alpha(600, 31), 1.0e-4, 3, 600)
    assert r["pushes"] > 600 * 3 * 1500


@pytest.mark.parametrize("pusher", ["poly2", "rk4"])
@pytest.mark.parametrize("gather", [1, 2])
def test_config3_vmec_alphas_staged_gathers(vmec_full, cuda_device, pusher, gather):
    """The same workload with the records staged in shared memory one push ahead -- per-lane bulk copies (1) and the
    warp-cooperative cp.async gather (2), which the library selects by itself only on meshes much larger than the L2 -- against
    the oracle: 2048 particles = 64 full warps, so the warp-level copy path, its per-lane fall-back (refills, pushes redone by
    the complete ladder, warps that empty towards the end of the queue) and the hand-over between them are all exercised."""
    mesh, settings = vmec_full
    st = _with(settings, ipusher=1) if pusher == "rk4" else _with(settings, ipusher=2, poly_order=2)
    r = cross_check(mesh, st, workloads.particles_vmec_alpha(2048, 37), 1.0e-4, 2, 256, gather=gather)
    assert r["pushes"] > 2048 * 2 * 1500


def test_config1_efit_flux_deuterons(efit_flux_full, cuda_device):
    """BASELINE config 1 exactly as SURVEY.md 8d states it: ASDEX Upgrade g-file, symmetry flux coordinates 100x40x40, 10^3 D+
    of 3 keV, s in U[0.2, 0.9], pitch in U[-1, 1], order 2, ten calls of 1e-3 s, PCG64(2024)."""
    mesh, settings = efit_flux_full
    assert mesh.ntetr == 960_000 and settings.poly_order == 2
    r = cross_check(mesh, settings, workloads.particles_flux(1000, 2024), 1.0e-3, 10, 500)
    assert r["pushes"] > 1000 * 10 * 2000


def test_config2_efit_flux_order4_conservation(efit_flux_full, cuda_device):
    """BASELINE config 2 (reduced particle count for the CPU side): order 4, energy / magnetic moment / p_phi conservation
    in the axisymmetric field, diagnosed on the device after every call."""
    mesh, settings = efit_flux_full
    st = _with(settings, poly_order=4)
    x, vpar, vperp = workloads.particles_flux(500, 7)
    g = Gorilla(mesh, st)
    b, i, f = workloads.fresh_state(500)
    g.orbit_timestep_gorilla(x, vpar, vperp, 0.0, b, i, f)
    e0, p0, m0 = g.invariants(x, vpar, vperp, i)
    g.close()
    r = cross_check(mesh, st, workloads.particles_flux(500, 7), 1.0e-4, 3, 500)
    alive = r["ind"] > 0
    assert alive.sum() > 450
    assert np.abs(r["energy"][alive] / e0[alive] - 1).max() < 1e-9
    assert np.abs(r["perpinv"][alive] / m0[alive] - 1).max() < 1e-12
    assert np.abs(r["p_phi"][alive] / p0[alive] - 1).max() < 1e-7     # axisymmetry: p_phi conserved to the mesh's accuracy


def test_config4_west_soledge3x_rk4_strong_e(cuda_device, product_lib):
    """BASELINE config 4: WEST equilibrium on the SOLEDGE3X-EIRENE mesh at n2 = 60 (4 242 060 tetrahedra), strong electric
    field eps_Phi = -1.5e-5, RK4 pusher, W74+ started uniformly over the poloidal mesh (open field line region included)."""
    grid, settings = workloads.west_soledge3x(DATA, n2=60)
    mesh = build_mesh(grid, settings)
    assert mesh.ntetr == 23567 * 3 * 60 and settings.ipusher == 1 and settings.boole_strong_electric_field
    r = cross_check(mesh, settings, workloads.particles_on_triangles(DATA, 600, 5), 1.0e-4, 3, 600)
    assert r["lost"] > 50 and r["pushes"] > 600 * 100
    # the polynomial pusher on the same mesh and field
    r2 = cross_check(mesh, _with(settings, ipusher=2, poly_order=2), workloads.particles_on_triangles(DATA, 500, 6), 1.0e-4, 2, 500)
    assert r2["pushes"] > 500 * 100


def test_demo_particle_of_test_gorilla_main(cuda_device, product_lib):
    """The one orbit the reference's own driver integrates (SRC/test_gorilla_main.f90:98-150, i_option = 2) with the
    blueprint inputs INPUT/gorilla.inp / INPUT/tetra_grid.inp: VMEC grid 100x40x40, ispecies = 1, order 2, two calls of
    t_step = 0.1 s from x = (0.5, 0.1, 0.63).  The first dump of the gfortran build has this as its obvious target
    (tests/reference_dump.py); until then it is pinned oracle <-> CUDA and by the invariants."""
    grid, settings = workloads.vmec_qi(DATA / "netcdf_file_for_test.nc")
    settings = _with(settings, ispecies=1, poly_order=2)
    mesh = build_mesh(grid, settings)
    x = np.array([[0.5, 0.1, 0.63]])
    vpar, vperp = np.array([37525024.533239894]), np.array([38283182.426206760])
    r = cross_check(mesh, settings, (x, vpar, vperp), 0.1, 2, 1)
    assert r["pushes"] > 200_000 and r["ind"][0] > 0
    v2 = vpar[0] ** 2 + vperp[0] ** 2
    assert abs((r["vpar"][0] ** 2 + r["vperp"][0] ** 2) / v2 - 1) < 1e-6     # no potential: |v| conserved (order 2 accuracy)
