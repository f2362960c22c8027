"""Strong-electric-field mode (boole_strong_electric_field): the extra ExB-drift terms of
the ODE (pusher_tetra_poly.f90:175,1519-1526,2728-2732,2799; pusher_tetra_rk.f90:126-134,152,672), the record
fields behind them (tetra_physics_mod.f90:531-536,625-630,708-749,820-855,884-893) and the vertex drift
(strong_electric_field_mod.f90).  Cylindrical grids only, as in the reference (gorilla_settings_mod.f90:139)."""
import numpy as np
import pytest

import workloads
from gorilla_b200 import api, build_mesh
from host_mirror_binding import HostMirror
from oracle_binding import OracleMesh

CLIGHT = 2.9979e10
EPS_PHI = -1.5e-5   # MATLAB/example_8.m (WEST case of SURVEY.md config 4)


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


@pytest.fixture(scope="module")
def strong_mesh(product_lib):
    grid, settings = workloads.analytic_tokamak(16, 16, 16)
    settings.eps_Phi = EPS_PHI
    settings.boole_strong_electric_field = True
    return build_mesh(grid, settings), grid, settings


@pytest.fixture(scope="module")
def weak_mesh(product_lib):
    grid, settings = workloads.analytic_tokamak(16, 16, 16)
    settings.eps_Phi = EPS_PHI
    return build_mesh(grid, settings), grid, settings


def _with(settings, **kw):
    return type(settings)(**{**settings.__dict__, **kw})


def test_record_fields_follow_the_reference_formulas(strong_mesh, weak_mesh):
    mesh, grid, _ = strong_mesh
    tp, tw = mesh.tetra_physics, weak_mesh[0].tetra_physics
    # everything that is not a strong-field quantity is unchanged; Er_mod is replaced by v_E_mod_average
    keep = np.ones(142, bool)
    for lo, hi in ((33, 39), (43, 47), (49, 50), (56, 59), (92, 107), (125, 134), (138, 142)):
        keep[lo:hi] = False
    assert same(tp[:, keep], tw[:, keep])
    assert np.all(tp[:, 37] == 0.0) and np.all(tw[:, 37] > 0.0) and np.all(tp[:, 38] > 0.0)
    # v_E = c E x h / B with E = -grad(psi)*eps_Phi: purely poloidal E -> |v_E| ~ c |E| / B; sanity of magnitude and
    # of the relation v2Emod_1 = v_E . v_E (cylindrical metric) at the first vertex
    R1 = tp[:, 31]
    v1, v2, v3, v2mod = tp[:, 33], tp[:, 34], tp[:, 35], tp[:, 36]
    assert np.allclose(v2mod, v1 * v1 + v2 * v2 / R1 ** 2 + v3 * v3, rtol=1e-14)
    assert 1e5 < np.sqrt(v2mod).mean() < 1e8
    # gamma matrix trace, curl of v_E and the two contractions, re-derived from the stored gradients
    gvE = tp[:, 92:101].reshape(-1, 3, 3)         # gvE[k] = grad of covariant component k
    curl = np.stack([gvE[:, 2, 1] - gvE[:, 1, 2], gvE[:, 0, 2] - gvE[:, 2, 0], gvE[:, 1, 0] - gvE[:, 0, 1]], 1)
    assert same(curl, tp[:, 101:104])
    gam = tp[:, 125:134].reshape(-1, 3, 3)        # column-major: [j][i] = gammat(i+1, j+1)
    assert same(tp[:, 49], gam[:, 0, 0] + gam[:, 1, 1] + gam[:, 2, 2])
    gv2 = tp[:, 104:107]
    assert same(tp[:, 46], gv2[:, 0] * curl[:, 0] + gv2[:, 1] * curl[:, 1] + gv2[:, 2] * curl[:, 2])
    anorm = tp[:, 9:21].reshape(-1, 4, 3)
    acoef_se = ((0.0 + curl[:, None, 0] * anorm[:, :, 0]) + curl[:, None, 1] * anorm[:, :, 1]) + curl[:, None, 2] * anorm[:, :, 2]
    assert same(tp[:, 138:142], acoef_se)


def test_settings_rules(product_lib, strong_mesh):
    mesh, _, settings = strong_mesh
    for bad in (_with(settings, i_precomp=1), _with(settings, ipusher=1, boole_newton_precalc=True)):
        with pytest.raises(api.GorillaError):
            api.Gorilla(mesh, bad)


def run_pair(mesh, settings, n, seed, t_step, cap, force_full=False, nsteps=1):
    om, hm = OracleMesh(mesh, settings), HostMirror(mesh, settings)
    xa, va, wa = workloads.particles_cyl(n, seed)
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    ia, ta, fa = workloads.fresh_state(n)
    ib, tb, fb = workloads.fresh_state(n)
    for _ in range(nsteps):
        ra = om.orbit_timestep_trace(xa, va, wa, t_step, ia, ta, fa, cap)
        rb = hm.orbit_timestep(xb, vb, wb, t_step, ib, tb, fb, cap, force_full=force_full)
        assert same(ra["trace_tetr"], rb["trace_tetr"]) and same(ra["trace_face"], rb["trace_face"])
        assert same(ra["n_pushes"], rb["n_pushes"])
        assert same(xa, xb) and same(va, vb) and same(wa, wb)
        assert same(ta, tb) and same(fa, fb) and same(ia, ib)
        assert same(ra["t_remain"], rb["t_remain"]) and same(ra["fallback"], rb["fallback"])
    return ra, xa, va, ta


@pytest.mark.parametrize("K", [1, 2, 3, 4])
@pytest.mark.parametrize("force_full", [False, True])
def test_host_mirror_parity_polynomial(strong_mesh, K, force_full):
    mesh, _, settings = strong_mesh
    ra, *_ = run_pair(mesh, _with(settings, poly_order=K), 150, 3, 2e-5, 256, force_full=force_full)
    assert ra["n_pushes"].sum() > 3000


@pytest.mark.parametrize("t_step", [2e-5, -1e-5])
def test_host_mirror_parity_rk4(strong_mesh, t_step):
    mesh, _, settings = strong_mesh
    run_pair(mesh, _with(settings, ipusher=1), 150, 4, t_step, 256, nsteps=2)


def test_strong_terms_change_the_orbit_and_conserve_the_extended_energy(strong_mesh, weak_mesh):
    """With the ExB terms the conserved energy is m v^2/2 + e Phi + m v_E^2/2 (supporting_functions_mod.f90:299).
    On this coarse grid a potential of this size limits energy conservation to ~1e-6 with or without the strong-
    field terms (it converges with the mesh); what is asserted is that the extended energy is conserved as well as
    the weak-field energy is in weak-field mode, and far better than the energy without the v_E^2 term."""
    mesh, _, settings = strong_mesh
    n = 100

    def go(mesh, settings, K, plain_energy_too=False):
        om = OracleMesh(mesh, _with(settings, poly_order=K))
        om_plain = OracleMesh(mesh, _with(settings, poly_order=K, boole_strong_electric_field=False))
        x, vpar, vperp = workloads.particles_cyl(n, 2)
        binit, ind, ifc = workloads.fresh_state(n)
        om.orbit_timestep_batch(x, vpar, vperp, 0.0, binit, ind, ifc)
        e0, p0, mu0 = om.invariants(x, vpar, vperp, ind)
        q0 = om_plain.invariants(x, vpar, vperp, ind)[0]
        total = 0
        for _ in range(4):
            total += om.orbit_timestep_batch(x, vpar, vperp, 5e-5, binit, ind, ifc, nthreads=4)
        e1, p1, mu1 = om.invariants(x, vpar, vperp, ind)
        q1 = om_plain.invariants(x, vpar, vperp, ind)[0]
        ok = ind > 0
        return (np.abs(e1 / e0 - 1)[ok].max(), np.abs(p1 / p0 - 1)[ok].max(), np.abs(mu1 / mu0 - 1)[ok].max(), x, ok,
                total, np.abs(q1 / q0 - 1)[ok].max())

    dE4, dP4, dMu4, xs, oks, total, dE_plain = go(mesh, settings, 4)
    assert total > 8000 and oks.sum() > 90
    dEw, dPw, _, xw, okw, _, _ = go(weak_mesh[0], weak_mesh[2], 4)
    assert dMu4 < 1e-13
    assert dE4 < 1e-5 and dE4 < 10 * dEw and dP4 < 10 * dPw
    assert dE_plain > 30 * dE4        # m v_E^2 / 2 is part of the invariant
    both = oks & okw
    assert np.abs(xs[both] - xw[both]).max() > 1e-3   # the drift terms matter at this field strength


# ---------------------------------------------------------------------------------------------- GPU
def _gpu_pair(mesh, settings, n, seed, t_step, cap, force_full=False, gather=None):
    from gorilla_b200 import Gorilla
    om, g = OracleMesh(mesh, settings), Gorilla(mesh, settings)
    g._debug_force_full(force_full)
    if gather is not None:
        g.set_gather(gather)
        assert g.get_gather() == gather
    xa, va, wa = workloads.particles_cyl(n, seed)
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    ia, ta, fa = workloads.fresh_state(n)
    ib, tb, fb = workloads.fresh_state(n)
    ra = om.orbit_timestep_trace(xa, va, wa, t_step, ia, ta, fa, cap)
    tro, npu = np.zeros(n), np.zeros(n, np.int64)
    tt, tf = g.orbit_timestep_gorilla(xb, vb, wb, t_step, ib, tb, fb, t_remain_out=tro, n_pushes=npu, trace_cap=cap)
    c = g.counters()
    assert same(ra["trace_tetr"], tt) and same(ra["trace_face"], tf), "visited tetra sequence differs"
    assert same(ra["n_pushes"], npu) and c.n_pushes == int(ra["n_pushes"].sum())
    assert same(xa, xb) and same(va, vb) and same(wa, wb) and same(ta, tb) and same(fa, fb)
    assert same(ra["t_remain"], tro)
    assert tuple(int(v) for v in ra["fallback"]) == c.n_fallback
    e, p, mu = g.invariants(xb, vb, wb, tb)
    eo, po, muo = om.invariants(xa, va, wa, ta)
    assert same(e, eo) and same(p, po) and same(mu, muo)
    g.close()
    return c


@pytest.mark.gpu
@pytest.mark.parametrize("K", [2, 3, 4])
def test_gpu_parity_polynomial(strong_mesh, cuda_device, K):
    mesh, _, settings = strong_mesh
    c = _gpu_pair(mesh, _with(settings, poly_order=K), 600, 3, 2e-5, 128)
    assert c.n_pushes > 15000
    _gpu_pair(mesh, _with(settings, poly_order=K), 200, 5, 1e-5, 64, force_full=True)


@pytest.mark.gpu
def test_gpu_parity_rk4(strong_mesh, cuda_device):
    mesh, _, settings = strong_mesh
    _gpu_pair(mesh, _with(settings, ipusher=1), 600, 4, 2e-5, 128)
    _gpu_pair(mesh, _with(settings, ipusher=1), 200, 6, -1e-5, 64, force_full=True)


@pytest.mark.gpu
@pytest.mark.parametrize("gather", [1, 2])
@pytest.mark.parametrize("pusher", ["poly2", "rk4"])
def test_gpu_parity_staged_gathers(strong_mesh, cuda_device, pusher, gather):
    """Strong-E kernels with the records staged in shared memory one push ahead: per-lane bulk copies of the magnetic record
    (1; Phi / strong-E sub-records stay per-lane loads) and the warp-cooperative gather (2), which for these kernels stages
    everything a push reads -- geom, bpart, phi and the hot part of se, 45 16-byte pieces per lane -- at two CTAs per SM.
    2048 particles = 64 full warps; forward and backward time."""
    mesh, _, settings = strong_mesh
    st = _with(settings, ipusher=1) if pusher == "rk4" else _with(settings, poly_order=2)
    c = _gpu_pair(mesh, st, 2048, 11, 2e-5, 64, gather=gather)
    assert c.n_pushes > 50000
    _gpu_pair(mesh, st, 700, 12, -1e-5, 64, gather=gather)
