"""Adaptive energy-controlled sub-stepping of the polynomial pusher (boole_adaptive_time_steps):
overhead_adaptive_time_steps / adaptive_time_steps_equidistant / _update_eta / _exit_time, pusher_tetra_poly.f90:830-1254,
and its call sites in the attempts (:316-319, 391-399, 564-568, 811-814, 2897-2902, 2966-2971).

CPU: oracle physics + oracle <-> host compile of the device headers, bit for bit.  GPU: C ABI <-> oracle, bit for bit."""
import numpy as np
import pytest

import workloads
from gorilla_b200 import api, build_mesh
from host_mirror_binding import HostMirror
from oracle_binding import OracleMesh


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def _with(settings, **kw):
    return type(settings)(**{**settings.__dict__, **kw})


def _adaptive(settings, K, dE, max_n=40, **kw):
    return _with(settings, poly_order=K, boole_adaptive_time_steps=True, desired_delta_energy=dE,
                 max_n_intermediate_steps=max_n, **kw)


@pytest.fixture(scope="module")
def strong_mesh(product_lib):
    grid, settings = workloads.analytic_tokamak(14, 14, 14)
    settings.eps_Phi = -1.5e-5
    settings.boole_strong_electric_field = True
    return build_mesh(grid, settings), grid, settings


def test_energy_error_follows_the_requested_bound(small_mesh):
    """Order 2 on the coarse test grid loses ~1e-7 of the energy in 30 crossings; with the adaptive scheme the error per
    tetrahedron is bounded by desired_delta_energy, so the accumulated error drops by orders of magnitude, while the
    visited tetrahedra stay the same where the plain order-2 orbit was already accurate."""
    mesh, _, settings = small_mesh
    n, t_step = 120, 2e-5
    res = {}
    for key, st in (("plain", _with(settings, poly_order=2)), ("1e-10", _adaptive(settings, 2, 1e-10)),
                    ("1e-13", _adaptive(settings, 2, 1e-13))):
        om = OracleMesh(mesh, st)
        x, vpar, vperp = workloads.particles_cyl(n, 1)
        s = workloads.fresh_state(n)
        om.orbit_timestep_batch(x, vpar, vperp, 0.0, *s)
        e0 = om.invariants(x, vpar, vperp, s[1])[0]
        r = om.orbit_timestep_trace(x, vpar, vperp, t_step, *s, 8)
        e1, _, mu1 = om.invariants(x, vpar, vperp, s[1])
        ok = s[1] > 0
        assert ok.sum() > 110
        res[key] = (np.abs(e1 / e0 - 1)[ok].max(), r, x.copy(), ok)
    assert res["plain"][1]["n_adaptive"] == 0 and res["1e-10"][1]["n_adaptive"] > 50
    assert res["1e-13"][1]["n_adaptive"] > 3 * res["1e-10"][1]["n_adaptive"]
    assert res["plain"][0] > 1e-8
    assert res["1e-10"][0] < 40 * 1e-10 and res["1e-10"][0] < res["plain"][0] / 30
    assert res["1e-13"][0] < res["1e-10"][0]
    both = res["plain"][3] & res["1e-13"][3]
    assert np.abs(res["plain"][2][both] - res["1e-13"][2][both]).max() < 0.05     # same orbits, better integrated
    assert same(res["plain"][1]["trace_tetr"][:, :3], res["1e-13"][1]["trace_tetr"][:, :3])


def test_settings_rules(product_lib, small_mesh):
    mesh, _, settings = small_mesh
    good = _adaptive(settings, 2, 1e-10)
    for bad, code in ((_with(good, desired_delta_energy=0.0), 1), (_with(good, max_n_intermediate_steps=1), 1),
                      (_with(good, ipusher=1), 1), (_with(good, handover_processing_kind=2), 2)):
        with pytest.raises(api.GorillaError) as ei:
            api.Gorilla(mesh, bad)
        assert ei.value.code == code


def run_pair(mesh, settings, n, seed, t_step, cap, force_full=False, nsteps=1):
    om, hm = OracleMesh(mesh, settings), HostMirror(mesh, settings)
    xa, va, wa = workloads.particles_cyl(n, seed)
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    ia, ta, fa = workloads.fresh_state(n)
    ib, tb, fb = workloads.fresh_state(n)
    nad = 0
    for _ in range(nsteps):
        ra = om.orbit_timestep_trace(xa, va, wa, t_step, ia, ta, fa, cap)
        rb = hm.orbit_timestep(xb, vb, wb, t_step, ib, tb, fb, cap, force_full=force_full)
        assert same(ra["trace_tetr"], rb["trace_tetr"]) and same(ra["trace_face"], rb["trace_face"])
        assert same(ra["n_pushes"], rb["n_pushes"])
        assert same(xa, xb) and same(va, vb) and same(wa, wb)
        assert same(ta, tb) and same(fa, fb) and same(ia, ib)
        assert same(ra["t_remain"], rb["t_remain"]) and same(ra["fallback"], rb["fallback"])
        assert (ra["n_adaptive"] > 0) == (rb["n_adaptive"] > 0)     # oracle counts segments, the device pushes
        nad += ra["n_adaptive"]
    return nad


@pytest.mark.parametrize("K,dE", [(1, 1e-10), (2, 1e-10), (2, 1e-14), (3, 1e-14), (3, 1e-16), (4, 1e-16)])
def test_host_mirror_parity(small_mesh, K, dE):
    mesh, _, settings = small_mesh
    st = _adaptive(settings, K, dE, max_n=25)
    nad = run_pair(mesh, st, 100, 3, 4e-6, 48, nsteps=2)
    if K <= 2 or dE <= 1e-16:
        assert nad > (5 if K <= 3 else 0)
    run_pair(mesh, st, 60, 4, -3e-6, 32, force_full=True)


def test_host_mirror_parity_small_step_budget(small_mesh):
    """max_n_intermediate_steps = 2..3: the partition loop ends at once / after one refinement."""
    mesh, _, settings = small_mesh
    for max_n in (2, 3):
        assert run_pair(mesh, _adaptive(settings, 2, 1e-13, max_n=max_n), 80, 9, 4e-6, 32) > 20


@pytest.mark.parametrize("K", [2, 3])
def test_host_mirror_parity_phi_and_strong_field(small_mesh_phi, strong_mesh, K):
    for mesh, _, settings in (small_mesh_phi, strong_mesh):
        assert run_pair(mesh, _adaptive(settings, K, 1e-12 if K == 2 else 1e-15, max_n=20), 80, 5, 4e-6, 48) > 0


def test_no_guess_variant(small_mesh):
    mesh, _, settings = small_mesh
    run_pair(mesh, _adaptive(settings, 3, 1e-15, max_n=20, boole_guess=False), 80, 6, 4e-6, 48)


# ---------------------------------------------------------------------------------------------- GPU
def _gpu_pair(mesh, settings, n, seed, t_step, cap, force_full=False, use_group=True):
    from gorilla_b200 import Gorilla
    om, g = OracleMesh(mesh, settings), Gorilla(mesh, settings)
    g._debug_force_full(force_full)
    g._debug_use_group(use_group)
    xa, va, wa = workloads.particles_cyl(n, seed)
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    ia, ta, fa = workloads.fresh_state(n)
    ib, tb, fb = workloads.fresh_state(n)
    for _ in range(2):
        ra = om.orbit_timestep_trace(xa, va, wa, t_step, ia, ta, fa, cap)
        tro, npu = np.zeros(n), np.zeros(n, np.int64)
        tt, tf = g.orbit_timestep_gorilla(xb, vb, wb, t_step, ib, tb, fb, t_remain_out=tro, n_pushes=npu, trace_cap=cap)
        c = g.counters()
        assert same(ra["trace_tetr"], tt) and same(ra["trace_face"], tf), "visited tetra sequence differs"
        assert same(ra["n_pushes"], npu) and c.n_pushes == int(ra["n_pushes"].sum())
        assert same(xa, xb) and same(va, vb) and same(wa, wb) and same(ta, tb) and same(fa, fb)
        assert same(ra["t_remain"], tro)
        assert tuple(int(v) for v in ra["fallback"]) == c.n_fallback
        assert (ra["n_adaptive"] > 0) == (c.n_adaptive > 0)
    g.close()
    return c


@pytest.mark.gpu
@pytest.mark.parametrize("K,dE", [(1, 1e-10), (2, 1e-10), (2, 1e-14), (3, 1e-15), (4, 1e-16)])
def test_gpu_parity(small_mesh, cuda_device, K, dE):
    mesh, _, settings = small_mesh
    st = _adaptive(settings, K, dE, max_n=25)
    c = _gpu_pair(mesh, st, 600, 3, 6e-6, 96)
    assert c.n_pushes > 4000
    if K == 2:
        assert c.n_adaptive > 100
    _gpu_pair(mesh, st, 200, 4, -3e-6, 48, force_full=True)
    if K >= 3:
        _gpu_pair(mesh, st, 256, 5, 4e-6, 48, use_group=False)


@pytest.mark.gpu
def test_gpu_parity_phi_and_strong_field(small_mesh_phi, strong_mesh, cuda_device):
    for mesh, _, settings in (small_mesh_phi, strong_mesh):
        _gpu_pair(mesh, _adaptive(settings, 2, 1e-12, max_n=20), 400, 5, 6e-6, 64)


@pytest.mark.gpu
def test_gpu_energy_conservation_at_size(small_mesh, cuda_device):
    """2e5 particles, order 2: the adaptive scheme bounds the energy error (size-independent property)."""
    from gorilla_b200 import Gorilla
    mesh, _, settings = small_mesh
    n = 200_000
    out = {}
    for key, st in (("plain", _with(settings, poly_order=2)), ("adaptive", _adaptive(settings, 2, 1e-11, max_n=40))):
        g = Gorilla(mesh, st)
        x, vpar, vperp = workloads.particles_cyl(n, 31)
        s = workloads.fresh_state(n)
        g.orbit_timestep_gorilla(x, vpar, vperp, 0.0, *s)
        e0 = g.invariants(x, vpar, vperp, s[1])[0]
        g.orbit_timestep_gorilla(x, vpar, vperp, 1e-5, *s)
        c = g.counters()
        e1 = g.invariants(x, vpar, vperp, s[1])[0]
        ok = s[1] > 0
        out[key] = (np.abs(e1 / e0 - 1)[ok].max(), c)
        g.close()
    assert out["plain"][1].n_adaptive == 0 and out["adaptive"][1].n_adaptive > 1000
    assert out["adaptive"][0] < out["plain"][0] / 20 and out["adaptive"][0] < 1e-8
