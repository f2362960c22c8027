"""Adaptive energy-controlled sub-stepping COMBINED with the consumers of the step lists: Hamiltonian
time tracing (pusher_tetra_poly.f90:463-557: the loop over number_of_integration_steps, findloc over t_hamiltonian_list in
the stop-inside case), the optional quantities (:662-667) and J_par (par_adiab_inv_tetra_poly, :3173-3291), with
tau_steps_list / intermediate_z0_list of 3 * max_n_intermediate_steps entries (manage_intermediate_steps_arrays, :98-101).
On the device these are the EXT = 5 kernels (gb_orbit_k{1..4}ax.cu): every push takes the complete path and keeps its lists
in a per-thread global scratch region.

CPU: oracle physics + oracle <-> host compile of the device headers, bit for bit.  GPU: C ABI <-> oracle, bit for bit."""
import numpy as np
import pytest

import workloads
from gorilla_b200 import api
from host_mirror_binding import HostMirror
from oracle_binding import OracleMesh

ALL_OQ = dict(boole_time_Hamiltonian=True, boole_gyrophase=True, boole_vpar_int=True, boole_vpar2_int=True)


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def _with(settings, **kw):
    return type(settings)(**{**settings.__dict__, **kw})


def _adaptive(settings, K, dE, max_n=25, **kw):
    return _with(settings, poly_order=K, boole_adaptive_time_steps=True, desired_delta_energy=dE,
                 max_n_intermediate_steps=max_n, **kw)


def _state(n):
    return np.zeros(n), np.zeros(n, np.int32), np.zeros(n, np.int32)


def _sorted(ev):
    return ev[np.lexsort((ev["kind"], ev["push"], ev["particle"]))]


# ---------------------------------------------------------------------------------------------- oracle physics
@pytest.mark.parametrize("K,dE", [(2, 1e-12), (3, 1e-15)])
def test_hamiltonian_time_of_a_sub_stepped_orbit_is_the_time_step(small_mesh, K, dE):
    """With sub-stepping a push consists of many integration steps; their Hamiltonian times still add up to the time step
    exactly (the stop-inside root is taken in the sub-step in which t_remain is reached), while with dt/dtau = const per
    cell it is only approximately so."""
    mesh, _, settings = small_mesh
    n, t_step = 80, 4e-6
    err, nad = {}, {}
    for tt in (1, 2):
        om = OracleMesh(mesh, _adaptive(settings, K, dE, i_time_tracing_option=tt, **ALL_OQ))
        x, vpar, vperp = workloads.particles_cyl(n, 5)
        st = workloads.fresh_state(n)
        r = om.orbit_timestep_trace(x, vpar, vperp, t_step, *st, 64)
        ok = st[1] > 0
        err[tt] = np.abs(r["optional_quantities"][ok, 0] / t_step - 1).max()
        nad[tt] = r["n_adaptive"]
    assert nad[1] > 3 and nad[2] > 3          # sub-stepping really happens
    assert err[2] < 1e-13 and 1e-6 < err[1] < 0.2


def test_optional_quantities_agree_with_the_plain_scheme(small_mesh):
    """The sub-stepped orbit is the same orbit, better integrated: gyrophase and the v_par integrals agree with the
    non-adaptive run to the accuracy of order 2."""
    mesh, _, settings = small_mesh
    n, t_step = 80, 4e-6
    out = {}
    for key, st in (("plain", _with(settings, poly_order=2, i_time_tracing_option=2, **ALL_OQ)),
                    ("adaptive", _adaptive(settings, 2, 1e-12, i_time_tracing_option=2, **ALL_OQ))):
        om = OracleMesh(mesh, st)
        x, vpar, vperp = workloads.particles_cyl(n, 5)
        s = workloads.fresh_state(n)
        r = om.orbit_timestep_trace(x, vpar, vperp, t_step, *s, 64)
        out[key] = (r["optional_quantities"].copy(), s[1] > 0, r["n_adaptive"])
    both = out["plain"][1] & out["adaptive"][1]
    a, b = out["plain"][0][both], out["adaptive"][0][both]
    assert out["adaptive"][2] > 20 and out["plain"][2] == 0
    assert not same(a, b)
    scale = np.abs(a).max(axis=0)
    assert (np.abs(a - b).max(axis=0) / scale).max() < 1e-4


# ---------------------------------------------------------------------------------------------- host mirror
def run_pair(mesh, settings, n, seed, t_step, cap, nsteps=1, optional=True):
    om, hm = OracleMesh(mesh, settings), HostMirror(mesh, settings)
    xa, va, wa = workloads.particles_cyl(n, seed)
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    ia, ta, fa = workloads.fresh_state(n)
    ib, tb, fb = workloads.fresh_state(n)
    nad = 0
    for _ in range(nsteps):
        ra = om.orbit_timestep_trace(xa, va, wa, t_step, ia, ta, fa, cap)
        rb = hm.orbit_timestep(xb, vb, wb, t_step, ib, tb, fb, cap, optional=optional)
        assert same(ra["trace_tetr"], rb["trace_tetr"]) and same(ra["trace_face"], rb["trace_face"])
        assert same(ra["n_pushes"], rb["n_pushes"])
        assert same(xa, xb) and same(va, vb) and same(wa, wb)
        assert same(ta, tb) and same(fa, fb) and same(ia, ib)
        assert same(ra["t_remain"], rb["t_remain"]) and same(ra["fallback"], rb["fallback"])
        if optional:
            assert same(ra["optional_quantities"], rb["optional_quantities"])
        nad += ra["n_adaptive"]
    return nad


@pytest.mark.parametrize("K,dE", [(1, 1e-10), (2, 1e-12), (3, 1e-15), (4, 1e-16)])
@pytest.mark.parametrize("tt", [1, 2])
def test_host_mirror_parity(small_mesh, K, dE, tt):
    mesh, _, settings = small_mesh
    st = _adaptive(settings, K, dE, i_time_tracing_option=tt, **ALL_OQ)
    nad = run_pair(mesh, st, 80, 5, 4e-6, 48, nsteps=2)
    if K <= 3:
        assert nad > 0
    run_pair(mesh, st, 40, 6, -3e-6, 32)                                   # backward time


def test_host_mirror_parity_hamiltonian_time_only_and_small_budget(small_mesh):
    """Hamiltonian time tracing without optional quantities through the plain entry point; max_n_intermediate_steps small
    enough for partitions that end on the budget."""
    mesh, _, settings = small_mesh
    run_pair(mesh, _adaptive(settings, 3, 1e-15, i_time_tracing_option=2), 60, 8, 4e-6, 48, optional=False)
    for max_n in (3, 6):
        assert run_pair(mesh, _adaptive(settings, 2, 1e-13, max_n=max_n, i_time_tracing_option=2, **ALL_OQ), 60, 9, 4e-6, 32) > 10


def test_host_mirror_parity_phi_and_strong_field(small_mesh_phi, product_lib):
    from gorilla_b200 import build_mesh
    grid, settings = workloads.analytic_tokamak(14, 14, 14)
    settings.eps_Phi = -1.5e-5
    settings.boole_strong_electric_field = True
    strong = (build_mesh(grid, settings), grid, settings)
    for mesh, _, st0 in (small_mesh_phi, strong):
        assert run_pair(mesh, _adaptive(st0, 2, 1e-12, i_time_tracing_option=2, **ALL_OQ), 60, 5, 4e-6, 48) > 0


@pytest.mark.parametrize("K,dE", [(2, 1e-11), (3, 1e-14), (4, 1e-16)])
def test_host_mirror_parity_events(small_mesh, K, dE):
    """J_par / banana tips / toroidal mappings over the long lists (the turning step is found by findloc over all sub-steps)."""
    mesh, _, settings = small_mesh
    st = _adaptive(settings, K, dE)
    om, hm = OracleMesh(mesh, st), HostMirror(mesh, st)
    n = 40
    xa, va, wa = workloads.particles_cyl(n, 5)
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    sa, sb = workloads.fresh_state(n), workloads.fresh_state(n)
    Ja, cva, cpa = _state(n)
    Jb, cvb, cpb = _state(n)
    for _ in range(2):
        eva, nea, npa = om.orbit_timestep_events(xa, va, wa, 4e-4, *sa, Ja, cva, cpa, 100000, n_skip_phi_0=2)
        evb, neb, npb = hm.orbit_timestep_events(xb, vb, wb, 4e-4, *sb, Jb, cvb, cpb, 100000, n_skip_phi_0=2)
        assert nea == neb and np.array_equal(eva, evb)
        assert np.array_equal(Ja, Jb) and np.array_equal(cva, cvb) and np.array_equal(cpa, cpb)
        assert np.array_equal(xa, xb) and np.array_equal(va, vb) and np.array_equal(wa, wb) and np.array_equal(npa, npb)
    assert (eva["kind"] == 2).sum() > 10 and (eva["kind"] == 1).sum() > 30


# ---------------------------------------------------------------------------------------------- GPU
def _gpu_pair(mesh, settings, n, seed, t_step, cap, optional=True):
    from gorilla_b200 import Gorilla
    om, g = OracleMesh(mesh, settings), Gorilla(mesh, settings)
    xa, va, wa = workloads.particles_cyl(n, seed)
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    ia, ta, fa = workloads.fresh_state(n)
    ib, tb, fb = workloads.fresh_state(n)
    nad = 0
    for _ in range(2):
        ra = om.orbit_timestep_trace(xa, va, wa, t_step, ia, ta, fa, cap)
        tro, npu = np.zeros(n), np.zeros(n, np.int64)
        oq = np.full((n, 4), 3.0) if optional else None
        tt, tf = g.orbit_timestep_gorilla(xb, vb, wb, t_step, ib, tb, fb, t_remain_out=tro, n_pushes=npu, trace_cap=cap,
                                          optional_quantities=oq)
        c = g.counters()
        assert same(ra["trace_tetr"], tt) and same(ra["trace_face"], tf), "visited tetra sequence differs"
        assert same(ra["n_pushes"], npu) and c.n_pushes == int(ra["n_pushes"].sum())
        assert same(xa, xb) and same(va, vb) and same(wa, wb) and same(ta, tb) and same(fa, fb)
        assert same(ra["t_remain"], tro)
        assert tuple(int(v) for v in ra["fallback"]) == c.n_fallback
        assert (ra["n_adaptive"] > 0) == (c.n_adaptive > 0)
        if optional:
            assert same(ra["optional_quantities"], oq), "optional quantities differ"
        nad += c.n_adaptive
    g.close()
    return nad


@pytest.mark.gpu
@pytest.mark.parametrize("K,dE", [(1, 1e-10), (2, 1e-12), (3, 1e-15), (4, 1e-16)])
@pytest.mark.parametrize("tt", [1, 2])
def test_gpu_parity(small_mesh, cuda_device, K, dE, tt):
    mesh, _, settings = small_mesh
    nad = _gpu_pair(mesh, _adaptive(settings, K, dE, i_time_tracing_option=tt, **ALL_OQ), 600, 5, 5e-6, 64)
    if K <= 3:
        assert nad > 20


@pytest.mark.gpu
def test_gpu_parity_hamiltonian_time_only_phi_and_default_budget(small_mesh, small_mesh_phi, cuda_device):
    mesh, _, settings = small_mesh
    _gpu_pair(mesh, _adaptive(settings, 3, 1e-15, i_time_tracing_option=2), 400, 8, 5e-6, 48, optional=False)
    mesh, _, settings = small_mesh_phi
    # max_n_intermediate_steps = 10000 (the default of gorilla.inp): 30 000 list entries per thread
    assert _gpu_pair(mesh, _adaptive(settings, 2, 1e-12, max_n=10000, i_time_tracing_option=2, **ALL_OQ), 300, 5, 4e-6, 48) > 0


@pytest.mark.gpu
@pytest.mark.parametrize("K,dE", [(2, 1e-11), (4, 1e-16)])
def test_gpu_parity_events(small_mesh, cuda_device, K, dE):
    from gorilla_b200 import Gorilla
    mesh, _, settings = small_mesh
    st = _adaptive(settings, K, dE)
    om, g = OracleMesh(mesh, st), Gorilla(mesh, st)
    n = 160
    xa, va, wa = workloads.particles_cyl(n, 5)
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    sa, sb = workloads.fresh_state(n), workloads.fresh_state(n)
    Ja, cva, cpa = _state(n)
    Jb, cvb, cpb = _state(n)
    for _ in range(2):
        eva, nea, npa = om.orbit_timestep_events(xa, va, wa, 4e-4, *sa, Ja, cva, cpa, 400000, n_skip_phi_0=2)
        npb = np.zeros(n, np.int64)
        evb, neb = g.orbit_timestep_gorilla_events(xb, vb, wb, 4e-4, *sb, Jb, cvb, cpb, 400000, n_pushes=npb, n_skip_phi_0=2)
        assert nea == neb and np.array_equal(_sorted(eva), evb)
        assert np.array_equal(Ja, Jb) and np.array_equal(cva, cvb) and np.array_equal(cpa, cpb)
        assert np.array_equal(xa, xb) and np.array_equal(va, vb) and np.array_equal(wa, wb) and np.array_equal(npa, npb)
    assert (eva["kind"] == 2).sum() > 30
    g.close()


@pytest.mark.gpu
def test_gpu_refuses_lists_beyond_the_scratch_limit(small_mesh, cuda_device):
    from gorilla_b200 import Gorilla
    mesh, _, settings = small_mesh
    g = Gorilla(mesh, _adaptive(settings, 2, 1e-12, max_n=200000, i_time_tracing_option=2))
    n = 200000
    x, vpar, vperp = workloads.particles_cyl(n, 1)
    with pytest.raises(api.GorillaError) as ei:
        g.orbit_timestep_gorilla(x, vpar, vperp, 1e-6, *workloads.fresh_state(n))
    assert ei.value.code == 2
    g.close()
