"""Hamiltonian time tracing (i_time_tracing_option = 2) and the optional quantities of pusher_tetra_poly
(SURVEY.md 8f row 2): pusher_tetra_poly.f90:463-557 (time tracing / stop-inside root in Hamiltonian time), :621-631,
:662-667, :2117-2536 (calc_t_hamiltonian, get_t_hamiltonian_root, calc_optional_quantities, z_series_coef,
poly_multiplication_coef), :3000-3150 (moment_integration); hamiltonian_time record tetra_physics_mod.f90:105-114,926-944.

CPU: oracle physics + oracle <-> host compile of the device headers, bit for bit.  GPU: C ABI <-> oracle, bit for bit."""
import numpy as np
import pytest

import workloads
from gorilla_b200 import api, build_mesh
from host_mirror_binding import HostMirror
from oracle_binding import OracleMesh

ALL_OQ = dict(boole_time_Hamiltonian=True, boole_gyrophase=True, boole_vpar_int=True, boole_vpar2_int=True)


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def _with(settings, **kw):
    return type(settings)(**{**settings.__dict__, **kw})


@pytest.fixture(scope="module")
def strong_mesh(product_lib):
    grid, settings = workloads.analytic_tokamak(14, 14, 14)
    settings.eps_Phi = -1.5e-5
    settings.boole_strong_electric_field = True
    return build_mesh(grid, settings), grid, settings


# ---------------------------------------------------------------------------------------------- oracle physics
def test_settings_rules(product_lib, small_mesh):
    mesh, _, settings = small_mesh
    for bad in (_with(settings, i_time_tracing_option=2, ipusher=1),       # gorilla_settings_mod.f90:124-129
                _with(settings, boole_gyrophase=True),                     # :132-135
                _with(settings, i_time_tracing_option=3)):
        with pytest.raises(api.GorillaError):
            api.Gorilla(mesh, bad)


@pytest.mark.parametrize("K", [2, 3, 4])
def test_hamiltonian_time_is_order_consistent(small_mesh, K):
    """With Hamiltonian time tracing the accumulated t_hamiltonian of a completed step IS the time step (up to the
    5th-order term dropped in the final cell at order 4, pusher_tetra_poly.f90:587-594); with dt/dtau = const per cell it
    is only approximately so.  Gyrophase ~ -omega_c t, int v_par dt ~ v_par t."""
    mesh, _, settings = small_mesh
    n, t_step = 150, 1e-5
    out = {}
    for tt in (1, 2):
        om = OracleMesh(mesh, _with(settings, poly_order=K, i_time_tracing_option=tt, **ALL_OQ))
        x, vpar, vperp = workloads.particles_cyl(n, 11)
        v0 = vpar.copy()
        binit, ind, ifc = workloads.fresh_state(n)
        oq, tro = np.zeros((n, 4)), np.zeros(n)
        om.orbit_timestep_batch(x, vpar, vperp, t_step, binit, ind, ifc, t_remain_out=tro, optional_quantities=oq, nthreads=4)
        ok = ind > 0
        assert ok.sum() > 140 and np.all(tro[ok] == 0.0)
        out[tt] = (x, oq, ok, v0, vpar)
    x1, oq1, ok1, _, _ = out[1]
    x2, oq2, ok2, v0, v1 = out[2]
    both = ok1 & ok2
    err2 = np.abs(oq2[both, 0] / t_step - 1.0).max()
    err1 = np.abs(oq1[both, 0] / t_step - 1.0).max()
    assert err2 < (1e-8 if K == 4 else 1e-13) and 1e-4 < err1 < 0.1
    assert 1e-3 < np.abs(x1[both] - x2[both]).max() < 2.0          # the orbits differ at the size of the time error
    # deuterons in ~2e4 G: omega_c = e B / (m c) ~ 1e8 rad/s
    om_c = -oq2[both, 1] / t_step
    assert np.all((om_c > 5e7) & (om_c < 2e8))
    vmean = 0.5 * (v0 + v1)[both]
    big = np.abs(vmean) > 1e7
    assert np.abs(oq2[both, 2][big] / (vmean[big] * t_step) - 1.0).max() < 0.1
    assert np.all(oq2[both, 3] >= 0.0)
    assert np.abs(oq2[both, 3][big] / (vmean[big] ** 2 * t_step) - 1.0).max() < 0.3


def test_optional_quantities_are_switchable(small_mesh):
    mesh, _, settings = small_mesh
    n = 40
    ref = None
    for flags in (ALL_OQ, dict(boole_vpar_int=True), dict(boole_time_Hamiltonian=True, boole_gyrophase=True), {}):
        om = OracleMesh(mesh, _with(settings, poly_order=2, **flags))
        x, vpar, vperp = workloads.particles_cyl(n, 12)
        binit, ind, ifc = workloads.fresh_state(n)
        oq = np.full((n, 4), 7.0)
        om.orbit_timestep_batch(x, vpar, vperp, 4e-6, binit, ind, ifc, optional_quantities=oq)
        if ref is None:
            ref = oq.copy()
            assert np.all(ref[ind > 0] != 0.0)
        on = [bool(flags.get(k)) for k in ("boole_time_Hamiltonian", "boole_gyrophase", "boole_vpar_int", "boole_vpar2_int")]
        for q in range(4):
            assert same(oq[:, q], ref[:, q] if on[q] else np.zeros(n))


# ---------------------------------------------------------------------------------------------- host mirror parity
def run_pair(mesh, settings, n, seed, t_step, cap, force_full=False, nsteps=1, sampler=None):
    om, hm = OracleMesh(mesh, settings), HostMirror(mesh, settings)
    xa, va, wa = (sampler or workloads.particles_cyl)(n, seed)
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    ia, ta, fa = workloads.fresh_state(n)
    ib, tb, fb = workloads.fresh_state(n)
    for _ in range(nsteps):
        ra = om.orbit_timestep_trace(xa, va, wa, t_step, ia, ta, fa, cap)
        rb = hm.orbit_timestep(xb, vb, wb, t_step, ib, tb, fb, cap, force_full=force_full, optional=True)
        assert same(ra["trace_tetr"], rb["trace_tetr"]) and same(ra["trace_face"], rb["trace_face"])
        assert same(ra["n_pushes"], rb["n_pushes"])
        assert same(xa, xb) and same(va, vb) and same(wa, wb)
        assert same(ta, tb) and same(fa, fb) and same(ia, ib)
        assert same(ra["t_remain"], rb["t_remain"]) and same(ra["fallback"], rb["fallback"])
        assert same(ra["optional_quantities"], rb["optional_quantities"])
    return ra


@pytest.mark.parametrize("K", [1, 2, 3, 4])
@pytest.mark.parametrize("tt", [1, 2])
def test_host_mirror_parity(small_mesh, K, tt):
    mesh, _, settings = small_mesh
    st = _with(settings, poly_order=K, i_time_tracing_option=tt, **ALL_OQ)
    ra = run_pair(mesh, st, 120, 3, 3e-6, 64, nsteps=2)
    assert ra["n_pushes"].sum() > 800 and np.count_nonzero(ra["optional_quantities"][:, 0]) > 100
    run_pair(mesh, st, 60, 4, -2e-6, 32, force_full=True)
    if K == 1:   # order 1 turns on faces all the time: two integration steps per push (prolonged trajectory)
        assert ra["fallback"][2] > 0


@pytest.mark.parametrize("K", [2, 4])
def test_host_mirror_parity_phi_and_strong_field(small_mesh_phi, strong_mesh, K):
    for mesh, _, settings in (small_mesh_phi, strong_mesh):
        st = _with(settings, poly_order=K, i_time_tracing_option=2, **ALL_OQ)
        ra = run_pair(mesh, st, 100, 5, 4e-6, 64)
        assert ra["n_pushes"].sum() > 500
        run_pair(mesh, _with(st, i_time_tracing_option=1, boole_gyrophase=False, boole_vpar2_int=False), 60, 6, 4e-6, 32,
                 force_full=True)


def test_host_mirror_parity_without_optional_quantities(small_mesh):
    """Hamiltonian time tracing alone (no optional quantity requested) through the plain entry point."""
    mesh, _, settings = small_mesh
    st = _with(settings, poly_order=3, i_time_tracing_option=2)
    om, hm = OracleMesh(mesh, st), HostMirror(mesh, st)
    xa, va, wa = workloads.particles_cyl(100, 8)
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    sa, sb = workloads.fresh_state(100), workloads.fresh_state(100)
    ra = om.orbit_timestep_trace(xa, va, wa, 5e-6, *sa, 48)
    rb = hm.orbit_timestep(xb, vb, wb, 5e-6, *sb, 48)
    assert same(ra["trace_tetr"], rb["trace_tetr"]) and same(xa, xb) and same(va, vb) and same(wa, wb)
    assert same(ra["t_remain"], rb["t_remain"]) and rb["optional_quantities"] is None


# ---------------------------------------------------------------------------------------------- GPU
def _gpu_pair(mesh, settings, n, seed, t_step, cap, force_full=False, optional=True, use_group=True):
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
        oq = np.full((n, 4), 3.0) if optional else None
        tt, tf = g.orbit_timestep_gorilla(xb, vb, wb, t_step, ib, tb, fb, t_remain_out=tro, n_pushes=npu, trace_cap=cap,
                                          optional_quantities=oq)
        c = g.counters()
        assert same(ra["trace_tetr"], tt) and same(ra["trace_face"], tf), "visited tetra sequence differs"
        assert same(ra["n_pushes"], npu) and c.n_pushes == int(ra["n_pushes"].sum())
        assert same(xa, xb) and same(va, vb) and same(wa, wb) and same(ta, tb) and same(fa, fb)
        assert same(ra["t_remain"], tro)
        assert tuple(int(v) for v in ra["fallback"]) == c.n_fallback
        if optional:
            assert same(ra["optional_quantities"], oq), "optional quantities differ"
    g.close()
    return c


@pytest.mark.gpu
@pytest.mark.parametrize("K", [1, 2, 3, 4])
@pytest.mark.parametrize("tt", [1, 2])
def test_gpu_parity(small_mesh, cuda_device, K, tt):
    mesh, _, settings = small_mesh
    st = _with(settings, poly_order=K, i_time_tracing_option=tt, **ALL_OQ)
    c = _gpu_pair(mesh, st, 700, 3, 6e-6, 96)
    assert c.n_pushes > 5000
    _gpu_pair(mesh, st, 200, 4, -3e-6, 48, force_full=True)
    if K >= 3:   # one-particle-per-lane kernel instead of the lock-step kernel
        _gpu_pair(mesh, st, 300, 5, 4e-6, 48, use_group=False)


@pytest.mark.gpu
def test_gpu_parity_time_tracing_only(small_mesh, cuda_device):
    mesh, _, settings = small_mesh
    for K in (2, 4):
        _gpu_pair(mesh, _with(settings, poly_order=K, i_time_tracing_option=2), 500, 7, 6e-6, 64, optional=False)


@pytest.mark.gpu
@pytest.mark.parametrize("K", [2, 4])
def test_gpu_parity_phi_and_strong_field(small_mesh_phi, strong_mesh, cuda_device, K):
    for mesh, _, settings in (small_mesh_phi, strong_mesh):
        _gpu_pair(mesh, _with(settings, poly_order=K, i_time_tracing_option=2, **ALL_OQ), 400, 5, 6e-6, 64)


@pytest.mark.gpu
def test_gpu_optional_dev_entry_point_and_properties(small_mesh, cuda_device):
    """Device-pointer entry point at a size the oracle does not walk: the Hamiltonian time of every completed step equals
    the time step, skipped (lost / not initialised) particles report zeros, rk pusher refuses."""
    import torch
    from gorilla_b200 import Gorilla
    mesh, _, settings = small_mesh
    st = _with(settings, poly_order=2, i_time_tracing_option=2, **ALL_OQ)
    g = Gorilla(mesh, st)
    n, t_step = 100_000, 4e-6
    x, vpar, vperp = workloads.particles_cyl(n, 21)
    dev = cuda_device
    tx, tv, tw = (torch.from_numpy(a).to(dev) for a in (x, vpar, vperp))
    binit = torch.zeros(n, dtype=torch.int32, device=dev)
    ind = torch.full((n,), -1, dtype=torch.int32, device=dev)
    ifc = torch.full((n,), -1, dtype=torch.int32, device=dev)
    oq = torch.full((n, 4), 5.0, dtype=torch.float64, device=dev)
    tro = torch.zeros(n, dtype=torch.float64, device=dev)
    g.orbit_timestep_gorilla_optional_dev(tx, tv, tw, t_step, binit, ind, ifc, oq, t_remain_out=tro)
    torch.cuda.synchronize()
    ok = (ind > 0).cpu().numpy()
    oqh = oq.cpu().numpy()
    assert ok.sum() > 0.95 * n
    assert np.abs(oqh[ok, 0] / t_step - 1.0).max() < 1e-12
    assert np.all(oqh[ok, 1] < 0.0) and np.all(oqh[ok, 3] >= 0.0)
    # second call: lost particles are skipped and report zeros
    g.orbit_timestep_gorilla_optional_dev(tx, tv, tw, t_step, binit, ind, ifc, oq)
    torch.cuda.synchronize()
    lost = ~ok
    if lost.any():
        assert np.all(oq.cpu().numpy()[lost] == 0.0)
    g.close()
    g1 = Gorilla(mesh, _with(settings, ipusher=1))
    with pytest.raises(api.GorillaError):
        g1.orbit_timestep_gorilla(x, vpar, vperp, t_step, *workloads.fresh_state(n), optional_quantities=np.zeros((n, 4)))
    g1.close()
