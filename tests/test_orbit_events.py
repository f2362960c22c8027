"""Orbit events: toroidal (phi = 0) mappings and banana tips / parallel adiabatic invariant J_par as
the reference's plotting driver captures them -- gorilla_plot_orbit_integration, gorilla_plot_mod.f90:433-658 (events
:585-638), modules par_adiab_inv_poly_mod (pusher_tetra_poly.f90:3156-3429) and par_adiab_inv_rk_mod (pusher_tetra_rk.f90:
2589-2798) -- written to an event buffer instead of files.

CPU: oracle physics + oracle <-> host compile of the device headers, bit for bit.  GPU: C ABI <-> oracle, bit for bit."""
import numpy as np
import pytest

import workloads
from gorilla_b200 import api
from host_mirror_binding import HostMirror
from oracle_binding import OracleMesh


def _with(settings, **kw):
    return type(settings)(**{**settings.__dict__, **kw})


def _sorted(ev):
    return ev[np.lexsort((ev["kind"], ev["push"], ev["particle"]))]


def _state(n):
    return np.zeros(n), np.zeros(n, np.int32), np.zeros(n, np.int32)


@pytest.mark.parametrize("K", [3, 4])
def test_j_par_is_conserved_and_mappings_count_toroidal_turns(small_mesh, K):
    """Trapped deuterons in the axisymmetric test field: J_par of successive complete bounces agrees to the accuracy of the
    polynomial order; passing particles produce one mapping per toroidal turn, with the sign of the direction."""
    mesh, _, settings = small_mesh
    om = OracleMesh(mesh, _with(settings, poly_order=K))
    n = 60
    x, vpar, vperp = workloads.particles_cyl(n, 1)
    v0 = vpar.copy()
    st = workloads.fresh_state(n)
    J, cv, cp = _state(n)
    ev, nev, npush = om.orbit_timestep_events(x, vpar, vperp, 1.2e-3, *st, J, cv, cp, 200000)
    assert nev == len(ev) and nev > 300
    tips = ev[ev["kind"] == 2]
    assert len(tips) > 40
    checked = 0
    for p in np.unique(tips["particle"]):
        j = tips["value"][tips["particle"] == p, 0]
        if len(j) >= 3:
            assert np.ptp(j) / np.abs(j).mean() < 2e-4        # J_par of complete bounces
            e = tips["value"][tips["particle"] == p, 1]
            assert np.ptp(e) / np.abs(e).mean() < 1e-9       # total energy at the tips
            checked += 1
    assert checked >= 5
    # banana tips: counters are consecutive from 2 (the first two bounces are not reported, :3240), v_par changes sign there
    for p in np.unique(tips["particle"]):
        c = tips["counter"][tips["particle"] == p]
        assert c[0] == 2 and np.all(np.diff(c) == 1) and cv[p] == c[-1] + 1
    maps = ev[ev["kind"] == 1]
    passing = [p for p in range(n) if cv[p] == 0 and st[1][p] > 0]
    assert len(passing) > 10
    orient = set()
    for p in passing:
        c = maps["counter"][maps["particle"] == p]
        assert len(c) == abs(cp[p]) and np.all(np.abs(np.diff(c)) == 1)
        if cp[p] != 0:
            orient.add(int(np.sign(cp[p]) * np.sign(v0[p])))
    assert len(orient) == 1            # co- and counter-passing particles map in opposite directions
    # the mapping position is the hand-over point: phi sits on the period boundary (0 after the periodic shift, or 2 pi)
    ph = maps["x"][:, 1]
    assert np.all((np.abs(ph) < 1e-9) | (np.abs(ph - 2 * np.pi) < 1e-9))


def test_skip_counters_and_switches(small_mesh):
    mesh, _, settings = small_mesh
    om = OracleMesh(mesh, _with(settings, poly_order=2))
    n = 30
    out = {}
    for key, kw in (("all", {}), ("skip", dict(n_skip_phi_0=3, n_skip_vpar_0=2)), ("phi_only", dict(poincare_vpar_0=False, J_par=False)),
                    ("vpar_only", dict(poincare_phi_0=False))):
        x, vpar, vperp = workloads.particles_cyl(n, 2)
        st = workloads.fresh_state(n)
        J, cv, cp = _state(n)
        ev, nev, _ = om.orbit_timestep_events(x, vpar, vperp, 8e-4, *st, J, cv, cp, 100000, **kw)
        out[key] = (_sorted(ev), cv.copy(), cp.copy(), x.copy())
    ev_all = out["all"][0]
    sk = out["skip"][0]
    want = ev_all[((ev_all["kind"] == 1) & (ev_all["counter"] % 3 == 0)) | ((ev_all["kind"] == 2) & (ev_all["counter"] % 2 == 0))]
    assert np.array_equal(sk, want)
    assert np.array_equal(out["phi_only"][0], ev_all[ev_all["kind"] == 1]) and np.all(out["phi_only"][1] == 0)
    assert np.array_equal(out["vpar_only"][0], ev_all[ev_all["kind"] == 2])
    assert np.array_equal(out["vpar_only"][2], out["all"][2])          # the toroidal counter runs regardless (:601-605)
    for k in out:
        assert np.array_equal(out[k][3], out["all"][3])                # capturing events never changes the orbit


@pytest.mark.parametrize("K", [2, 3, 4])
@pytest.mark.parametrize("force_full", [False, True])
def test_host_mirror_parity(small_mesh, K, force_full):
    mesh, _, settings = small_mesh
    st = _with(settings, poly_order=K)
    om, hm = OracleMesh(mesh, st), HostMirror(mesh, st)
    n = 40
    xa, va, wa = workloads.particles_cyl(n, 5)
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    sa, sb = workloads.fresh_state(n), workloads.fresh_state(n)
    Ja, cva, cpa = _state(n)
    Jb, cvb, cpb = _state(n)
    for _ in range(2):   # state carried across two calls
        eva, nea, npa = om.orbit_timestep_events(xa, va, wa, 5e-4, *sa, Ja, cva, cpa, 100000, n_skip_phi_0=2)
        evb, neb, npb = hm.orbit_timestep_events(xb, vb, wb, 5e-4, *sb, Jb, cvb, cpb, 100000, n_skip_phi_0=2,
                                                 force_full=force_full)
        assert nea == neb and np.array_equal(eva, evb)
        assert np.array_equal(Ja, Jb) and np.array_equal(cva, cvb) and np.array_equal(cpa, cpb)
        assert np.array_equal(xa, xb) and np.array_equal(va, vb) and np.array_equal(wa, wb) and np.array_equal(npa, npb)
    assert (eva["kind"] == 2).sum() > 10 and (eva["kind"] == 1).sum() > 50


def test_host_mirror_parity_strong_field_and_backward_time(product_lib):
    from gorilla_b200 import build_mesh
    grid, settings = workloads.analytic_tokamak(14, 14, 14)
    settings.eps_Phi = -1.5e-5
    settings.boole_strong_electric_field = True
    mesh = build_mesh(grid, settings)
    st = _with(settings, poly_order=2)
    om, hm = OracleMesh(mesh, st), HostMirror(mesh, st)
    n = 30
    for t_step in (4e-4, -4e-4):
        xa, va, wa = workloads.particles_cyl(n, 6)
        xb, vb, wb = xa.copy(), va.copy(), wa.copy()
        sa, sb = workloads.fresh_state(n), workloads.fresh_state(n)
        Ja, cva, cpa = _state(n)
        Jb, cvb, cpb = _state(n)
        eva, nea, _ = om.orbit_timestep_events(xa, va, wa, t_step, *sa, Ja, cva, cpa, 100000)
        evb, neb, _ = hm.orbit_timestep_events(xb, vb, wb, t_step, *sb, Jb, cvb, cpb, 100000)
        assert nea == neb and nea > 50 and np.array_equal(eva, evb)
        assert np.array_equal(Ja, Jb) and np.array_equal(cva, cvb) and np.array_equal(cpa, cpb) and np.array_equal(xa, xb)


def test_refused_configurations(product_lib, small_mesh):
    mesh, _, settings = small_mesh
    om = OracleMesh(mesh, _with(settings, poly_order=1))
    x, vpar, vperp = workloads.particles_cyl(2, 1)
    st = workloads.fresh_state(2)
    J, cv, cp = _state(2)
    with pytest.raises(AssertionError):     # GOR_ERR_CONFIG: par_adiab_tau has no case(1)
        om.orbit_timestep_events(x, vpar, vperp, 1e-5, *st, J, cv, cp, 10)


# ---------------------------------------------------------------------------------------------- RK pusher
def test_rk_pusher_j_par_agrees_with_the_polynomial_pusher(small_mesh):
    """module par_adiab_inv_rk_mod (SRC/pusher_tetra_rk.f90:2589-2798): J_par integrated as a fifth ODE45 equation.  The
    same banana tips are found as with the order-4 polynomial pusher, J_par agrees to the accuracy of the two pushers."""
    mesh, _, settings = small_mesh
    res = {}
    for key, kw in (("poly", dict(ipusher=2, poly_order=4)), ("rk", dict(ipusher=1))):
        om = OracleMesh(mesh, _with(settings, **kw))
        n = 40
        x, vpar, vperp = workloads.particles_cyl(n, 5)
        lam = np.linspace(-0.3, 0.3, n)                 # deeply trapped: several bounces within the time step
        vmod = np.hypot(vpar, vperp)
        vpar[:] = lam * vmod
        vperp[:] = np.sqrt(vmod ** 2 - vpar ** 2)
        st = workloads.fresh_state(n)
        J, cv, cp = _state(n)
        ev, nev, _ = om.orbit_timestep_events(x, vpar, vperp, 6e-4, *st, J, cv, cp, 100000)
        res[key] = (ev[ev["kind"] == 2], cv.copy(), J.copy())
    a, b = res["poly"][0], res["rk"][0]
    assert len(a) > 20 and np.array_equal(res["poly"][1], res["rk"][1])
    da = {(int(e["particle"]), int(e["counter"])): e for e in a}
    db = {(int(e["particle"]), int(e["counter"])): e for e in b}
    assert set(da) == set(db)
    rel = [abs(da[k]["value"][0] / db[k]["value"][0] - 1) for k in da]
    assert max(rel) < 1e-7
    pos = [np.abs(da[k]["x"] - db[k]["x"]).max() for k in da]
    assert max(pos) < 1e-4                               # |v_par| <= 10 cm/s (RK) vs the exact root (polynomial)
    for p in np.unique(b["particle"]):
        j = b["value"][b["particle"] == p, 0]
        if len(j) >= 3:
            assert np.ptp(j) / np.abs(j).mean() < 2e-4


@pytest.mark.parametrize("force_full", [False, True])
def test_rk_pusher_host_mirror_parity(small_mesh, force_full):
    mesh, _, settings = small_mesh
    st = _with(settings, ipusher=1)
    om, hm = OracleMesh(mesh, st), HostMirror(mesh, st)
    n = 40
    xa, va, wa = workloads.particles_cyl(n, 5)
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    sa, sb = workloads.fresh_state(n), workloads.fresh_state(n)
    Ja, cva, cpa = _state(n)
    Jb, cvb, cpb = _state(n)
    for _ in range(2):
        eva, nea, npa = om.orbit_timestep_events(xa, va, wa, 5e-4, *sa, Ja, cva, cpa, 100000, n_skip_phi_0=2)
        evb, neb, npb = hm.orbit_timestep_events(xb, vb, wb, 5e-4, *sb, Jb, cvb, cpb, 100000, n_skip_phi_0=2,
                                                 force_full=force_full)
        assert nea == neb and np.array_equal(eva, evb)      # host pow() on both sides: bit for bit
        assert np.array_equal(Ja, Jb) and np.array_equal(cva, cvb) and np.array_equal(cpa, cpb)
        assert np.array_equal(xa, xb) and np.array_equal(va, vb) and np.array_equal(wa, wb) and np.array_equal(npa, npb)
    assert (eva["kind"] == 2).sum() > 10 and (eva["kind"] == 1).sum() > 50


@pytest.mark.gpu
def test_rk_pusher_gpu_parity(small_mesh, small_mesh_phi, cuda_device):
    """The orbit itself is bit-identical; J_par goes through RKF45's pow(x, 0.2) step-size control, where the CUDA libm is not
    glibc's to the last bit: values to 1e-10 (north_star's bound), everything discrete exactly."""
    from gorilla_b200 import Gorilla
    for mesh, _, settings in (small_mesh, small_mesh_phi):
        st = _with(settings, ipusher=1)
        om, g = OracleMesh(mesh, st), Gorilla(mesh, st)
        n = 160
        xa, va, wa = workloads.particles_cyl(n, 5)
        xb, vb, wb = xa.copy(), va.copy(), wa.copy()
        sa, sb = workloads.fresh_state(n), workloads.fresh_state(n)
        Ja, cva, cpa = _state(n)
        Jb, cvb, cpb = _state(n)
        for _ in range(2):
            eva, nea, npa = om.orbit_timestep_events(xa, va, wa, 5e-4, *sa, Ja, cva, cpa, 400000, n_skip_phi_0=2)
            npb = np.zeros(n, np.int64)
            evb, neb = g.orbit_timestep_gorilla_events(xb, vb, wb, 5e-4, *sb, Jb, cvb, cpb, 400000, n_pushes=npb, n_skip_phi_0=2)
            eva = _sorted(eva)
            assert nea == neb and nea > 500
            for f in ("particle", "push", "kind", "counter"):
                assert np.array_equal(eva[f], evb[f]), f
            assert np.allclose(eva["x"], evb["x"], rtol=1e-10, atol=1e-12)
            assert np.allclose(eva["value"], evb["value"], rtol=1e-10, atol=0)
            assert np.allclose(Ja, Jb, rtol=1e-10, atol=1e-300) and np.array_equal(cva, cvb) and np.array_equal(cpa, cpb)
            assert np.array_equal(xa, xb) and np.array_equal(va, vb) and np.array_equal(wa, wb) and np.array_equal(npa, npb)
        assert (eva["kind"] == 2).sum() > 30
        g.close()


# ---------------------------------------------------------------------------------------------- GPU
def _gpu_pair(mesh, settings, n, seed, t_step, ncalls=2, cap=400000, use_group=True, **kw):
    from gorilla_b200 import Gorilla
    om, g = OracleMesh(mesh, settings), Gorilla(mesh, settings)
    g._debug_use_group(use_group)
    xa, va, wa = workloads.particles_cyl(n, seed)
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    sa, sb = workloads.fresh_state(n), workloads.fresh_state(n)
    Ja, cva, cpa = _state(n)
    Jb, cvb, cpb = _state(n)
    total = 0
    for _ in range(ncalls):
        eva, nea, npa = om.orbit_timestep_events(xa, va, wa, t_step, *sa, Ja, cva, cpa, cap, **kw)
        npb = np.zeros(n, np.int64)
        evb, neb = g.orbit_timestep_gorilla_events(
            xb, vb, wb, t_step, *sb, Jb, cvb, cpb, cap, n_pushes=npb,
            **{("boole_" + k if k in ("poincare_phi_0", "poincare_vpar_0", "J_par", "full_orbit") else k): v for k, v in kw.items()})
        assert nea == neb, "number of events differs"
        assert np.array_equal(_sorted(eva), evb), "events differ"
        assert np.array_equal(Ja, Jb) and np.array_equal(cva, cvb) and np.array_equal(cpa, cpb)
        assert np.array_equal(xa, xb) and np.array_equal(va, vb) and np.array_equal(wa, wb) and np.array_equal(npa, npb)
        assert np.array_equal(sa[1], sb[1]) and np.array_equal(sa[2], sb[2])
        total += nea
    g.close()
    return total, eva


@pytest.mark.gpu
@pytest.mark.parametrize("K", [2, 3, 4])
def test_gpu_parity(small_mesh, cuda_device, K):
    mesh, _, settings = small_mesh
    total, ev = _gpu_pair(mesh, _with(settings, poly_order=K), 160, 5, 5e-4, n_skip_phi_0=2)
    assert total > 1000 and (ev["kind"] == 2).sum() > 30
    if K >= 3:
        _gpu_pair(mesh, _with(settings, poly_order=K), 64, 7, 3e-4, ncalls=1, use_group=False)


@pytest.mark.gpu
def test_gpu_parity_with_hamiltonian_time_and_phi(small_mesh_phi, cuda_device):
    mesh, _, settings = small_mesh_phi
    _gpu_pair(mesh, _with(settings, poly_order=2, i_time_tracing_option=2), 128, 9, 4e-4)


@pytest.mark.gpu
def test_gpu_event_buffer_overflow_and_refusals(small_mesh, cuda_device):
    from gorilla_b200 import Gorilla
    mesh, _, settings = small_mesh
    g = Gorilla(mesh, _with(settings, poly_order=2))
    n = 200
    x, vpar, vperp = workloads.particles_cyl(n, 3)
    st = workloads.fresh_state(n)
    J, cv, cp = _state(n)
    ev, nev = g.orbit_timestep_gorilla_events(x, vpar, vperp, 4e-4, *st, J, cv, cp, 64)
    assert nev > 64 and len(ev) == 64 and np.all(ev["kind"] > 0)       # surplus dropped, count still complete
    assert nev >= np.abs(cp).sum() * 0 + (cv.clip(2) - 2).sum()        # at least the reported banana tips
    g.close()
    for bad in (_with(settings, poly_order=1),):
        gb = Gorilla(mesh, bad)
        with pytest.raises(api.GorillaError):
            gb.orbit_timestep_gorilla_events(x, vpar, vperp, 1e-5, *workloads.fresh_state(n), *_state(n), 10)
        gb.close()


# ------------------------------------------------ boole_full_orbit (gorilla_plot_mod.f90:553-579): kind 3 events
@pytest.mark.parametrize("pusher,K,force_full", [(2, 2, False), (2, 2, True), (2, 4, False), (1, 4, False), (1, 4, True)])
def test_full_orbit_events(small_mesh, pusher, K, force_full):
    """The orbit point, p_phi and E_tot after every n_skip_full_orbit-th push (the reference's full_orbit_plot / p_phi /
    e_tot files), the push that ends the time step included: oracle <-> host compile of the device headers bit for bit,
    alone and together with the other event kinds, and the records against what the call itself returns."""
    mesh, _, settings = small_mesh
    st = _with(settings, ipusher=pusher, poly_order=K)
    om, hm = OracleMesh(mesh, st), HostMirror(mesh, st)
    n, t_step = 24, 2e-4
    for kw in (dict(poincare_phi_0=False, poincare_vpar_0=False, J_par=False, full_orbit=True, n_skip_full_orbit=1),
               dict(full_orbit=True, n_skip_full_orbit=3, n_skip_phi_0=2)):
        xa, va, wa = workloads.particles_cyl(n, 11)
        xb, vb, wb = xa.copy(), va.copy(), wa.copy()
        sa, sb = workloads.fresh_state(n), workloads.fresh_state(n)
        om.orbit_timestep_batch(xa, va, wa, 0.0, *sa)
        hm.orbit_timestep(xb, vb, wb, 0.0, *sb, 0)
        e0, p0, _ = om.invariants(xa, va, wa, sa[1])
        Ja, cva, cpa = _state(n)
        Jb, cvb, cpb = _state(n)
        eva, nea, npa = om.orbit_timestep_events(xa, va, wa, t_step, *sa, Ja, cva, cpa, 200000, **kw)
        evb, neb, npb = hm.orbit_timestep_events(xb, vb, wb, t_step, *sb, Jb, cvb, cpb, 200000, force_full=force_full, **kw)
        assert nea == neb and np.array_equal(_sorted(eva), _sorted(evb))
        assert np.array_equal(xa, xb) and np.array_equal(va, vb) and np.array_equal(npa, npb)
        assert np.array_equal(Ja, Jb) and np.array_equal(cva, cvb) and np.array_equal(cpa, cpb)
        fo = _sorted(eva[eva["kind"] == api.EVENT_FULL_ORBIT])
        nskip = kw["n_skip_full_orbit"]
        assert len(fo) == (npa // nskip).sum() and len(fo) > 500
        if kw.get("J_par", True):
            assert (eva["kind"] == api.EVENT_PHI_0).sum() > 0
        else:
            assert len(fo) == nea
        for i in range(n):
            f = fo[fo["particle"] == i]
            assert np.array_equal(f["counter"], nskip * np.arange(1, len(f) + 1)) and np.array_equal(f["push"], f["counter"] - 1)
            assert np.all(np.diff(f["t"]) > 0) and f["t"][-1] <= t_step
            if nskip == 1 and sa[1][i] > 0:        # the last record is the state the call returns
                assert np.array_equal(f["x"][-1], xa[i]) and f["t"][-1] == t_step
            if sa[1][i] > 0:                       # axisymmetric field, no potential: both invariants hold along the orbit
                assert np.abs(f["value"][:, 1] / e0[i] - 1).max() < (2e-3 if K == 2 and pusher == 2 else 1e-6)
                assert np.abs(f["value"][:, 0] - p0[i]).max() < 2e-2 * abs(p0).mean()


def test_full_orbit_refusals(small_mesh):
    mesh, _, settings = small_mesh
    om = OracleMesh(mesh, settings)
    n = 2
    x, v, w = workloads.particles_cyl(n, 1)
    with pytest.raises(AssertionError):
        om.orbit_timestep_events(x, v, w, 1e-5, *workloads.fresh_state(n), *_state(n), 10, full_orbit=True, n_skip_full_orbit=0)


@pytest.mark.gpu
@pytest.mark.parametrize("pusher,K", [(2, 2), (2, 4), (1, 4)])
def test_gpu_parity_full_orbit(small_mesh, cuda_device, pusher, K):
    mesh, _, settings = small_mesh
    st = _with(settings, ipusher=pusher, poly_order=K)
    # RK pusher: J_par goes through RKF45's pow() (not bit-identical on the GPU, see test_rk_pusher_gpu_parity): keep to the
    # kinds that come from the orbit itself
    extra = dict(poincare_vpar_0=False, J_par=False) if pusher == 1 else {}
    total, ev = _gpu_pair(mesh, st, 128, 13, 2e-4, ncalls=2, full_orbit=True, n_skip_full_orbit=2, n_skip_phi_0=2, **extra)
    assert (ev["kind"] == api.EVENT_FULL_ORBIT).sum() > 2000
    _gpu_pair(mesh, st, 64, 14, 1e-4, ncalls=1, poincare_phi_0=False, poincare_vpar_0=False, J_par=False, full_orbit=True)
