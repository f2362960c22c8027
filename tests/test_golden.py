"""Committed golden vectors (tests/golden/orbits_v1.npz and orbits_v2.npz -- the latter: Hamiltonian time tracing with the
optional quantities, adaptive sub-stepping, orbit events -- made by tests/golden/make_golden.py from the CPU oracle): eleven
small cases -- polynomial orders 1-4, backward time, no face guess, RK4, electrostatic potential, strong electric field --
with final phase-space state, visited-tetra trace and push counts.  CPU: the oracle still reproduces them bit for bit;
GPU: the CUDA path reproduces them WITHOUT the oracle in the loop.  orbits_v3.npz: the full-orbit output (event kind 3) with
the elapsed time of every event, polynomial and RK pusher -- checked on the CPU against the oracle and against the device
headers compiled for the host (the GPU side of these events is tests/test_orbit_events.py::test_gpu_parity_full_orbit)."""
import sys
from pathlib import Path

import numpy as np
import pytest

import workloads
from gorilla_b200 import build_mesh

sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))
from make_golden import CAP, CASES, CASES_V2, CASES_V3, EV_CAP, N, N_EV, run_case_v2, run_case_v3  # noqa: E402

GOLD = np.load(Path(__file__).resolve().parent / "golden" / "orbits_v1.npz")
GOLD2 = np.load(Path(__file__).resolve().parent / "golden" / "orbits_v2.npz")
GOLD3 = np.load(Path(__file__).resolve().parent / "golden" / "orbits_v3.npz")


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and bool(np.all((a == b) | (np.isnan(a.astype(float)) & np.isnan(b.astype(float)))))


def _setup(over, seed):
    grid, st = workloads.analytic_tokamak(10, 10, 10)
    st = type(st)(**{**st.__dict__, **over})
    mesh = build_mesh(grid, st)
    x, vpar, vperp = workloads.particles_cyl(N, seed)
    return mesh, st, x, vpar, vperp, workloads.fresh_state(N)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_reproduces_golden(product_lib, case):
    from oracle_binding import OracleMesh
    name, over, t_step, seed = case
    mesh, st, x, vpar, vperp, (binit, ind, ifc) = _setup(over, seed)
    r = OracleMesh(mesh, st).orbit_timestep_trace(x, vpar, vperp, t_step, binit, ind, ifc, CAP)
    g = lambda k: GOLD[f"{name}/{k}"]  # noqa: E731
    assert same(r["trace_tetr"], g("trace_tetr")) and same(r["trace_face"], g("trace_face"))
    assert same(r["n_pushes"], g("n_pushes")) and same(r["t_remain"], g("t_remain"))
    assert same(x, g("x")) and same(vpar, g("vpar")) and same(vperp, g("vperp"))
    assert same(ind, g("ind_tetr")) and same(ifc, g("iface"))
    assert int(g("n_pushes").sum()) > 500


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_cuda_reproduces_golden(cuda_device, product_lib, case):
    from gorilla_b200 import Gorilla
    name, over, t_step, seed = case
    mesh, st, x, vpar, vperp, (binit, ind, ifc) = _setup(over, seed)
    gk = Gorilla(mesh, st)
    tro, npu = np.zeros(N), np.zeros(N, np.int64)
    tt, tf = gk.orbit_timestep_gorilla(x, vpar, vperp, t_step, binit, ind, ifc, t_remain_out=tro, n_pushes=npu, trace_cap=CAP)
    gk.close()
    g = lambda k: GOLD[f"{name}/{k}"]  # noqa: E731
    assert same(tt, g("trace_tetr")) and same(tf, g("trace_face")), "visited tetra sequence differs from the golden vector"
    assert same(npu, g("n_pushes")) and same(tro, g("t_remain"))
    assert same(x, g("x")) and same(vpar, g("vpar")) and same(vperp, g("vperp"))
    assert same(ind, g("ind_tetr")) and same(ifc, g("iface"))


# ---- orbits_v2.npz: Hamiltonian time + optional quantities, adaptive sub-stepping, orbit events -------------------
@pytest.mark.parametrize("case", CASES_V2, ids=[c[0] for c in CASES_V2])
def test_oracle_reproduces_golden_v2(product_lib, case):
    name, over, t_step, seed, kind = case
    r = run_case_v2(over, t_step, seed, kind)
    keys = [k.split("/", 1)[1] for k in GOLD2.files if k.startswith(name + "/")]
    assert sorted(keys) == sorted(r.keys())
    for k in keys:
        assert same(r[k], GOLD2[f"{name}/{k}"]), k
    assert int(GOLD2[f"{name}/n_pushes"].sum()) > 500
    if kind == "events":
        assert (GOLD2[f"{name}/ev_kind"] == 2).sum() > 5 and (GOLD2[f"{name}/ev_kind"] == 1).sum() > 20


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES_V2, ids=[c[0] for c in CASES_V2])
def test_cuda_reproduces_golden_v2(cuda_device, product_lib, case):
    from gorilla_b200 import Gorilla
    name, over, t_step, seed, kind = case
    n = N_EV if kind == "events" else N
    grid, st = workloads.analytic_tokamak(10, 10, 10)
    st = type(st)(**{**st.__dict__, **over})
    mesh = build_mesh(grid, st)
    x, vpar, vperp = workloads.particles_cyl(n, seed)
    binit, ind, ifc = workloads.fresh_state(n)
    gk = Gorilla(mesh, st)
    g = lambda k: GOLD2[f"{name}/{k}"]  # noqa: E731
    npu = np.zeros(n, np.int64)
    if kind == "events":
        J, cv, cp = np.zeros(n), np.zeros(n, np.int32), np.zeros(n, np.int32)
        ev, nev = gk.orbit_timestep_gorilla_events(x, vpar, vperp, t_step, binit, ind, ifc, J, cv, cp, EV_CAP, n_skip_phi_0=2,
                                                   n_pushes=npu)
        assert nev == len(g("ev_kind"))
        for k in ("particle", "kind", "counter", "push", "x", "value"):
            assert same(ev[k], g("ev_" + k)), k
        assert same(J, g("par_adiab_inv")) and same(cv, g("counter_vpar_0")) and same(cp, g("counter_phi_0"))
    else:
        tro = np.zeros(n)
        oq = np.zeros((n, 4)) if kind == "optq" else None
        tt, tf = gk.orbit_timestep_gorilla(x, vpar, vperp, t_step, binit, ind, ifc, t_remain_out=tro, n_pushes=npu,
                                           trace_cap=CAP, optional_quantities=oq)
        assert same(tt, g("trace_tetr")) and same(tf, g("trace_face")), "visited tetra sequence differs from the golden vector"
        assert same(tro, g("t_remain"))
        if oq is not None:
            assert same(oq, g("optional_quantities"))
    gk.close()
    assert same(npu, g("n_pushes"))
    assert same(x, g("x")) and same(vpar, g("vpar")) and same(vperp, g("vperp"))
    assert same(ind, g("ind_tetr")) and same(ifc, g("iface"))


# ---- orbits_v3.npz: full-orbit output (event kind 3) and the elapsed time of every event ----------------------------------
@pytest.mark.parametrize("case", CASES_V3, ids=[c[0] for c in CASES_V3])
def test_oracle_and_device_algorithm_reproduce_golden_v3(product_lib, case):
    """The oracle, and the device headers compiled for the host (tests/host_mirror), against the committed vectors."""
    from host_mirror_binding import HostMirror
    name, over, t_step, seed, switches = case
    r = run_case_v3(over, t_step, seed, switches)
    keys = [k.split("/", 1)[1] for k in GOLD3.files if k.startswith(name + "/")]
    assert sorted(keys) == sorted(r.keys())
    for k in keys:
        assert same(r[k], GOLD3[f"{name}/{k}"]), k
    assert (GOLD3[f"{name}/ev_kind"] == 3).sum() > 1500
    grid, st = workloads.analytic_tokamak(10, 10, 10)
    st = type(st)(**{**st.__dict__, **over})
    hm = HostMirror(build_mesh(grid, st), st)
    x, vpar, vperp = workloads.particles_cyl(N_EV, seed)
    binit, ind, ifc = workloads.fresh_state(N_EV)
    hm.orbit_timestep(x, vpar, vperp, 0.0, binit, ind, ifc, 0)      # the mirror's event call wants localised particles
    J, cv, cp = np.zeros(N_EV), np.zeros(N_EV, np.int32), np.zeros(N_EV, np.int32)
    ev, nev, npush = hm.orbit_timestep_events(x, vpar, vperp, t_step, binit, ind, ifc, J, cv, cp, EV_CAP, **switches)
    ev = ev[np.lexsort((ev["kind"], ev["push"], ev["particle"]))]
    g = lambda k: GOLD3[f"{name}/{k}"]  # noqa: E731
    assert nev == len(g("ev_kind")) and same(npush, g("n_pushes"))
    for k in ("particle", "kind", "counter", "push", "x", "value", "t"):
        assert same(ev[k], g("ev_" + k)), k
    assert same(x, g("x")) and same(vpar, g("vpar")) and same(ind, g("ind_tetr")) and same(ifc, g("iface"))
    assert same(J, g("par_adiab_inv")) and same(cv, g("counter_vpar_0")) and same(cp, g("counter_phi_0"))
