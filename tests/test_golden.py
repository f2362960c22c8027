"""Committed golden vectors (tests/golden/orbits_v1.npz, made by tests/golden/make_golden.py from the CPU oracle): eleven
small cases -- polynomial orders 1-4, backward time, no face guess, RK4, electrostatic potential, strong electric field --
with final phase-space state, visited-tetra trace and push counts.  CPU: the oracle still reproduces them bit for bit;
GPU: the CUDA path reproduces them WITHOUT the oracle in the loop."""
import sys
from pathlib import Path

import numpy as np
import pytest

import workloads
from gorilla_b200 import build_mesh

sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))
from make_golden import CAP, CASES, N  # noqa: E402

GOLD = np.load(Path(__file__).resolve().parent / "golden" / "orbits_v1.npz")


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
