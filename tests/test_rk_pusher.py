"""ipusher = 1 (RK4 pusher, SRC/pusher_tetra_rk.f90): oracle physics checks, device-algorithm parity on the host
(tests/host_mirror) including the Newton / last-line-of-defence / bisection ladders, and CUDA parity (gpu)."""
import numpy as np
import pytest

import workloads
from host_mirror_binding import HostMirror
from oracle_binding import OracleMesh


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def _settings(base, **kw):
    return type(base)(**{**base.__dict__, "ipusher": 1, **kw})


def test_rk4_agrees_with_order4_polynomial_pusher(small_mesh):
    mesh, _, settings = small_mesh
    out = {}
    for name, st in (("poly", type(settings)(**{**settings.__dict__, "poly_order": 4})), ("rk", _settings(settings))):
        om = OracleMesh(mesh, st)
        x, vpar, vperp = workloads.particles_cyl(200, 2)
        s = workloads.fresh_state(200)
        om.orbit_timestep_batch(x, vpar, vperp, 0.0, *s)
        e0, p0, mu0 = om.invariants(x, vpar, vperp, s[1])
        for _ in range(3):
            om.orbit_timestep_batch(x, vpar, vperp, 2e-5, *s, nthreads=4)
        e1, p1, mu1 = om.invariants(x, vpar, vperp, s[1])
        assert np.abs(e1 / e0 - 1).max() < 1e-10 and np.abs(mu1 / mu0 - 1).max() < 1e-13
        out[name] = (x, vpar, s[1])
    assert (out["poly"][2] == out["rk"][2]).mean() > 0.99
    ok = out["poly"][2] == out["rk"][2]
    assert np.median(np.abs(out["poly"][0] - out["rk"][0])[ok]) < 1e-8


@pytest.mark.parametrize("force_full", [False, True])
def test_host_mirror_parity_regular(small_mesh, force_full):
    mesh, _, settings = small_mesh
    st = _settings(settings)
    om, hm = OracleMesh(mesh, st), HostMirror(mesh, st)
    n = 300
    xa, va, wa = workloads.particles_cyl(n, 3)
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    sa, sb = workloads.fresh_state(n), workloads.fresh_state(n)
    for t_step in (5e-5, -3e-5):
        ra = om.orbit_timestep_trace(xa, va, wa, t_step, *sa, 256)
        rb = hm.orbit_timestep(xb, vb, wb, t_step, *sb, 256, force_full=force_full)
        assert same(ra["trace_tetr"], rb["trace_tetr"]) and same(ra["trace_face"], rb["trace_face"])
        assert same(xa, xb) and same(va, vb) and same(wa, wb) and same(sa[1], sb[1]) and same(sa[2], sb[2])
        assert same(ra["t_remain"], rb["t_remain"]) and same(ra["n_pushes"], rb["n_pushes"])
        assert same(ra["fallback"], rb["fallback"])
    assert ra["n_pushes"].sum() > 10000


STRESS = [  # (n1, n2, n3, energy_eV, t_step, ispecies): coarse cells / fast particles push the ladders
    (4, 4, 4, 3e3, 2e-4, 2), (6, 5, 6, 3e5, 5e-5, 2), (6, 6, 6, 3e4, 2e-5, 1), (5, 7, 5, 3e4, 3e-4, 2),
]


@pytest.mark.parametrize("case", STRESS)
def test_host_mirror_parity_fallback_ladders(product_lib, case):
    """Newton failures, three-plane switches, last line of defence and bisection all occur in these regimes."""
    from gorilla_b200 import build_mesh
    n1, n2, n3, energy, t_step, sp = case
    grid, settings = workloads.analytic_tokamak(n1, n2, n3)
    st = _settings(settings, ispecies=sp)
    mesh = build_mesh(grid, st)
    om, hm = OracleMesh(mesh, st), HostMirror(mesh, st)
    n = 300
    mass = 2 * workloads.AMP if sp == 2 else 9.1094e-28
    xa, va, wa = workloads.particles_cyl(n, 11, energy_ev=energy, mass=mass, rmin_frac=0.05, rmax_frac=0.9)
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    sa, sb = workloads.fresh_state(n), workloads.fresh_state(n)
    tot = np.zeros(4, np.int64)
    for _ in range(2):
        ra = om.orbit_timestep_trace(xa, va, wa, t_step, *sa, 64)
        rb = hm.orbit_timestep(xb, vb, wb, t_step, *sb, 64)
        assert same(ra["trace_tetr"], rb["trace_tetr"]) and same(xa, xb) and same(va, vb) and same(wa, wb)
        assert same(sa[1], sb[1]) and same(sa[2], sb[2]) and same(ra["t_remain"], rb["t_remain"])
        assert same(ra["fallback"], rb["fallback"])
        tot += ra["fallback"]
    assert tot[0] + tot[2] > 0   # the ladders were actually entered


def test_vmec_flux_coordinates(product_lib):
    from pathlib import Path
    from gorilla_b200 import build_mesh
    nc = Path(__file__).resolve().parent.parent / "data" / "equilibria" / "netcdf_file_for_test.nc"
    grid, settings = workloads.vmec_qi(nc, 12, 8, 10)
    st = _settings(settings)
    mesh = build_mesh(grid, st)
    om, hm = OracleMesh(mesh, st), HostMirror(mesh, st)
    n = 150
    xa, va, wa = workloads.particles_vmec_alpha(n, 3)
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    sa, sb = workloads.fresh_state(n), workloads.fresh_state(n)
    ra = om.orbit_timestep_trace(xa, va, wa, 3e-5, *sa, 256)
    rb = hm.orbit_timestep(xb, vb, wb, 3e-5, *sb, 256)
    assert same(ra["trace_tetr"], rb["trace_tetr"]) and same(xa, xb) and same(va, vb) and same(wa, wb)
    assert ra["n_pushes"].sum() > 5000


@pytest.mark.gpu
@pytest.mark.parametrize("force_full", [False, True])
def test_gpu_parity_regular(small_mesh, cuda_device, force_full):
    from gorilla_b200 import Gorilla
    mesh, _, settings = small_mesh
    st = _settings(settings)
    om, g = OracleMesh(mesh, st), Gorilla(mesh, st)
    g._debug_force_full(force_full)
    n = 1000
    xa, va, wa = workloads.particles_cyl(n, 3)
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    sa, sb = workloads.fresh_state(n), workloads.fresh_state(n)
    for t_step in (5e-5, -3e-5):
        ra = om.orbit_timestep_trace(xa, va, wa, t_step, *sa, 128)
        tro, npu = np.zeros(n), np.zeros(n, np.int64)
        tt, tf = g.orbit_timestep_gorilla(xb, vb, wb, t_step, *sb, t_remain_out=tro, n_pushes=npu, trace_cap=128)
        c = g.counters()
        assert same(ra["trace_tetr"], tt) and same(ra["trace_face"], tf)
        assert same(xa, xb) and same(va, vb) and same(wa, wb) and same(sa[1], sb[1]) and same(sa[2], sb[2])
        assert same(ra["t_remain"], tro) and same(ra["n_pushes"], npu)
        assert tuple(int(v) for v in ra["fallback"]) == c.n_fallback
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", STRESS)
def test_gpu_parity_fallback_ladders(cuda_device, product_lib, case):
    from gorilla_b200 import Gorilla, build_mesh
    n1, n2, n3, energy, t_step, sp = case
    grid, settings = workloads.analytic_tokamak(n1, n2, n3)
    st = _settings(settings, ispecies=sp)
    mesh = build_mesh(grid, st)
    om, g = OracleMesh(mesh, st), Gorilla(mesh, st)
    n = 600
    mass = 2 * workloads.AMP if sp == 2 else 9.1094e-28
    xa, va, wa = workloads.particles_cyl(n, 11, energy_ev=energy, mass=mass, rmin_frac=0.05, rmax_frac=0.9)
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    sa, sb = workloads.fresh_state(n), workloads.fresh_state(n)
    for _ in range(2):
        ra = om.orbit_timestep_trace(xa, va, wa, t_step, *sa, 64)
        tro = np.zeros(n)
        tt, tf = g.orbit_timestep_gorilla(xb, vb, wb, t_step, *sb, t_remain_out=tro, trace_cap=64)
        assert same(ra["trace_tetr"], tt) and same(xa, xb) and same(va, vb) and same(wa, wb)
        assert same(sa[1], sb[1]) and same(sa[2], sb[2]) and same(ra["t_remain"], tro)
        assert tuple(int(v) for v in ra["fallback"]) == g.counters().n_fallback
    g.close()
