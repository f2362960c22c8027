"""Precomputed-coefficient modes of the polynomial pusher: i_precomp = 1 (orders 2-4) and 2 (order 2)
with the tetra_physics_poly4 records (SRC/tetra_physics_poly_precomp_mod.f90:21-45,160-476; analytic_coeff_with_precomp
SRC/pusher_tetra_poly.f90:1590-1725, analytic_integration_with_precomp :2530-2650, normal_velocity_func :2734-2738)."""
import dataclasses

import numpy as np
import pytest

import workloads
from gorilla_b200 import api, build_mesh
from host_mirror_binding import HostMirror
from oracle_binding import OracleMesh


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


CASES = [(2, 1), (3, 1), (4, 1), (2, 2)]   # (poly_order, i_precomp)


def _run_oracle(mesh, st, n, seed, t_step, cap, nsteps=2):
    om = OracleMesh(mesh, st)
    x, vpar, vperp = workloads.particles_cyl(n, seed)
    s = workloads.fresh_state(n)
    out = []
    for _ in range(nsteps):
        r = om.orbit_timestep_trace(x, vpar, vperp, t_step, *s, cap)
        out.append(r)
    return x, vpar, vperp, s, out


def test_poly4_records_follow_the_reference_structure(small_mesh, oracle_lib):
    """The matrices are sums over all orderings of k-fold products of alpha = cm_over_e*alpmat and beta = (-clight*betmat |
    curlA): numpy forms them from the binomial structure, amat_k(p) = (beta + p alpha)^k for any p."""
    mesh, _, settings = small_mesh
    om = OracleMesh(mesh, dataclasses.replace(settings, poly_order=2, i_precomp=1))
    p4, tp, cm = om.poly4, mesh.tetra_physics, mesh.scalars["cm_over_e"]
    assert p4.shape == (mesh.ntetr, 544) and np.isfinite(p4).all()
    t = np.arange(0, mesh.ntetr, 997)
    alp, bet = np.zeros((t.size, 4, 4)), np.zeros((t.size, 4, 4))
    alp[:, :3, :3] = cm * tp[t, 107:116].reshape(-1, 3, 3).transpose(0, 2, 1)
    alp[:, 3, 3] = cm * tp[t, 47]
    bet[:, :3, :3] = -2.9979e10 * tp[t, 116:125].reshape(-1, 3, 3).transpose(0, 2, 1)
    bet[:, :3, 3] = tp[t, 21:24]
    bet[:, 3, 3] = -2.9979e10 * tp[t, 48]
    mat = lambda k: p4[t, 16 * k:16 * k + 16].reshape(-1, 4, 4).transpose(0, 2, 1)   # noqa: E731  column-major -> [i, j]
    first = {1: 0, 2: 2, 3: 5, 4: 9}
    p = -3.7e-9     # any number: the records must reproduce (beta + p alpha)^k
    for k in (1, 2, 3, 4):
        want = np.linalg.matrix_power(bet + p * alp, k)
        got = sum(p ** q * mat(first[k] + q) for q in range(k + 1))
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
    # anorm_in_amat(:, n) = (anorm(:, n), 0) . amat
    an = np.concatenate([tp[t, 9:21].reshape(-1, 4, 3), np.zeros((t.size, 4, 1))], axis=2)
    for k in range(14):
        got = p4[t, 224 + 16 * k:224 + 16 * k + 16].reshape(-1, 4, 4)      # [n][j]
        assert np.abs(got - np.einsum("tni,tij->tnj", an, mat(k))).max() <= 1e-12 * max(np.abs(got).max(), 1e-300)


@pytest.mark.parametrize("K,ip", CASES)
def test_oracle_precomp_equals_plain_to_rounding(small_mesh, oracle_lib, K, ip):
    """sign_sqg = +1 and forward time: the precomputed coefficients are the same numbers as i_precomp = 0 up to rounding
    (i_precomp = 1), so the orbits agree to ~1e-12; i_precomp = 2 validates the exit velocity with the never-assigned module
    variable b (zero), which sends a few pushes into the fall-back ladder -- the orbits that stay on the main path agree."""
    mesh, _, settings = small_mesh
    assert mesh.scalars["sign_sqg"] == 1
    x0, v0, _, s0, r0 = _run_oracle(mesh, dataclasses.replace(settings, poly_order=K), 120, 5, 1e-5, 64)
    x1, v1, _, s1, r1 = _run_oracle(mesh, dataclasses.replace(settings, poly_order=K, i_precomp=ip), 120, 5, 1e-5, 64)
    agree = np.all(r0[1]["trace_tetr"] == r1[1]["trace_tetr"], axis=1) & np.all(r0[0]["trace_tetr"] == r1[0]["trace_tetr"], axis=1)
    assert agree.mean() > (0.999 if ip == 1 else 0.9)
    assert np.abs(x0[agree] - x1[agree]).max() < 1e-9 and np.abs(v0[agree] / v1[agree] - 1).max() < 1e-10


@pytest.mark.parametrize("K,ip", CASES)
def test_host_mirror_precomp_bit_exact(small_mesh, small_mesh_phi, oracle_lib, host_mirror_lib, K, ip):
    """Device headers compiled for the host == oracle, bit for bit: fast path and complete ladder, with and without Phi,
    backward time (sign_rhs = -1: the precomputed coefficients carry no sign, faithfully wrong), repeated calls."""
    for mesh, _, settings in (small_mesh, small_mesh_phi):
        st = dataclasses.replace(settings, poly_order=K, i_precomp=ip)
        for t_step, force_full in ((1e-5, False), (1e-5, True), (-6e-6, False)):
            om, hm = OracleMesh(mesh, st), HostMirror(mesh, st)
            xa, va, wa = workloads.particles_cyl(150, 11)
            xb, vb, wb = xa.copy(), va.copy(), wa.copy()
            sa, sb = workloads.fresh_state(150), workloads.fresh_state(150)
            for _ in range(2):
                ra = om.orbit_timestep_trace(xa, va, wa, t_step, *sa, 48)
                rb = hm.orbit_timestep(xb, vb, wb, t_step, *sb, trace_cap=48, force_full=force_full)
                assert same(ra["trace_tetr"], rb["trace_tetr"]) and same(ra["trace_face"], rb["trace_face"])
                assert same(xa, xb) and same(va, vb) and same(wa, wb) and same(ra["t_remain"], rb["t_remain"])
                assert same(sa[1], sb[1]) and same(sa[2], sb[2]) and same(ra["n_pushes"], rb["n_pushes"])
                # (a push whose first AND second attempt prolong the trajectory counts twice in the oracle, once per push on
                # the device; with sign_rhs = -1 the unsigned precomputed coefficients make that common)
                fa, fb = tuple(int(v) for v in ra["fallback"]), tuple(int(v) for v in rb["fallback"])
                assert fa[:2] == fb[:2] and fa[3] == fb[3] and fa[2] >= fb[2]


def test_unsupported_combinations_are_refused(small_mesh, product_lib):
    mesh, _, settings = small_mesh
    for kw in (dict(i_precomp=3), dict(i_precomp=1, poly_order=1), dict(i_precomp=2, poly_order=3),
               dict(i_precomp=1, poly_order=2, i_time_tracing_option=2), dict(i_precomp=1, poly_order=2, boole_adaptive_time_steps=True),
               dict(i_precomp=1, poly_order=2, boole_vpar_int=True)):
        with pytest.raises(api.GorillaError) as ei:
            api.Gorilla(mesh, dataclasses.replace(settings, **kw))
        assert ei.value.code == 2, kw


@pytest.mark.gpu
@pytest.mark.parametrize("K,ip", CASES)
def test_cuda_precomp_bit_exact(small_mesh, small_mesh_phi, cuda_device, K, ip):
    from gorilla_b200 import Gorilla
    for mesh, _, settings in (small_mesh, small_mesh_phi):
        st = dataclasses.replace(settings, poly_order=K, i_precomp=ip)
        for t_step, force_full in ((2e-5, False), (1e-5, True), (-6e-6, False)):
            om, g = OracleMesh(mesh, st), Gorilla(mesh, st)
            g._debug_force_full(force_full)
            n = 600
            xa, va, wa = workloads.particles_cyl(n, 13, rmax_frac=0.97)
            xb, vb, wb = xa.copy(), va.copy(), wa.copy()
            sa, sb = workloads.fresh_state(n), workloads.fresh_state(n)
            for _ in range(2):
                ra = om.orbit_timestep_trace(xa, va, wa, t_step, *sa, 96)
                tro, npu = np.zeros(n), np.zeros(n, np.int64)
                tt, tf = g.orbit_timestep_gorilla(xb, vb, wb, t_step, *sb, t_remain_out=tro, n_pushes=npu, trace_cap=96)
                c = g.counters()
                assert same(ra["trace_tetr"], tt) and same(ra["trace_face"], tf)
                assert same(xa, xb) and same(va, vb) and same(wa, wb) and same(ra["t_remain"], tro)
                assert same(sa[1], sb[1]) and same(sa[2], sb[2]) and same(ra["n_pushes"], npu)
                fa = tuple(int(v) for v in ra["fallback"])
                assert fa[:2] == c.n_fallback[:2] and fa[3] == c.n_fallback[3] and fa[2] >= c.n_fallback[2]
                assert c.n_pushes == int(npu.sum())
            g.close()


# ---- boole_newton_precalc: the RK pusher takes normal velocity / acceleration and the quadratic start guess from the
# tetra_physics_poly4 records (SRC/pusher_tetra_rk.f90:579-632, 2451-2527)
def _rk_pair(run_b, mesh, st, n, seed, t_step, cap, nsteps=2, **pk):
    om = OracleMesh(mesh, st)
    xa, va, wa = workloads.particles_cyl(n, seed, **pk)
    xb, vb, wb = xa.copy(), va.copy(), wa.copy()
    sa, sb = workloads.fresh_state(n), workloads.fresh_state(n)
    for _ in range(nsteps):
        ra = om.orbit_timestep_trace(xa, va, wa, t_step, *sa, cap)
        tt, tf, npu, tro, fb = run_b(xb, vb, wb, t_step, sb, cap)
        assert same(ra["trace_tetr"], tt) and same(ra["trace_face"], tf)
        assert same(xa, xb) and same(va, vb) and same(wa, wb) and same(ra["t_remain"], tro)
        assert same(sa[1], sb[1]) and same(sa[2], sb[2]) and same(ra["n_pushes"], npu)
        assert tuple(int(v) for v in ra["fallback"]) == tuple(int(v) for v in fb)
    return xa, va, sa[1]


def test_newton_precalc_host_mirror_bit_exact_and_consistent(small_mesh, small_mesh_phi, oracle_lib, host_mirror_lib):
    for mesh, _, settings in (small_mesh, small_mesh_phi):
        res = {}
        for pre in (False, True):
            st = dataclasses.replace(settings, ipusher=1, boole_newton_precalc=pre)
            hm = HostMirror(mesh, st)

            def run_b(x, v, w, t, s, cap):
                r = hm.orbit_timestep(x, v, w, t, *s, trace_cap=cap)
                return r["trace_tetr"], r["trace_face"], r["n_pushes"], r["t_remain"], r["fallback"]
            for t_step in (1e-5, -7e-6):
                res[(pre, t_step)] = _rk_pair(run_b, mesh, st, 200, 17, t_step, 64, rmax_frac=0.97)
        # the analytic normal velocity equals n . dz/dtau for sign_rhs = +1: same orbits to the Newton tolerance
        (x0, v0, i0), (x1, v1, i1) = res[(False, 1e-5)], res[(True, 1e-5)]
        ok = (i0 == i1) & (i0 > 0)
        assert ok.mean() > 0.97 and np.abs(x0[ok] - x1[ok]).max() < 1e-6


@pytest.mark.gpu
def test_newton_precalc_cuda_bit_exact(small_mesh, small_mesh_phi, cuda_device):
    from gorilla_b200 import Gorilla
    for mesh, _, settings in (small_mesh, small_mesh_phi):
        st = dataclasses.replace(settings, ipusher=1, boole_newton_precalc=True)
        for t_step, force_full in ((2e-5, False), (1e-5, True), (-7e-6, False)):
            g = Gorilla(mesh, st)
            g._debug_force_full(force_full)

            def run_b(x, v, w, t, s, cap):
                n = x.shape[0]
                tro, npu = np.zeros(n), np.zeros(n, np.int64)
                tt, tf = g.orbit_timestep_gorilla(x, v, w, t, *s, t_remain_out=tro, n_pushes=npu, trace_cap=cap)
                return tt, tf, npu, tro, g.counters().n_fallback
            _rk_pair(run_b, mesh, st, 800, 19, t_step, 96, rmax_frac=0.97)
            g.close()
