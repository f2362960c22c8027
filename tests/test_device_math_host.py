"""Pins the product's explicit FP64 complex arithmetic (gorilla_b200/csrc/gb_math.cuh, gb_roots.cuh), compiled
for the host, bit-for-bit against the real thing: gcc -fcx-fortran-rules (the lowering gfortran uses) and glibc
cabs/csqrt/cexp, probed through the oracle library, and against the oracle's solver chain."""
import ctypes as C

import numpy as np

DP = C.POINTER(C.c_double)


def _bits(a):
    return np.asarray(a, dtype=np.float64).view(np.int64)


def _rand(rng, n, lo=-150, hi=150):
    s = rng.choice([-1.0, 1.0], size=n)
    return s * rng.random(n) * 10.0 ** rng.uniform(lo, hi, size=n)


def test_hypot_csqrt_cdiv_cmul_rmul_bit_exact(oracle_lib, host_mirror_lib):
    O, H = oracle_lib, host_mirror_lib
    O.gor_probe_cabs.restype = C.c_double
    O.gor_probe_cabs.argtypes = [C.c_double] * 2
    for f in (O.gor_probe_csqrt,):
        f.argtypes = [C.c_double, C.c_double, DP]
    O.gor_probe_cdiv.argtypes = [C.c_double] * 4 + [DP]
    O.gor_probe_cmul.argtypes = [C.c_double] * 4 + [DP]
    O.gor_probe_rmul.argtypes = [C.c_double] * 3 + [DP]
    rng = np.random.default_rng(2024)
    n = 60000
    for lo, hi in ((-3, 3), (-150, 150)):
        a, b, c, d = (_rand(rng, n, lo, hi) for _ in range(4))
        # special operands the solver produces all the time: purely real / purely imaginary / signed zeros
        b[::7] = 0.0
        b[3::7] = -0.0
        a[5::11] = 0.0
        a[6::11] = -0.0
        o1, o2 = (C.c_double * 2)(), (C.c_double * 2)()
        for i in range(n):
            assert _bits(O.gor_probe_cabs(a[i], b[i])) == _bits(H.hm_hypot(a[i], b[i]))
            O.gor_probe_csqrt(a[i], b[i], o1); H.hm_csqrt(a[i], b[i], o2)
            assert list(_bits(o1[:])) == list(_bits(o2[:])), ("csqrt", a[i], b[i])
            O.gor_probe_cmul(a[i], b[i], c[i], d[i], o1); H.hm_cmul(a[i], b[i], c[i], d[i], o2)
            assert list(_bits(o1[:])) == list(_bits(o2[:])), ("cmul", a[i], b[i], c[i], d[i])
            O.gor_probe_rmul(c[i], a[i], b[i], o1); H.hm_rmul(c[i], a[i], b[i], o2)
            assert list(_bits(o1[:])) == list(_bits(o2[:])), ("rmul", c[i], a[i], b[i])
            if c[i] != 0.0 or d[i] != 0.0:
                O.gor_probe_cdiv(a[i], b[i], c[i], d[i], o1); H.hm_cdiv(a[i], b[i], c[i], d[i], o2)
                assert list(_bits(o1[:])) == list(_bits(o2[:])), ("cdiv", a[i], b[i], c[i], d[i])


def test_frac_jump_phase_table_matches_glibc_cexp(oracle_lib, host_mirror_lib):
    o1, o2 = (C.c_double * 2)(), (C.c_double * 2)()
    for k in range(10):
        oracle_lib.gor_frac_jump_phase(k, o1)
        host_mirror_lib.hm_frac_jump_phase(k, o2)
        assert list(_bits(o1[:])) == list(_bits(o2[:]))


def test_cmplx_roots_gen_bit_exact_vs_oracle(oracle_lib, host_mirror_lib):
    """Laguerre -> SG -> Newton, deflation, Viete, polish: every root identical to the last bit,
    for real-coefficient monic polynomials (the only kind the pusher produces) of degree 2..4."""
    rng = np.random.default_rng(99)
    for deg in (2, 3, 4):
        for trial in range(4000):
            if trial % 3 == 0:  # well separated real roots
                q = np.poly(rng.normal(size=deg) * 10.0 ** rng.integers(-3, 4, size=deg))[1:][::-1]
            elif trial % 3 == 1:  # generic
                q = rng.normal(size=deg) * 10.0 ** rng.integers(-6, 7, size=deg)
            else:  # nearly multiple roots
                r0 = rng.normal()
                q = np.poly(r0 + rng.normal(size=deg) * 1e-7)[1:][::-1]
            poly = np.zeros(2 * (deg + 1))
            poly[0:2 * deg:2] = q
            poly[2 * deg] = 1.0
            a, b = np.zeros(2 * deg), np.zeros(2 * deg)
            oracle_lib.gor_cmplx_roots_gen(deg, poly.ctypes.data_as(DP), a.ctypes.data_as(DP))
            host_mirror_lib.hm_cmplx_roots_gen(deg, poly.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p))
            assert list(_bits(a)) == list(_bits(b)), (deg, q)


def test_exit_time_solvers_bit_exact_vs_oracle(oracle_lib, host_mirror_lib):
    O, H = oracle_lib, host_mirror_lib
    rng = np.random.default_rng(5)
    for _ in range(3000):
        a, b, c, d, e = rng.normal(size=5) * 10.0 ** rng.integers(-4, 5, size=5)
        assert _bits(O.gor_quadratic_solver1(a, b, c)) == _bits(H.hm_quadratic_solver1(a, b, c))
        assert _bits(O.gor_quadratic_solver2(a, b, c)) == _bits(H.hm_quadratic_solver2(a, b, c))
        assert _bits(O.gor_cubic_solver(a, b, c, d)) == _bits(H.hm_cubic_solver(a, b, c, d))
        for s in (0, 1, 3, 4, 5):  # 2 and 6 go through libm pow(): compared with a tolerance below
            assert _bits(O.gor_quartic_solver(s, a, b, c, d, e)) == _bits(H.hm_quartic_solver(s, a, b, c, d, e))
        for s in (2, 6):
            x, y = O.gor_quartic_solver(s, a, b, c, d, e), H.hm_quartic_solver(s, a, b, c, d, e)
            assert x == y  # same libm on the host; on the device pow() may differ in the last bit
