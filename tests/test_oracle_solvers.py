"""Known-answer tests that pin the oracle's solver chain (SRC/pusher_tetra_poly.f90:1767-2021,
SRC/contrib/Polynomial234RootSolvers.f90, SRC/contrib/cmplx_roots_sg.f90).  The reference ships no
vectors for this path, so the answers are constructed: polynomials built from chosen roots."""
import ctypes as C

import numpy as np
import pytest


def _roots(L, deg, coeffs):
    n = C.c_int()
    root = (C.c_double * (2 * deg))()
    if deg == 2:
        L.gor_quadratic_roots(coeffs[1], coeffs[2], C.byref(n), root)
    elif deg == 3:
        L.gor_cubic_roots(coeffs[1], coeffs[2], coeffs[3], C.byref(n), root)
    else:
        L.gor_quartic_roots(coeffs[1], coeffs[2], coeffs[3], coeffs[4], C.byref(n), root)
    r = np.array(root[:]).reshape(2, deg).T  # Fortran root(n,2)
    return n.value, r


@pytest.mark.parametrize("roots", [
    [1.0, 2.0], [-3.0, 0.25], [1.0, 2.0, 3.0], [0.5, -0.5, 4.0], [1.0, 2.0, 3.0, 4.0],
    [1e-3, 1e3, 5.0, 7.0], [-1.0, -2.0, 0.1, 50.0],
])
def test_real_roots_sorted_descending(oracle_lib, roots):
    c = np.poly(roots)
    n, r = _roots(oracle_lib, len(roots), c)
    assert n == len(roots)
    assert np.all(r[:, 1] == 0.0)  # real roots are emitted with an exactly-zero imaginary part
    np.testing.assert_allclose(r[:, 0], sorted(roots, reverse=True), rtol=1e-12)


def test_complex_pairs_come_after_real_roots(oracle_lib):
    c = np.poly([1.0, -2.0, 0.5 + 1.0j, 0.5 - 1.0j]).real
    n, r = _roots(oracle_lib, 4, c)
    assert n == 2
    np.testing.assert_allclose(r[:2, 0], [1.0, -2.0], rtol=1e-13)
    assert np.all(r[:2, 1] == 0.0)
    np.testing.assert_allclose(sorted(r[2:, 1]), [-1.0, 1.0], rtol=1e-13)
    np.testing.assert_allclose(r[2:, 0], [0.5, 0.5], rtol=1e-13)


def test_random_polynomials_residual(oracle_lib):
    rng = np.random.default_rng(7)
    for deg in (2, 3, 4):
        for _ in range(300):
            q = rng.normal(size=deg) * 10.0 ** rng.integers(-2, 3, size=deg)
            poly = np.zeros(2 * (deg + 1))
            poly[0:2 * deg:2] = q
            poly[2 * deg] = 1.0
            out = np.zeros(2 * deg)
            oracle_lib.gor_cmplx_roots_gen(deg, poly.ctypes.data_as(C.POINTER(C.c_double)),
                                           out.ctypes.data_as(C.POINTER(C.c_double)))
            z = out[0::2] + 1j * out[1::2]
            coeff = np.concatenate([[1.0], q[::-1]])
            ref = np.roots(coeff)
            # every computed root is a root: compare multisets
            for zi in z:
                assert np.min(np.abs(ref - zi)) <= 1e-8 * max(1.0, abs(zi))


def test_exit_time_solvers_pick_smallest_positive_root(oracle_lib):
    L = oracle_lib
    # f(tau) = a/24 tau^4 + b/6 tau^3 + c/2 tau^2 + d tau + e with roots r
    r = [0.3, -1.0, 2.0, 5.0]
    c = np.poly(r)
    for s in range(7):
        tau = L.gor_quartic_solver(s, 24 * c[0], 6 * c[1], 2 * c[2], c[3], c[4])
        assert tau == pytest.approx(0.3, rel=1e-13)
    r3 = [-0.7, 1.5, 0.2]
    c3 = np.poly(r3)
    assert L.gor_cubic_solver(6 * c3[0], 2 * c3[1], c3[2], c3[3]) == pytest.approx(0.2, rel=1e-13)
    r2 = [4.0, 0.5]
    c2 = np.poly(r2)
    assert L.gor_quadratic_solver1(2 * c2[0], c2[1], c2[2]) == pytest.approx(0.5, rel=1e-14)
    assert L.gor_quadratic_solver2(2 * c2[0], c2[1], c2[2]) == pytest.approx(0.5, rel=1e-13)
    # no positive real root -> huge(0.d0)
    huge = np.finfo(np.float64).max
    assert L.gor_quadratic_solver1(2.0, 3.0, 2.0) == huge          # complex pair
    assert L.gor_cubic_solver(6.0, 2 * 6.0, 11.0, 6.0) == huge     # roots -1,-2,-3
    assert L.gor_quadratic_solver1(0.0, -2.0, 1.0) == pytest.approx(0.5)  # degenerate a = 0 -> linear


def test_quadratic_solver1_sign_cases(oracle_lib):
    """Every branch of the sign-case tree (:1827-1891) against the closed form."""
    rng = np.random.default_rng(11)
    huge = np.finfo(np.float64).max
    for _ in range(4000):
        a, b, c = rng.normal(size=3)
        if rng.random() < 0.15:
            a = 0.0
        if rng.random() < 0.1:
            c = 0.0
        got = oracle_lib.gor_quadratic_solver1(a, b, c)
        # smallest strictly positive root of a/2 t^2 + b t + c (c = 0: the non-trivial root)
        if a != 0.0:
            disc = b * b - 2 * a * c
            cand = [] if disc < 0 else [(-b + np.sqrt(disc)) / a, (-b - np.sqrt(disc)) / a]
        else:
            cand = [-c / b] if b != 0.0 else []
        cand = [t for t in cand if t > 0]
        if c == 0.0 and a != 0.0:
            cand = [t for t in [-2 * b / a] if t > 0]
        if not cand:
            assert got == huge or got <= 0.0 or not np.isfinite(got)
        else:
            assert got == pytest.approx(min(cand), rel=1e-9)
