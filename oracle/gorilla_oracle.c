/*
 * gorilla_oracle.c -- CPU ORACLE (test infrastructure, see gorilla_oracle.h).
 *
 * Build: gcc -std=gnu11 -O2 -ffp-contract=off -fcx-fortran-rules -fopenmp -fPIC -shared
 *   -ffp-contract=off  : the reference ISA (x86-64 baseline) has no FMA (CMakeLists.txt:24-25)
 *   -fcx-fortran-rules : gfortran's complex * and / lowering (Smith-type division, no NaN recovery)
 * Mixed real*complex follows Fortran semantics: the real operand is promoted to (r, 0.0) and a
 * full complex product is formed (this matters for the sign of zero imaginary parts that
 * csqrt() later branches on).
 */
#define _GNU_SOURCE
#include "gorilla_oracle.h"
#include <complex.h>
#include <float.h>
#include <math.h>
#include <stdbool.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* SRC/constants_mod.f90:3-8 */
static const double PI = 3.141592653589793238462643383;
static const double CLIGHT = 2.9979e10;
static const double EPS = 1.e-10;
#define HUGE_D DBL_MAX

typedef double _Complex cplx;
static inline cplx rc(double r) { return CMPLX(r, 0.0); }               /* real -> complex promotion */
static inline cplx rmul(double r, cplx z) { return rc(r) * z; }         /* Fortran real*complex */
static inline bool ceq(cplx a, cplx b) { return creal(a) == creal(b) && cimag(a) == cimag(b); }
static inline double abs2(cplx p) { return creal(conj(p) * p); }        /* real(conjg(p)*p) */

/* ------------------------------------------------------------------------------------------------
 * Skowron & Gould solver -- SRC/contrib/cmplx_roots_sg.f90
 * ---------------------------------------------------------------------------------------------- */
static const double FRAC_JUMPS[10] = {0.64109297, 0.91577881, 0.25921289, 0.50487203, 0.08177045,
                                      0.13653241, 0.306162,   0.37794326, 0.04618805, 0.75132137};
static const double SG_PI = 3.141592653589793;
static const double FRAC_ERR = 2.0e-15;

static inline cplx frac_jump_phase(int k) /* exp(cmplx(0, FRAC_JUMPS(k+1)*2*pi)) */
{
  return cexp(CMPLX(0.0, FRAC_JUMPS[k] * 2 * SG_PI));
}
void gor_frac_jump_phase(int k, double out[2])
{
  cplx e = frac_jump_phase(k);
  out[0] = creal(e);
  out[1] = cimag(e);
}

/* cmplx_roots_sg.f90:558-731 */
static void cmplx_laguerre(const cplx *poly, int degree, cplx *root, int *iter, bool *success)
{
  const int MAX_ITERS = 200, FRAC_JUMP_EVERY = 10, FRAC_JUMP_LEN = 10;
  cplx p, dp, d2p_half, denom, denom_sqrt, dx, newroot, fac_netwon = 0, fac_extra, F_half;
  const cplx c_one = CMPLX(1.0, 0.0), zero = CMPLX(0.0, 0.0);
  double ek, absroot, abs2p, faq, stopping_crit2;
  *iter = 0;
  *success = true;
  bool good_to_go = false;
  double one_nth = 1.0 / degree;
  double n_1_nth = (degree - 1.0) * one_nth;
  double two_n_div_n_1 = 2.0 / n_1_nth;
  cplx c_one_nth = CMPLX(one_nth, 0.0);

  for (int i = 1; i <= MAX_ITERS; i++) {
    ek = cabs(poly[degree]);
    absroot = cabs(*root);
    p = poly[degree];
    dp = zero;
    d2p_half = zero;
    for (int k = degree; k >= 1; k--) {
      d2p_half = dp + d2p_half * (*root);
      dp = p + dp * (*root);
      p = poly[k - 1] + p * (*root);
      ek = absroot * ek + cabs(p);
    }
    *iter = *iter + 1;
    abs2p = abs2(p);
    if (abs2p == 0.0) return;
    stopping_crit2 = (FRAC_ERR * ek) * (FRAC_ERR * ek);
    if (abs2p < stopping_crit2) {
      if (abs2p < 0.01 * stopping_crit2) return;
      good_to_go = true;
    } else {
      good_to_go = false;
    }
    faq = 1.0;
    denom = zero;
    if (!ceq(dp, zero)) {
      fac_netwon = p / dp;
      fac_extra = d2p_half / dp;
      F_half = fac_netwon * fac_extra;
      denom_sqrt = csqrt(c_one - rmul(two_n_div_n_1, F_half));
      if (creal(denom_sqrt) >= 0.0)
        denom = c_one_nth + rmul(n_1_nth, denom_sqrt);
      else
        denom = c_one_nth - rmul(n_1_nth, denom_sqrt);
    }
    if (ceq(denom, zero)) {
      dx = rmul(absroot + 1.0, frac_jump_phase(i % FRAC_JUMP_LEN));
    } else {
      dx = fac_netwon / denom;
    }
    newroot = *root - dx;
    if (ceq(newroot, *root)) return;
    if (good_to_go) {
      *root = newroot;
      return;
    }
    if (i % FRAC_JUMP_EVERY == 0) {
      faq = FRAC_JUMPS[(i / FRAC_JUMP_EVERY - 1) % FRAC_JUMP_LEN];
      newroot = *root - rmul(faq, dx);
    }
    *root = newroot;
  }
  *success = false;
}

/* cmplx_roots_sg.f90:906-1305 */
static void cmplx_laguerre2newton(const cplx *poly, int degree, cplx *root, int *iter, bool *success,
                                  int starting_mode)
{
  const int MAX_ITERS = 50, FRAC_JUMP_EVERY = 10, FRAC_JUMP_LEN = 10;
  cplx p, dp, d2p_half, denom, denom_sqrt, dx, newroot, fac_netwon = 0, fac_extra, F_half;
  const cplx c_one = CMPLX(1.0, 0.0), zero = CMPLX(0.0, 0.0);
  cplx c_one_nth = zero;
  double ek, absroot, abs2p, abs2_F_half, faq, stopping_crit2 = 0.0;
  double one_nth = 0, n_1_nth = 0, two_n_div_n_1 = 0;
  int i, j = 1, mode = starting_mode;
  bool good_to_go = false;
  *iter = 0;
  *success = true;

  for (;;) {
    /* ------------------------------------------------ mode 2: Laguerre */
    if (mode >= 2) {
      one_nth = 1.0 / degree;
      n_1_nth = (degree - 1.0) * one_nth;
      two_n_div_n_1 = 2.0 / n_1_nth;
      c_one_nth = CMPLX(one_nth, 0.0);
      for (i = 1; i <= MAX_ITERS; i++) {
        faq = 1.0;
        ek = cabs(poly[degree]);
        absroot = cabs(*root);
        p = poly[degree];
        dp = zero;
        d2p_half = zero;
        for (int k = degree; k >= 1; k--) {
          d2p_half = dp + d2p_half * (*root);
          dp = p + dp * (*root);
          p = poly[k - 1] + p * (*root);
          ek = absroot * ek + cabs(p);
        }
        abs2p = abs2(p);
        *iter = *iter + 1;
        if (abs2p == 0.0) return;
        stopping_crit2 = (FRAC_ERR * ek) * (FRAC_ERR * ek);
        if (abs2p < stopping_crit2) {
          if (abs2p < 0.01 * stopping_crit2) return;
          good_to_go = true;
        } else {
          good_to_go = false;
        }
        denom = zero;
        if (!ceq(dp, zero)) {
          fac_netwon = p / dp;
          fac_extra = d2p_half / dp;
          F_half = fac_netwon * fac_extra;
          abs2_F_half = abs2(F_half);
          if (abs2_F_half <= 0.0625) {
            if (abs2_F_half <= 0.000625)
              mode = 0;
            else
              mode = 1;
          }
          denom_sqrt = csqrt(c_one - rmul(two_n_div_n_1, F_half));
          if (creal(denom_sqrt) >= 0.0)
            denom = c_one_nth + rmul(n_1_nth, denom_sqrt);
          else
            denom = c_one_nth - rmul(n_1_nth, denom_sqrt);
        }
        if (ceq(denom, zero)) {
          dx = rmul(cabs(*root) + 1.0, frac_jump_phase(i % FRAC_JUMP_LEN));
        } else {
          dx = fac_netwon / denom;
        }
        newroot = *root - dx;
        if (ceq(newroot, *root)) return;
        if (good_to_go) {
          *root = newroot;
          return;
        }
        if (mode != 2) {
          *root = newroot;
          j = i + 1;
          break;
        }
        if (i % FRAC_JUMP_EVERY == 0) {
          faq = FRAC_JUMPS[(i / FRAC_JUMP_EVERY - 1) % FRAC_JUMP_LEN];
          newroot = *root - rmul(faq, dx);
        }
        *root = newroot;
      }
      if (i >= MAX_ITERS) {
        *success = false;
        return;
      }
    }
    /* ------------------------------------------------ mode 1: SG */
    if (mode == 1) {
      for (i = j; i <= MAX_ITERS; i++) {
        faq = 1.0;
        p = poly[degree];
        dp = zero;
        d2p_half = zero;
        if ((i - j) % 10 == 0) {
          ek = cabs(poly[degree]);
          absroot = cabs(*root);
          for (int k = degree; k >= 1; k--) {
            d2p_half = dp + d2p_half * (*root);
            dp = p + dp * (*root);
            p = poly[k - 1] + p * (*root);
            ek = absroot * ek + cabs(p);
          }
          stopping_crit2 = (FRAC_ERR * ek) * (FRAC_ERR * ek);
        } else {
          for (int k = degree; k >= 1; k--) {
            d2p_half = dp + d2p_half * (*root);
            dp = p + dp * (*root);
            p = poly[k - 1] + p * (*root);
          }
        }
        abs2p = abs2(p);
        *iter = *iter + 1;
        if (abs2p == 0.0) return;
        if (abs2p < stopping_crit2) {
          if (ceq(dp, zero)) return;
          if (abs2p < 0.01 * stopping_crit2) return;
          good_to_go = true;
        } else {
          good_to_go = false;
        }
        if (ceq(dp, zero)) {
          dx = rmul(cabs(*root) + 1.0, frac_jump_phase(i % FRAC_JUMP_LEN));
        } else {
          fac_netwon = p / dp;
          fac_extra = d2p_half / dp;
          F_half = fac_netwon * fac_extra;
          abs2_F_half = abs2(F_half);
          if (abs2_F_half <= 0.000625) mode = 0;
          dx = fac_netwon * (c_one + F_half);
        }
        newroot = *root - dx;
        if (ceq(newroot, *root)) return;
        if (good_to_go) {
          *root = newroot;
          return;
        }
        if (mode != 1) {
          *root = newroot;
          j = i + 1;
          break;
        }
        if (i % FRAC_JUMP_EVERY == 0) {
          faq = FRAC_JUMPS[(i / FRAC_JUMP_EVERY - 1) % FRAC_JUMP_LEN];
          newroot = *root - rmul(faq, dx);
        }
        *root = newroot;
      }
      if (i >= MAX_ITERS) {
        *success = false;
        return;
      }
    }
    /* ------------------------------------------------ mode 0: Newton */
    if (mode == 0) {
      for (i = j; i <= j + 10; i++) {
        faq = 1.0;
        p = poly[degree];
        dp = zero;
        if (i == j) {
          ek = cabs(poly[degree]);
          absroot = cabs(*root);
          for (int k = degree; k >= 1; k--) {
            dp = p + dp * (*root);
            p = poly[k - 1] + p * (*root);
            ek = absroot * ek + cabs(p);
          }
          stopping_crit2 = (FRAC_ERR * ek) * (FRAC_ERR * ek);
        } else {
          for (int k = degree; k >= 1; k--) {
            dp = p + dp * (*root);
            p = poly[k - 1] + p * (*root);
          }
        }
        abs2p = abs2(p);
        *iter = *iter + 1;
        if (abs2p == 0.0) return;
        if (abs2p < stopping_crit2) {
          if (ceq(dp, zero)) return;
          if (abs2p < 0.01 * stopping_crit2) return;
          good_to_go = true;
        } else {
          good_to_go = false;
        }
        if (ceq(dp, zero)) {
          dx = rmul(cabs(*root) + 1.0, frac_jump_phase(i % FRAC_JUMP_LEN));
        } else {
          dx = p / dp;
        }
        newroot = *root - dx;
        if (ceq(newroot, *root)) return;
        if (good_to_go) {
          *root = newroot;
          return;
        }
        *root = newroot;
      }
      if (*iter >= MAX_ITERS) {
        *success = false;
        return;
      }
      mode = 2;
    }
  }
}

/* cmplx_roots_sg.f90:1311-1362 */
static void solve_quadratic_eq(cplx *x0, cplx *x1, const cplx *poly)
{
  cplx a = poly[2], b = poly[1], c = poly[0];
  cplx b2 = b * b;
  cplx delta = csqrt(b2 - rmul(4.0, a * c));
  if (creal(conj(b) * delta) >= 0.0)
    *x0 = rmul(-0.5, b + delta);
  else
    *x0 = rmul(-0.5, b - delta);
  if (ceq(*x0, CMPLX(0.0, 0.0))) {
    *x1 = CMPLX(0.0, 0.0);
  } else {
    *x1 = c / *x0;
    *x0 = *x0 / a;
  }
}

/* cmplx_roots_sg.f90:83-201 with polish_roots_after=.true., use_roots_as_starting_points=.false. */
static void cmplx_roots_gen(cplx *roots, const cplx *poly, int degree, gor_trace *tr)
{
  cplx poly2[5];
  const cplx zero = CMPLX(0.0, 0.0);
  int iter;
  bool success;
  cplx coef, prev;
  for (int i = 0; i <= degree; i++) poly2[i] = poly[i];
  for (int i = 0; i < degree; i++) roots[i] = zero;
  if (degree <= 1) {
    if (degree == 1) roots[0] = -poly[0] / poly[1];
    return;
  }
  for (int n = degree; n >= 3; n--) {
    cmplx_laguerre2newton(poly2, n, &roots[n - 1], &iter, &success, 2);
    if (tr) tr->n_solver_iters += iter;
    if (!success) {
      roots[n - 1] = zero;
      cmplx_laguerre(poly2, n, &roots[n - 1], &iter, &success);
      if (tr) tr->n_solver_iters += iter;
    }
    coef = poly2[n];
    for (int i = n; i >= 1; i--) {
      prev = poly2[i - 1];
      poly2[i - 1] = coef;
      coef = prev + roots[n - 1] * coef;
    }
  }
  cmplx_laguerre2newton(poly2, 2, &roots[1], &iter, &success, 2);
  if (tr) tr->n_solver_iters += iter;
  if (!success) {
    solve_quadratic_eq(&roots[1], &roots[0], poly2);
  } else {
    roots[0] = -(roots[1] + poly2[1] / poly2[2]);
  }
  for (int n = 0; n < degree; n++) {
    cmplx_laguerre(poly, degree, &roots[n], &iter, &success);
    if (tr) tr->n_solver_iters += iter;
  }
  if (tr) tr->n_solver_calls += 1;
}

void gor_cmplx_roots_gen(int degree, const double *poly_re_im, double *roots_re_im)
{
  cplx poly[5], roots[4];
  for (int i = 0; i <= degree; i++) poly[i] = CMPLX(poly_re_im[2 * i], poly_re_im[2 * i + 1]);
  cmplx_roots_gen(roots, poly, degree, NULL);
  for (int i = 0; i < degree; i++) {
    roots_re_im[2 * i] = creal(roots[i]);
    roots_re_im[2 * i + 1] = cimag(roots[i]);
  }
}

/* SRC/contrib/Polynomial234RootSolvers.f90:84-142 ; root is root(n,2) column-major */
static void pack_roots(const cplx *croots, int n, int *nReal, double *root)
{
  const double cmplx_tol = 1.0e-12;
  double re_part[4], im_part[4];
  bool is_real[4];
  int cnt = 0;
  for (int i = 0; i < n; i++) {
    re_part[i] = creal(croots[i]);
    im_part[i] = cimag(croots[i]);
    double tol_i = cmplx_tol * fmax(1.0, fabs(re_part[i]));
    is_real[i] = fabs(im_part[i]) <= tol_i;
    if (is_real[i]) cnt++;
  }
  *nReal = cnt;
  for (int i = 0; i < 2 * n; i++) root[i] = 0.0;
  int idx = 0;
  for (int i = 0; i < n; i++)
    if (is_real[i]) {
      root[idx] = re_part[i];
      root[n + idx] = 0.0;
      idx++;
    }
  /* sort_real_descending: insertion sort on the first nReal rows */
  for (int i = 1; i < cnt; i++) {
    double tmp_re = root[i], tmp_im = root[n + i];
    int j = i - 1;
    while (j >= 0) {
      if (root[j] >= tmp_re) break;
      root[j + 1] = root[j];
      root[n + j + 1] = root[n + j];
      j--;
    }
    root[j + 1] = tmp_re;
    root[n + j + 1] = tmp_im;
  }
  for (int i = 0; i < n; i++)
    if (!is_real[i]) {
      root[idx] = re_part[i];
      root[n + idx] = im_part[i];
      idx++;
    }
}

static void quadraticRoots(double q1, double q0, int *nReal, double *root, gor_trace *tr)
{
  cplx poly[3] = {CMPLX(q0, 0.0), CMPLX(q1, 0.0), CMPLX(1.0, 0.0)}, croots[2];
  cmplx_roots_gen(croots, poly, 2, tr);
  pack_roots(croots, 2, nReal, root);
}
static void cubicRoots(double c2, double c1, double c0, int *nReal, double *root, gor_trace *tr)
{
  cplx poly[4] = {CMPLX(c0, 0.0), CMPLX(c1, 0.0), CMPLX(c2, 0.0), CMPLX(1.0, 0.0)}, croots[3];
  cmplx_roots_gen(croots, poly, 3, tr);
  pack_roots(croots, 3, nReal, root);
}
static void quarticRoots(double q3, double q2, double q1, double q0, int *nReal, double *root,
                         gor_trace *tr)
{
  cplx poly[5] = {CMPLX(q0, 0.0), CMPLX(q1, 0.0), CMPLX(q2, 0.0), CMPLX(q3, 0.0), CMPLX(1.0, 0.0)},
       croots[4];
  cmplx_roots_gen(croots, poly, 4, tr);
  pack_roots(croots, 4, nReal, root);
}
void gor_quadratic_roots(double q1, double q0, int *nreal, double root[4])
{
  quadraticRoots(q1, q0, nreal, root, NULL);
}
void gor_cubic_roots(double c2, double c1, double c0, int *nreal, double root[6])
{
  cubicRoots(c2, c1, c0, nreal, root, NULL);
}
void gor_quartic_roots(double q3, double q2, double q1, double q0, int *nreal, double root[8])
{
  quarticRoots(q3, q2, q1, q0, nreal, root, NULL);
}

/* ------------------------------------------------------------------------------------------------
 * Exit-time solvers -- SRC/pusher_tetra_poly.f90:1767-2021
 * ---------------------------------------------------------------------------------------------- */
static double Linear_Solver(double a, double b)
{
  if (a == 0.0) return HUGE_D;
  return -b / a;
}

/* :1809-1893   f(tau) = a/2 tau^2 + b tau + c */
static double Quadratic_Solver1(double acoef, double bcoef, double ccoef)
{
  double dtau = HUGE_D, discr, dummy;
  if (ccoef > 0.0) {
    if (acoef > 0.0) {
      if (bcoef < 0.0) {
        discr = bcoef * bcoef - 2.0 * acoef * ccoef;
        if (discr > 0.0) {
          dummy = (-bcoef + sqrt(discr));
          if (fabs(dummy) > EPS)
            dtau = 2.0 * ccoef / dummy;
          else
            dtau = (-sqrt(discr) - bcoef) / acoef;
        } else if (discr == 0.0) {
          dtau = -bcoef / acoef;
        } else {
          return dtau;
        }
      } else {
        return dtau;
      }
    } else if (acoef < 0.0) {
      discr = bcoef * bcoef - 2.0 * acoef * ccoef;
      dummy = (-bcoef + sqrt(discr));
      if (fabs(dummy) > EPS)
        dtau = 2.0 * ccoef / dummy;
      else
        dtau = (-sqrt(discr) - bcoef) / acoef;
    } else {
      if (bcoef < 0.0)
        dtau = -ccoef / bcoef;
      else
        return dtau;
    }
  } else if (ccoef < 0.0) {
    if (acoef < 0.0) {
      if (bcoef > 0.0) {
        discr = bcoef * bcoef - 2.0 * acoef * ccoef;
        if (discr > 0.0)
          dtau = (sqrt(discr) - bcoef) / acoef;
        else if (discr == 0.0)
          dtau = -bcoef / acoef;
        else
          return dtau;
      } else {
        return dtau;
      }
    } else if (acoef > 0.0) {
      discr = bcoef * bcoef - 2.0 * acoef * ccoef;
      dtau = (sqrt(discr) - bcoef) / acoef;
    } else {
      if (bcoef > 0.0)
        dtau = -ccoef / bcoef;
      else
        return dtau;
    }
  } else {
    if (((acoef > 0.0) && (bcoef < 0.0)) || ((acoef < 0.0) && (bcoef > 0.0)))
      dtau = -2.0 * bcoef / acoef;
    else
      return dtau;
  }
  return dtau;
}

/* minval(root(:,1),1,mask) with mask = abs(Im)==0 .and. Re>0 ; empty mask -> huge */
static double min_positive_real(const double *root, int n, double lambda)
{
  double best = HUGE_D;
  for (int i = 0; i < n; i++) {
    double re = root[i] / lambda, im = root[n + i] / lambda;
    if (fabs(im) == 0.0 && re > 0.0)
      if (re < best) best = re;
  }
  return best;
}

/* :1897-1930 */
static double Quadratic_Solver2(double a, double b, double c, gor_trace *tr)
{
  double root[4], lambda, q0, q1;
  int nReal;
  lambda = b / c;
  q0 = 2.0 * (b * b) / (a * c);
  q1 = q0;
  quadraticRoots(q1, q0, &nReal, root, tr);
  return min_positive_real(root, 2, lambda);
}
/* :1934-1967 */
static double Cubic_Solver(double a, double b, double c, double d, gor_trace *tr)
{
  double root[6], lambda, c2, c1, c0;
  int nReal;
  lambda = b / (2.0 * c);
  c2 = 3.0 * lambda * b / a;
  c1 = 6.0 * c * (lambda * lambda) / a;
  c0 = 6.0 * d * ((lambda * lambda) * lambda) / a;
  cubicRoots(c2, c1, c0, &nReal, root, tr);
  return min_positive_real(root, 3, lambda);
}
/* :1971-2021 */
static double Quartic_Solver(int i_scaling, double a, double b, double c, double d, double e,
                             gor_trace *tr)
{
  double root[8], lambda = 0, q3, q2, q1, q0;
  int nReal;
  switch (i_scaling) {
    case 0: lambda = sqrt(fabs(b / (6.0 * d))); break;
    case 1: lambda = b / (3.0 * c); break;
    case 2: lambda = pow(fabs(b / (6.0 * e)), 1.0 / 3.0); break;
    case 3: lambda = c / (2.0 * d); break;
    case 4: lambda = sqrt(fabs(c / (2.0 * e))); break;
    case 5: lambda = d / e; break;
    case 6: lambda = pow(fabs(a / (24.0 * e)), 1.0 / 4.0); break;
  }
  double l2 = lambda * lambda;
  q3 = 4.0 * b * lambda / a;
  q2 = 12.0 * c * l2 / a;
  q1 = 24.0 * d * (l2 * lambda) / a;
  q0 = 24.0 * e * (l2 * l2) / a;
  quarticRoots(q3, q2, q1, q0, &nReal, root, tr);
  return min_positive_real(root, 4, lambda);
}
double gor_quadratic_solver1(double a, double b, double c) { return Quadratic_Solver1(a, b, c); }
double gor_quadratic_solver2(double a, double b, double c) { return Quadratic_Solver2(a, b, c, NULL); }
double gor_cubic_solver(double a, double b, double c, double d) { return Cubic_Solver(a, b, c, d, NULL); }
double gor_quartic_solver(int s, double a, double b, double c, double d, double e)
{
  return Quartic_Solver(s, a, b, c, d, e, NULL);
}

/* ------------------------------------------------------------------------------------------------
 * small helpers -- SRC/supporting_functions_mod.f90
 * ---------------------------------------------------------------------------------------------- */
static inline const double *rec(const gor_mesh *m, int ind_tetr)
{
  return m->tetra_physics + (int64_t)(ind_tetr - 1) * TP_NDOUBLES;
}
static inline const int32_t *grd(const gor_mesh *m, int ind_tetr)
{
  return m->tetra_grid + (int64_t)(ind_tetr - 1) * TG_NINTS;
}
static inline double dot3(const double *a, const double *b) /* sum(a*b), ascending */
{
  return ((0.0 + a[0] * b[0]) + a[1] * b[1]) + a[2] * b[2];
}
static inline int isign1(double t) { return signbit(t) ? -1 : 1; } /* int(sign(1.d0,t)) */

double gor_bmod(const gor_mesh *m, const double z[3], int32_t ind_tetr) /* :305 */
{
  const double *r = rec(m, ind_tetr);
  return r[TP_BMOD1] + dot3(r + TP_GB, z);
}
static double phi_elec_func(const gor_mesh *m, const double z[3], int ind_tetr) /* :342 */
{
  const double *r = rec(m, ind_tetr);
  return r[TP_PHI1] + dot3(r + TP_GPHI, z);
}
static double v2_E_mod_func(const gor_mesh *m, const double z[3], int ind_tetr) /* :445 */
{
  const double *r = rec(m, ind_tetr);
  return r[TP_V2EMOD_1] + dot3(z, r + TP_GV2EMOD);
}
static double vperp_func(const gor_mesh *m, const double z[3], double perpinv, int ind_tetr) /* :321 */
{
  if (perpinv != 0.0) return sqrt(2.0 * fabs(perpinv) * gor_bmod(m, z, ind_tetr));
  return 0.0;
}
double gor_energy_tot(const gor_mesh *m, const double z[4], double perpinv, int32_t ind_tetr) /* :279 */
{
  const double *r = rec(m, ind_tetr);
  double vperp = sqrt(2.0 * fabs(perpinv) * (r[TP_BMOD1] + dot3(r + TP_GB, z)));
  double e = m->particle_mass / 2.0 * (vperp * vperp + z[3] * z[3]) +
             m->particle_charge * phi_elec_func(m, z, ind_tetr);
  if (m->boole_strong_electric_field) e = e + 0.5 * m->particle_mass * v2_E_mod_func(m, z, ind_tetr);
  return e;
}
double gor_p_phi(const gor_mesh *m, double vpar, const double z[3], int32_t ind_tetr) /* :377 */
{
  const double *r = rec(m, ind_tetr);
  double hphi1;
  const double *ghphi;
  if (m->coord_system == 1) {
    hphi1 = r[TP_H2_1];
    ghphi = r + TP_GH2;
  } else {
    hphi1 = r[TP_H3_1];
    ghphi = r + TP_GH3;
  }
  double p = m->particle_mass * vpar * (hphi1 + dot3(ghphi, z)) +
             m->particle_mass / m->cm_over_e * (r[TP_APHI1] + dot3(r + TP_GAPHI, z));
  if (m->boole_strong_electric_field) {
    double vE2 = r[TP_VE2_1] + dot3(z, r + TP_GVE2);
    p = p + m->particle_mass * vE2;
  }
  return p;
}

/* ------------------------------------------------------------------------------------------------
 * pusher_handover2neighbour -- SRC/pusher_tetra_func_mod.f90:6-93 (handover_processing_kind = 1)
 * ---------------------------------------------------------------------------------------------- */
static void handover2neighbour(const gor_mesh *m, int ind_tetr, int *ind_tetr_out, int *iface_inout,
                               double x[3], int *iper_phi)
{
  const int32_t *g = grd(m, ind_tetr);
  int iface = *iface_inout;
  *ind_tetr_out = g[TG_NEIGHBOUR_TETR + iface - 1];
  *iface_inout = g[TG_NEIGHBOUR_FACE + iface - 1];
  *iper_phi = g[TG_PERBOU_PHI + iface - 1];
  int iper_theta = g[TG_PERBOU_THETA + iface - 1];
  if (m->handover_processing_kind == 2) { /* position exchange via Cartesian variables (skew coordinates), :59-89 */
    /* type tetrahedron_skew_coord, column-major: skew_coord_x1x2x3(3,3,4) @0, skew_coord_xyz @36, inv_skew_coord_x1x2x3 @72,
     * inv_skew_coord_xyz @108, skew_ref_x1x2x3(3,4) @144, skew_ref_xyz(3,4) @156.  A particle that leaves the domain
     * (neighbour -1) keeps its exit position: the reference indexes tetra_skew_coord(-1) there. */
    const int iface_out = *iface_inout;
    if (*ind_tetr_out < 1) return;
    const double *A = m->tetra_skew_coord + (int64_t)(ind_tetr - 1) * 168, *B = m->tetra_skew_coord + (int64_t)(*ind_tetr_out - 1) * 168;
    const int k = iface - 1, g = iface_out - 1;
    double b[3], t[3], x_lin_cart[3];
    for (int i = 0; i < 3; i++) b[i] = x[i] - A[144 + 3 * k + i];
    for (int i = 0; i < 3; i++) { /* matmul(inv_skew_coord_x1x2x3(:,:,iface), b) */
      double acc = 0.0;
      for (int j = 0; j < 3; j++) acc = acc + A[72 + 9 * k + i + 3 * j] * b[j];
      t[i] = acc;
    }
    for (int i = 0; i < 3; i++) { /* matmul(skew_coord_xyz(:,:,iface), .) */
      double acc = 0.0;
      for (int j = 0; j < 3; j++) acc = acc + A[36 + 9 * k + i + 3 * j] * t[j];
      x_lin_cart[i] = acc;
    }
    for (int i = 0; i < 3; i++) x_lin_cart[i] = x_lin_cart[i] + A[156 + 3 * k + i];
    for (int i = 0; i < 3; i++) b[i] = x_lin_cart[i] - B[156 + 3 * g + i];
    for (int i = 0; i < 3; i++) { /* matmul(inv_skew_coord_xyz(:,:,iface_out), b) */
      double acc = 0.0;
      for (int j = 0; j < 3; j++) acc = acc + B[108 + 9 * g + i + 3 * j] * b[j];
      t[i] = acc;
    }
    for (int i = 0; i < 3; i++) { /* matmul(skew_coord_x1x2x3(:,:,iface_out), .) */
      double acc = 0.0;
      for (int j = 0; j < 3; j++) acc = acc + B[9 * g + i + 3 * j] * t[j];
      x[i] = acc;
    }
    for (int i = 0; i < 3; i++) x[i] = x[i] + B[144 + 3 * g + i];
    return;
  }
  if (m->coord_system == 1) {
    if (*iper_phi == 1)
      x[1] = x[1] - 2.0 * PI / m->n_field_periods;
    else if (*iper_phi == -1)
      x[1] = x[1] + 2.0 * PI / m->n_field_periods;
  } else {
    if (*iper_phi == 1)
      x[2] = x[2] - 2.0 * PI / m->n_field_periods;
    else if (*iper_phi == -1)
      x[2] = x[2] + 2.0 * PI / m->n_field_periods;
    if (iper_theta == 1)
      x[1] = x[1] - 2.0 * PI;
    else if (iper_theta == -1)
      x[1] = x[1] + 2.0 * PI;
  }
}

/* ------------------------------------------------------------------------------------------------
 * pusher_tetra_poly -- SRC/pusher_tetra_poly.f90
 * (i_precomp = 0, i_time_tracing_option = 1 or 2, boole_adaptive_time_steps = .false.)
 * The THREADPRIVATE module variables (:6-13,45-61) live in this struct, one per particle.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const gor_mesh *m;
  const double *r; /* current record */
  int ind_tetr, iface_init, sign_rhs;
  double perpinv, perpinv2, vmod0, dt_dtau_const, bmod0, t_remain;
  double z_init[4], k1, k3;
  double b[4], amat[4][4], amat2[4][4], amat3[4][4], amat4[4][4]; /* amat[i][j] = amat(i+1,j+1) */
  double amat_in_z[4], amat2_in_z[4], amat3_in_z[4], amat4_in_z[4];
  double amat_in_b[4], amat2_in_b[4], amat3_in_b[4];
  int number_of_integration_steps;
  /* tau_steps_list / intermediate_z0_list (:36-38); non-adaptive scheme: two entries (:98-104) */
  double *tau_steps_list, (*intermediate_z0_list)[4];
  int list_cap; /* 2, or 3*max_n_intermediate_steps with the adaptive scheme (:98-104) */
  double list_tau_static[2], list_z0_static[2][4];
  double *thl_heap; /* t_hamiltonian_list (:471) when the lists are the long ones of the adaptive scheme: list_cap + 1 entries */
  bool removed; /* the push ended on one of the 'remove particle' returns */
  gor_trace *tr;
} poly_state;
static const double eps_tau = 100.0;

static void matmul44(double c[4][4], double a[4][4], double b[4][4])
{
  double t[4][4];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      double s = 0.0;
      for (int k = 0; k < 4; k++) s = s + a[i][k] * b[k][j];
      t[i][j] = s;
    }
  memcpy(c, t, sizeof(t));
}
static void matvec4(double c[4], double a[4][4], const double v[4])
{
  double t[4];
  for (int i = 0; i < 4; i++) {
    double s = 0.0;
    for (int k = 0; k < 4; k++) s = s + a[i][k] * v[k];
    t[i] = s;
  }
  memcpy(c, t, sizeof(t));
}
static inline const double *anorm_col(const poly_state *s, int iface) /* anorm(:,iface) */
{
  return s->r + TP_ANORM + 3 * (iface - 1);
}

/* :125-178 */
static void initialize_pusher_tetra_poly(poly_state *s, int ind_tetr, const double x[3], int iface,
                                         double vpar, double t_remain_in)
{
  const gor_mesh *m = s->m;
  s->t_remain = t_remain_in;
  s->ind_tetr = ind_tetr;
  s->r = rec(m, ind_tetr);
  s->sign_rhs = m->sign_sqg * isign1(s->t_remain);
  for (int i = 0; i < 3; i++) s->z_init[i] = x[i] - s->r[TP_X1 + i];
  s->z_init[3] = vpar;
  s->iface_init = iface;
  s->dt_dtau_const = s->r[TP_DT_DTAU_CONST];
  s->dt_dtau_const = s->dt_dtau_const * (double)s->sign_rhs;
  s->bmod0 = gor_bmod(m, s->z_init, ind_tetr);
  double phi_elec = phi_elec_func(m, s->z_init, ind_tetr);
  double vperp2 = -2.0 * s->perpinv * s->bmod0;
  double vpar2 = vpar * vpar;
  s->vmod0 = sqrt(vpar2 + vperp2);
  s->k1 = vperp2 + vpar2 + 2.0 * s->perpinv * s->r[TP_BMOD1];
  if (m->boole_strong_electric_field)
    s->k1 = s->k1 + (v2_E_mod_func(m, s->z_init, ind_tetr) - s->r[TP_V2EMOD_1]);
  s->k3 = s->r[TP_PHI1] - phi_elec;
}

/* :1486-1586 ; coef_mat[n][k] = coef_mat(n+1,k+1) */
static void analytic_coeff_without_precomp(poly_state *s, int poly_order, const bool boole_faces[4],
                                           const double z[4], double coef_mat[4][5])
{
  const gor_mesh *m = s->m;
  const double *r = s->r;
  const double cm_over_e = m->cm_over_e, perpinv = s->perpinv;
  for (int i = 0; i < 3; i++)
    s->b[i] = (r[TP_CURLH + i] * (s->k1) + perpinv * r[TP_GBXH1 + i]) * cm_over_e -
              CLIGHT * (2.0 * (s->k3) * r[TP_CURLH + i] + r[TP_GPHIXH1 + i]);
  s->b[3] = perpinv * r[TP_GBXCURLA] - CLIGHT / cm_over_e * r[TP_GPHIXCURLA];
  memset(s->amat, 0, sizeof(s->amat));
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) /* alpmat(i,j) column-major: [i + 3*j] */
      s->amat[i][j] = perpinv * cm_over_e * r[TP_ALPMAT + i + 3 * j] - CLIGHT * r[TP_BETMAT + i + 3 * j];
  s->amat[3][3] = perpinv * cm_over_e * r[TP_SPALPMAT] - CLIGHT * r[TP_SPBETMAT];
  for (int i = 0; i < 3; i++) s->amat[i][3] = r[TP_CURLA + i];
  if (m->boole_strong_electric_field) {
    for (int i = 0; i < 3; i++) s->b[i] = s->b[i] - 0.5 * cm_over_e * r[TP_GV2EMODXH1 + i];
    s->b[3] = s->b[3] + cm_over_e * perpinv * r[TP_GBXCURLVE] - CLIGHT * r[TP_GPHIXCURLVE] -
              0.5 * cm_over_e * r[TP_GV2EMODXCURLVE] - 0.5 * r[TP_GV2EMODXCURLA];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) s->amat[i][j] = s->amat[i][j] - 0.5 * cm_over_e * r[TP_GAMMAT + i + 3 * j];
    s->amat[3][3] = s->amat[3][3] - 0.5 * cm_over_e * r[TP_SPGAMMAT];
    for (int i = 0; i < 3; i++) s->amat[i][3] = s->amat[i][3] + cm_over_e * r[TP_CURLVE + i];
  }
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) s->amat[i][j] = s->amat[i][j] * (double)s->sign_rhs;
  for (int i = 0; i < 4; i++) s->b[i] = s->b[i] * (double)s->sign_rhs;
  double dist1 = -r[TP_DIST_REF];

  for (int n = 0; n < 4; n++) {
    if (!boole_faces[n]) continue;
    coef_mat[n][0] = dot3(r + TP_ANORM + 3 * n, z);
  }
  coef_mat[0][0] = coef_mat[0][0] - dist1;
  if (poly_order >= 1) {
    matvec4(s->amat_in_z, s->amat, z);
    for (int n = 0; n < 4; n++) {
      if (!boole_faces[n]) continue;
      coef_mat[n][1] = dot3(r + TP_ANORM + 3 * n, s->amat_in_z) + dot3(r + TP_ANORM + 3 * n, s->b);
    }
  }
  if (poly_order >= 2) {
    matmul44(s->amat2, s->amat, s->amat);
    matvec4(s->amat2_in_z, s->amat2, z);
    matvec4(s->amat_in_b, s->amat, s->b);
    for (int n = 0; n < 4; n++) {
      if (!boole_faces[n]) continue;
      coef_mat[n][2] = dot3(r + TP_ANORM + 3 * n, s->amat2_in_z) + dot3(r + TP_ANORM + 3 * n, s->amat_in_b);
    }
  }
  if (poly_order >= 3) {
    matmul44(s->amat3, s->amat, s->amat2);
    matvec4(s->amat3_in_z, s->amat3, z);
    matvec4(s->amat2_in_b, s->amat2, s->b);
    for (int n = 0; n < 4; n++) {
      if (!boole_faces[n]) continue;
      coef_mat[n][3] = dot3(r + TP_ANORM + 3 * n, s->amat3_in_z) + dot3(r + TP_ANORM + 3 * n, s->amat2_in_b);
    }
  }
  if (poly_order >= 4) {
    matmul44(s->amat4, s->amat, s->amat3);
    matvec4(s->amat4_in_z, s->amat4, z);
    matvec4(s->amat3_in_b, s->amat3, s->b);
    for (int n = 0; n < 4; n++) {
      if (!boole_faces[n]) continue;
      coef_mat[n][4] = dot3(r + TP_ANORM + 3 * n, s->amat4_in_z) + dot3(r + TP_ANORM + 3 * n, s->amat3_in_b);
    }
  }
}

/* ------------------------------------------------------------------------------------------------
 * Precomputed-coefficient modes: type tetrahedron_physics_precomp_poly4 and make_precomp_poly4
 * (SRC/tetra_physics_poly_precomp_mod.f90:21-45,160-476).  Matrices are handled as m[i][j] = M(i+1,j+1) and stored in the
 * record in Fortran order (P4_* offsets, gorilla_oracle.h).  matmul / sum accumulate in ascending index order from 0.
 * ---------------------------------------------------------------------------------------------- */
static void p4_store(double *rec4, int off, double a[4][4])
{
  for (int j = 0; j < 4; j++)
    for (int i = 0; i < 4; i++) rec4[off + i + 4 * j] = a[i][j];
}
static void p4_load(const double *rec4, int off, double a[4][4])
{
  for (int j = 0; j < 4; j++)
    for (int i = 0; i < 4; i++) a[i][j] = rec4[off + i + 4 * j];
}
static void add44(double c[4][4], double a[4][4], double b[4][4])
{
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) c[i][j] = a[i][j] + b[i][j];
}
static void make_precomp_poly4_one(const gor_mesh *m, int ind_tetr, double *o)
{
  const double *r = rec(m, ind_tetr);
  const double cm = m->cm_over_e;
  double alp[4][4], bet[4][4];
  memset(alp, 0, sizeof(alp));
  memset(bet, 0, sizeof(bet));
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      alp[i][j] = cm * r[TP_ALPMAT + i + 3 * j];
      bet[i][j] = -(CLIGHT * r[TP_BETMAT + i + 3 * j]);
    }
  alp[3][3] = cm * r[TP_SPALPMAT];
  for (int i = 0; i < 3; i++) bet[i][3] = r[TP_CURLA + i];
  bet[3][3] = -(CLIGHT * r[TP_SPBETMAT]);
  double aa[4][4], ab[4][4], bb[4][4], ba[4][4];
  matmul44(aa, alp, alp); matmul44(ab, alp, bet); matmul44(bb, bet, bet); matmul44(ba, bet, alp);
  double aaa[4][4], aab[4][4], aba[4][4], abb[4][4], baa[4][4], bab[4][4], bba[4][4], bbb[4][4];
  matmul44(aaa, alp, aa); matmul44(aab, alp, ab); matmul44(aba, alp, ba); matmul44(abb, alp, bb);
  matmul44(baa, bet, aa); matmul44(bab, bet, ab); matmul44(bba, bet, ba); matmul44(bbb, bet, bb);
  double aaaa[4][4], aaab[4][4], aaba[4][4], aabb[4][4], abaa[4][4], abab[4][4], abba[4][4], abbb[4][4];
  double baaa[4][4], baab[4][4], baba[4][4], babb[4][4], bbaa[4][4], bbab[4][4], bbba[4][4], bbbb[4][4];
  matmul44(aaaa, alp, aaa); matmul44(aaab, alp, aab); matmul44(aaba, alp, aba); matmul44(aabb, alp, abb);
  matmul44(abaa, alp, baa); matmul44(abab, alp, bab); matmul44(abba, alp, bba); matmul44(abbb, alp, bbb);
  matmul44(baaa, bet, aaa); matmul44(baab, bet, aab); matmul44(baba, bet, aba); matmul44(babb, bet, abb);
  matmul44(bbaa, bet, baa); matmul44(bbab, bet, bab); matmul44(bbba, bet, bba); matmul44(bbbb, bet, bbb);
  double t[4][4];
  p4_store(o, P4_AMAT1_0, bet);
  p4_store(o, P4_AMAT1_1, alp);
  p4_store(o, P4_AMAT2_0, bb);
  add44(t, ba, ab); p4_store(o, P4_AMAT2_1, t);                               /* bet_alp + alp_bet */
  p4_store(o, P4_AMAT2_2, aa);
  p4_store(o, P4_AMAT3_0, bbb);
  add44(t, bba, bab); add44(t, t, abb); p4_store(o, P4_AMAT3_1, t);           /* bet_bet_alp + bet_alp_bet + alp_bet_bet */
  add44(t, baa, aba); add44(t, t, aab); p4_store(o, P4_AMAT3_2, t);           /* bet_alp_alp + alp_bet_alp + alp_alp_bet */
  p4_store(o, P4_AMAT3_3, aaa);
  p4_store(o, P4_AMAT4_0, bbbb);
  add44(t, bbba, bbab); add44(t, t, babb); add44(t, t, abbb); p4_store(o, P4_AMAT4_1, t);
  add44(t, bbaa, baba); add44(t, t, baab); add44(t, t, abba); add44(t, t, abab); add44(t, t, aabb);
  p4_store(o, P4_AMAT4_2, t);
  add44(t, abaa, aaba); add44(t, t, aaab); add44(t, t, baaa); p4_store(o, P4_AMAT4_3, t);
  p4_store(o, P4_AMAT4_4, aaaa);
  /* n in a^k: anorm_in_amatK(:,n) = matmul(n_vec, amatK), n_vec = (anorm(:,n), 0) */
  for (int k = 0; k < 14; k++) {
    double a[4][4];
    p4_load(o, 16 * k, a);
    for (int n = 0; n < 4; n++) {
      const double nv[4] = {r[TP_ANORM + 3 * n], r[TP_ANORM + 3 * n + 1], r[TP_ANORM + 3 * n + 2], 0.0};
      for (int j = 0; j < 4; j++) {
        double sacc = 0.0;
        for (int i = 0; i < 4; i++) sacc = sacc + nv[i] * a[i][j];
        o[P4_AN_AMAT1_0 + 16 * k + j + 4 * n] = sacc;
      }
    }
  }
  /* factorised b-vector */
  for (int i = 0; i < 3; i++) {
    o[P4_B0 + i] = -CLIGHT * r[TP_GPHIXH1 + i];
    o[P4_B1 + i] = cm * r[TP_CURLH + i];
    o[P4_B2 + i] = cm * r[TP_GBXH1 + i];
    o[P4_B3 + i] = -2.0 * CLIGHT * r[TP_CURLH + i];
  }
  o[P4_B0 + 3] = -CLIGHT / cm * r[TP_GPHIXCURLA];
  o[P4_B1 + 3] = 0.0;
  o[P4_B2 + 3] = r[TP_GBXCURLA];
  o[P4_B3 + 3] = 0.0;
  for (int q = 0; q < 2; q++) {    /* amat1_0, amat1_1 in b0..b3 */
    double a[4][4];
    p4_load(o, q == 0 ? P4_AMAT1_0 : P4_AMAT1_1, a);
    for (int k = 0; k < 4; k++) matvec4(o + (q == 0 ? P4_A10_B0 : P4_A11_B0) + 4 * k, a, o + P4_B0 + 4 * k);
  }
  for (int k = 0; k < 4; k++)
    for (int n = 0; n < 4; n++) {
      o[P4_AN_B0 + 4 * k + n] = dot3(r + TP_ANORM + 3 * n, o + P4_B0 + 4 * k);
      for (int q = 0; q < 2; q++) {
        const double *col = o + (q == 0 ? P4_AN_AMAT1_0 : P4_AN_AMAT1_1) + 4 * n, *bk = o + P4_B0 + 4 * k;
        double sacc = 0.0;
        for (int i = 0; i < 4; i++) sacc = sacc + col[i] * bk[i];
        o[(q == 0 ? P4_AN_A10_B0 : P4_AN_A11_B0) + 4 * k + n] = sacc;
      }
    }
}
void gor_make_precomp_poly4(const gor_mesh *m, double *out)
{
#pragma omp parallel for schedule(static)
  for (int64_t t = 1; t <= m->ntetr; t++) make_precomp_poly4_one(m, (int)t, out + (t - 1) * P4_NDOUBLES);
}
static inline const double *p4rec(const poly_state *s)
{
  return s->m->tetra_physics_poly4 + ((int64_t)s->ind_tetr - 1) * P4_NDOUBLES;
}
/* sum over the four components of (c0 + f1*c1 + f2*c2 + ...)(:) * v(:): the shape of every precomputed coefficient */
static double p4_combo_dot(const double *p4, const int *off, const double *fac, int nterms, int n, const double v[4])
{
  double sacc = 0.0;
  for (int i = 0; i < 4; i++) {
    double e = p4[off[0] + i + 4 * n];
    for (int k = 1; k < nterms; k++) e = e + fac[k] * p4[off[k] + i + 4 * n];
    sacc = sacc + e * v[i];
  }
  return sacc;
}
/* analytic_coeff_with_precomp (:1590-1725); i_precomp = 2 exists for orders <= 2 only (the reference leaves the higher
 * coefficients unassigned) */
static void analytic_coeff_with_precomp(poly_state *s, int poly_order, int i_precomp, const bool boole_faces[4],
                                        const double z[4], double coef_mat[4][5])
{
  const gor_mesh *m = s->m;
  const double *r = s->r, *p4 = p4rec(s);
  const double cm_over_e = m->cm_over_e, perpinv = s->perpinv, perpinv2 = s->perpinv2;
  if (i_precomp == 1) { /* b without sign_rhs (:1607-1614) */
    for (int i = 0; i < 3; i++)
      s->b[i] = (r[TP_CURLH + i] * (s->k1) + perpinv * r[TP_GBXH1 + i]) * cm_over_e -
                CLIGHT * (2.0 * (s->k3) * r[TP_CURLH + i] + r[TP_GPHIXH1 + i]);
    s->b[3] = perpinv * r[TP_GBXCURLA] - CLIGHT / cm_over_e * r[TP_GPHIXCURLA];
  }
  const double dist1 = -r[TP_DIST_REF];
  const double perpinv3 = perpinv2 * perpinv, perpinv4 = perpinv2 * perpinv2;
  const double fac[5] = {1.0, perpinv, perpinv2, perpinv3, perpinv4};
  static const int A1[2] = {P4_AN_AMAT1_0, P4_AN_AMAT1_1}, A2[3] = {P4_AN_AMAT2_0, P4_AN_AMAT2_1, P4_AN_AMAT2_2},
                   A3[4] = {P4_AN_AMAT3_0, P4_AN_AMAT3_1, P4_AN_AMAT3_2, P4_AN_AMAT3_3},
                   A4[5] = {P4_AN_AMAT4_0, P4_AN_AMAT4_1, P4_AN_AMAT4_2, P4_AN_AMAT4_3, P4_AN_AMAT4_4};
  for (int n = 0; n < 4; n++) {
    if (!boole_faces[n]) continue;
    coef_mat[n][0] = dot3(r + TP_ANORM + 3 * n, z);
  }
  coef_mat[0][0] = coef_mat[0][0] - dist1;
  for (int n = 0; n < 4; n++) {
    if (!boole_faces[n]) continue;
    if (poly_order >= 1) {
      const double sz = p4_combo_dot(p4, A1, fac, 2, n, z);
      if (i_precomp == 1)
        coef_mat[n][1] = sz + dot3(r + TP_ANORM + 3 * n, s->b);
      else
        coef_mat[n][1] = sz + p4[P4_AN_B0 + n] + s->k1 * p4[P4_AN_B1 + n] + perpinv * p4[P4_AN_B2 + n] + s->k3 * p4[P4_AN_B3 + n];
    }
    if (poly_order >= 2) {
      const double sz = p4_combo_dot(p4, A2, fac, 3, n, z);
      if (i_precomp == 1)
        coef_mat[n][2] = sz + p4_combo_dot(p4, A1, fac, 2, n, s->b);
      else
        coef_mat[n][2] = sz + p4[P4_AN_A10_B0 + n] + perpinv * p4[P4_AN_A11_B0 + n] +
                         s->k1 * (p4[P4_AN_A10_B1 + n] + perpinv * p4[P4_AN_A11_B1 + n]) +
                         perpinv * (p4[P4_AN_A10_B2 + n] + perpinv * p4[P4_AN_A11_B2 + n]) +
                         s->k3 * (p4[P4_AN_A10_B3 + n] + perpinv * p4[P4_AN_A11_B3 + n]);
    }
    if (poly_order >= 3) coef_mat[n][3] = p4_combo_dot(p4, A3, fac, 4, n, z) + p4_combo_dot(p4, A2, fac, 3, n, s->b);
    if (poly_order >= 4) coef_mat[n][4] = p4_combo_dot(p4, A4, fac, 5, n, z) + p4_combo_dot(p4, A3, fac, 4, n, s->b);
  }
}
/* one element of (f_hi*M_hi + ... + f_1*M_1 + M_0): the operators of analytic_integration_with_precomp are written with the
 * HIGHEST power of perpinv first (:2558-2660) */
static double p4_combo_desc(const double *p4, const int *off, const double *fac, int nterms, int i, int j)
{
  double e = fac[nterms - 1] * p4[off[nterms - 1] + i + 4 * j];
  for (int k = nterms - 2; k >= 1; k--) e = e + fac[k] * p4[off[k] + i + 4 * j];
  return e + p4[off[0] + i + 4 * j];
}
/* analytic_integration_with_precomp (:2530-2650).  Does not touch number_of_integration_steps or the step lists (that
 * book-keeping lives in analytic_integration_without_precomp only); poly_order = 1 has no case: z is left unchanged. */
static void analytic_integration_with_precomp(poly_state *s, int poly_order, int i_precomp, double z[4], double tau)
{
  const double *p4 = p4rec(s);
  const double perpinv = s->perpinv, perpinv2 = s->perpinv2;
  const double perpinv3 = perpinv2 * perpinv, perpinv4 = perpinv2 * perpinv2;
  const double fac[5] = {1.0, perpinv, perpinv2, perpinv3, perpinv4};
  static const int M1[2] = {P4_AMAT1_0, P4_AMAT1_1}, M2[3] = {P4_AMAT2_0, P4_AMAT2_1, P4_AMAT2_2},
                   M3[4] = {P4_AMAT3_0, P4_AMAT3_1, P4_AMAT3_2, P4_AMAT3_3},
                   M4[5] = {P4_AMAT4_0, P4_AMAT4_1, P4_AMAT4_2, P4_AMAT4_3, P4_AMAT4_4};
  if (poly_order < 2 || poly_order > 4) return;
  if (i_precomp == 2 && poly_order != 2) return;
  const double tau2_half = tau * tau * 0.5, tau3_sixth = (tau * tau) * tau / 6.0;
  const double t2 = tau * tau, tau4_twentyfourth = (t2 * t2) / 24.0;
  double op_z[4][4], op_b[4][4];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      double e = tau * p4_combo_desc(p4, M1, fac, 2, i, j) + tau2_half * p4_combo_desc(p4, M2, fac, 3, i, j);
      if (poly_order >= 3) e = e + tau3_sixth * p4_combo_desc(p4, M3, fac, 4, i, j);
      if (poly_order >= 4) e = e + tau4_twentyfourth * p4_combo_desc(p4, M4, fac, 5, i, j);
      op_z[i][j] = e;
      double f = tau * (i == j ? 1.0 : 0.0) + tau2_half * p4_combo_desc(p4, M1, fac, 2, i, j);
      if (poly_order >= 3) f = f + tau3_sixth * p4_combo_desc(p4, M2, fac, 3, i, j);
      if (poly_order >= 4) f = f + tau4_twentyfourth * p4_combo_desc(p4, M3, fac, 4, i, j);
      op_b[i][j] = f;
    }
  double oz[4], ob[4];
  matvec4(oz, op_z, z);
  if (i_precomp == 1) {
    matvec4(ob, op_b, s->b);
  } else { /* operator_b_in_b (:2578-2587) */
    for (int i = 0; i < 4; i++)
      ob[i] = tau * (p4[P4_B0 + i] + s->k1 * p4[P4_B1 + i] + perpinv * p4[P4_B2 + i] + s->k3 * p4[P4_B3 + i]) +
              tau2_half * ((p4[P4_A10_B0 + i] + perpinv * p4[P4_A11_B0 + i]) +
                           s->k1 * (p4[P4_A10_B1 + i] + perpinv * p4[P4_A11_B1 + i]) +
                           perpinv * (p4[P4_A10_B2 + i] + perpinv * p4[P4_A11_B2 + i]) +
                           s->k3 * (p4[P4_A10_B3 + i] + perpinv * p4[P4_A11_B3 + i]));
  }
  for (int i = 0; i < 4; i++) z[i] = z[i] + ob[i] + oz[i];
}

/* :1258-1482.  dtau is only assigned when a valid root exists (intent(out) left untouched otherwise). */
static void analytic_approx(poly_state *s, int poly_order, const bool boole_faces[4], int i_scaling,
                            const double z[4], int *iface_inout, double *dtau, bool *boole_approx)
{
  double coef_mat[4][5];
  double dtau_vec[4] = {HUGE_D, HUGE_D, HUGE_D, HUGE_D};
  if (s->m->i_precomp == 0) analytic_coeff_without_precomp(s, poly_order, boole_faces, z, coef_mat);
  else analytic_coeff_with_precomp(s, poly_order, s->m->i_precomp, boole_faces, z, coef_mat);
  int iface = *iface_inout;
  for (int i = 1; i <= 4; i++) {
    if (!boole_faces[i - 1]) continue;
    const double *cm = coef_mat[i - 1];
    int solver = 0;
    double qa = 0, qb = 0, qc = 0, qd = 0, qe = 0; /* quart_a.. / cub_a.. / quad_a.. / lin_a.. */
    bool reduced = (i == iface) || (cm[0] == 0.0);
    switch (poly_order) {
      case 1:
        if (reduced) { dtau_vec[i - 1] = 0.0; continue; }
        solver = 1; qa = cm[1]; qb = cm[0];
        if (qa == 0.0) { dtau_vec[i - 1] = 0.0; continue; }
        break;
      case 2:
        if (reduced) {
          solver = 1; qa = cm[2] / 2.0; qb = cm[1];
          if (qa == 0.0) { dtau_vec[i - 1] = 0.0; continue; }
        } else {
          solver = 2; qa = cm[2]; qb = cm[1]; qc = cm[0];
          if (qa == 0.0) {
            if (qb != 0.0) { solver = 1; qa = qb; qb = qc; }
            else { dtau_vec[i - 1] = 0.0; continue; }
          }
        }
        break;
      case 3:
        if (reduced) {
          solver = 2; qa = cm[3] / 3.0; qb = cm[2] / 2.0; qc = cm[1];
          if (qa == 0.0) {
            if (qb != 0.0) { solver = 1; qa = qb; qb = qc; }
            else { dtau_vec[i - 1] = 0.0; continue; }
          }
        } else {
          solver = 3; qa = cm[3]; qb = cm[2]; qc = cm[1]; qd = cm[0];
          if (qa == 0.0) {
            if (qb != 0.0) { solver = 2; qa = qb; qb = qc; qc = qd; }
            else if (qc != 0.0) { solver = 1; qa = qc; qb = qd; }
            else { dtau_vec[i - 1] = 0.0; continue; }
          }
        }
        break;
      case 4:
        if (reduced) {
          solver = 3; qa = cm[4] / 4.0; qb = cm[3] / 3.0; qc = cm[2] / 2.0; qd = cm[1];
          if (qa == 0.0) {
            if (qb != 0.0) { solver = 2; qa = qb; qb = qc; qc = qd; }
            else if (qc != 0.0) { solver = 1; qa = qc; qb = qd; }
            else { dtau_vec[i - 1] = 0.0; continue; }
          }
        } else {
          solver = 4; qa = cm[4]; qb = cm[3]; qc = cm[2]; qd = cm[1]; qe = cm[0];
          if (qa == 0.0) {
            if (qb != 0.0) { solver = 3; qa = qb; qb = qc; qc = qd; qd = qe; }
            else if (qc != 0.0) { solver = 2; qa = qc; qb = qd; qc = qe; }
            else if (qd != 0.0) { solver = 1; qa = qd; qb = qe; }
            else { dtau_vec[i - 1] = 0.0; continue; }
          }
        }
        break;
    }
    switch (solver) {
      case 1: dtau_vec[i - 1] = Linear_Solver(qa, qb); break;
      case 2:
        dtau_vec[i - 1] = (i_scaling == 0) ? Quadratic_Solver1(qa, qb, qc) : Quadratic_Solver2(qa, qb, qc, s->tr);
        break;
      case 3: dtau_vec[i - 1] = Cubic_Solver(qa, qb, qc, qd, s->tr); break;
      case 4: dtau_vec[i - 1] = Quartic_Solver(i_scaling, qa, qb, qc, qd, qe, s->tr); break;
    }
  }
  int best = -1;
  for (int i = 0; i < 4; i++) {
    bool valid = (dtau_vec[i] < HUGE_D) && (dtau_vec[i] > 0.0);
    if (valid && (best < 0 || dtau_vec[i] < dtau_vec[best])) best = i;
  }
  if (best >= 0) {
    *boole_approx = true;
    *iface_inout = best + 1;
    *dtau = dtau_vec[best];
  } else {
    *boole_approx = false;
  }
}

/* :2047-2083 */
static void analytic_integration(poly_state *s, int poly_order, double z[4], double tau)
{
  if (s->m->i_precomp != 0) { /* :2034-2039 */
    analytic_integration_with_precomp(s, poly_order, s->m->i_precomp, z, tau);
    return;
  }
  s->number_of_integration_steps += 1;
  if (s->number_of_integration_steps <= s->list_cap) { /* the reference's lists hold list_cap entries */
    s->tau_steps_list[s->number_of_integration_steps - 1] = tau;
    memcpy(s->intermediate_z0_list[s->number_of_integration_steps - 1], z, 4 * sizeof(double));
  }
  if (poly_order >= 1)
    for (int i = 0; i < 4; i++) z[i] = z[i] + tau * (s->b[i] + s->amat_in_z[i]);
  if (poly_order >= 2) {
    double tau2_half = tau * tau * 0.5;
    for (int i = 0; i < 4; i++) z[i] = z[i] + tau2_half * (s->amat_in_b[i] + s->amat2_in_z[i]);
  }
  if (poly_order >= 3) {
    double tau3_sixth = (tau * tau) * tau / 6.0;
    for (int i = 0; i < 4; i++) z[i] = z[i] + tau3_sixth * (s->amat2_in_b[i] + s->amat3_in_z[i]);
  }
  if (poly_order >= 4) {
    double t2 = tau * tau;
    double tau4_twentyfourth = (t2 * t2) / 24.0;
    for (int i = 0; i < 4; i++) z[i] = z[i] + tau4_twentyfourth * (s->amat3_in_b[i] + s->amat4_in_z[i]);
  }
}
/* :2087-2113 */
static void set_integration_coef_manually(poly_state *s, int poly_order, const double z0[4])
{
  if (poly_order >= 1) matvec4(s->amat_in_z, s->amat, z0);
  if (poly_order >= 2) {
    matvec4(s->amat2_in_z, s->amat2, z0);
    matvec4(s->amat_in_b, s->amat, s->b);
  }
  if (poly_order >= 3) {
    matvec4(s->amat3_in_z, s->amat3, z0);
    matvec4(s->amat2_in_b, s->amat2, s->b);
  }
  if (poly_order >= 4) {
    matvec4(s->amat4_in_z, s->amat4, z0);
    matvec4(s->amat3_in_b, s->amat3, s->b);
  }
}

/* ------------------------------------------------------------------------------------------------
 * Hamiltonian time tracing (i_time_tracing_option = 2) and the optional quantities
 * (t_hamiltonian, gyrophase, int v_par dt, int v_par^2 dt) -- SRC/pusher_tetra_poly.f90:2117-2536,3000-3150.
 * type hamiltonian_time_type (SRC/tetra_physics_mod.f90:105-114) is built at :926-944 from fields of the
 * tetrahedron_physics record; it is re-formed here from the record with the same operations.
 * Integer powers: gfortran expands only x**2 inline (no -ffast-math in CMakeLists.txt:24); x**3..5 go through
 * libgcc __powidf2: x**3 = x*(x*x), x**4 = (x*x)*(x*x), x**5 = x*((x*x)*(x*x)).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  double h1_in_curlA, h1_in_curlh, vec_mismatch_der[3], vec_parcurr_der[3];
} ham_time;
static void hamiltonian_time_of(const double *r, ham_time *h) /* tetra_physics_mod.f90:926-944 */
{
  const double vec_h_1[3] = {r[TP_H1_1], r[TP_H2_1], r[TP_H3_1]};
  h->h1_in_curlA = dot3(vec_h_1, r + TP_CURLA);
  h->h1_in_curlh = dot3(vec_h_1, r + TP_CURLH);
  for (int i = 0; i < 3; i++) { /* matmul(mat_gh, v), mat_gh(:,j) = gh_j */
    h->vec_mismatch_der[i] = ((0.0 + r[TP_GH1 + i] * r[TP_CURLA]) + r[TP_GH2 + i] * r[TP_CURLA + 1]) + r[TP_GH3 + i] * r[TP_CURLA + 2];
    h->vec_parcurr_der[i] = ((0.0 + r[TP_GH1 + i] * r[TP_CURLH]) + r[TP_GH2 + i] * r[TP_CURLH + 1]) + r[TP_GH3 + i] * r[TP_CURLH + 2];
  }
}
/* :2436-2474 ; x_coef[i][k] = x_coef(i+1,k+1) */
static void z_series_coef(const poly_state *s, int poly_order, const double z0[4], double x_coef[3][5], double vpar_coef[5])
{
  for (int i = 0; i < 3; i++) x_coef[i][0] = z0[i];
  vpar_coef[0] = z0[3];
  if (poly_order >= 1) {
    for (int i = 0; i < 3; i++) x_coef[i][1] = s->b[i] + s->amat_in_z[i];
    vpar_coef[1] = s->b[3] + s->amat_in_z[3];
  }
  if (poly_order >= 2) {
    for (int i = 0; i < 3; i++) x_coef[i][2] = 0.5 * (s->amat_in_b[i] + s->amat2_in_z[i]);
    vpar_coef[2] = 0.5 * (s->amat_in_b[3] + s->amat2_in_z[3]);
  }
  if (poly_order >= 3) {
    for (int i = 0; i < 3; i++) x_coef[i][3] = 1.0 / 6.0 * (s->amat2_in_b[i] + s->amat3_in_z[i]);
    vpar_coef[3] = 1.0 / 6.0 * (s->amat2_in_b[3] + s->amat3_in_z[3]);
  }
  if (poly_order >= 4) {
    for (int i = 0; i < 3; i++) x_coef[i][4] = 1.0 / 24.0 * (s->amat3_in_b[i] + s->amat4_in_z[i]);
    vpar_coef[4] = 1.0 / 24.0 * (s->amat3_in_b[4 - 1] + s->amat4_in_z[3]);
  }
}
/* :2478-2520 (both operands of size n; terms of order > n-1 dropped) */
static void poly_multiplication_coef(const double *p1, const double *p2, int n, double *res)
{
  for (int i = 0; i < n; i++) res[i] = 0.0;
  for (int j = 0; j < n; j++)
    for (int k = 0; k < n; k++) {
      int cur_order = j + k;
      if (cur_order > n - 1) break;
      res[cur_order] = res[cur_order] + p1[j] * p2[k];
    }
}
/* scalar_integral_without_precomp :3063-3091 (undefined for poly_order = 0) */
static double moment_integration(int poly_order, double tau, const double *c)
{
  double r = 0.0;
  const double t2 = tau * tau;
  if (poly_order >= 1) r = c[0] * tau + t2 * 0.5 * c[1];
  if (poly_order >= 2) r = r + (tau * t2) / 3.0 * c[2];
  if (poly_order >= 3) r = r + (t2 * t2) / 4.0 * c[3];
  if (poly_order >= 4) r = r + (tau * (t2 * t2)) / 5.0 * c[4];
  return r;
}
/* :2214-2256 */
static void calc_t_hamiltonian(poly_state *s, int poly_order, const double z0[4], double tau, double *t_hamiltonian)
{
  const double cm_over_e = s->m->cm_over_e;
  const int n = poly_order + 1;
  double x_coef[3][5], vpar_coef[5], x_vpar_coef[3][5], mi_x[3], mi_xv[3];
  ham_time h;
  hamiltonian_time_of(s->r, &h);
  if (s->number_of_integration_steps > 1) set_integration_coef_manually(s, poly_order, z0);
  z_series_coef(s, poly_order, z0, x_coef, vpar_coef);
  for (int i = 0; i < 3; i++) poly_multiplication_coef(x_coef[i], vpar_coef, n, x_vpar_coef[i]);
  for (int i = 0; i < 3; i++) {
    mi_x[i] = moment_integration(poly_order, tau, x_coef[i]);
    mi_xv[i] = moment_integration(poly_order, tau, x_vpar_coef[i]);
  }
  double delta = h.h1_in_curlA * tau + cm_over_e * h.h1_in_curlh * moment_integration(poly_order, tau, vpar_coef) +
                 dot3(h.vec_mismatch_der, mi_x) + cm_over_e * dot3(h.vec_parcurr_der, mi_xv);
  delta = delta * (double)s->sign_rhs;
  *t_hamiltonian = *t_hamiltonian + delta;
}
/* :2340-2432 */
static void get_t_hamiltonian_root(poly_state *s, int poly_order, const double z0[4], double t_remain, double *tau_root)
{
  const double cm_over_e = s->m->cm_over_e;
  const int n = poly_order + 1;
  double x_coef[3][5], vpar_coef[5], x_vpar_coef[3][5], col[3], colv[3];
  double b_coef = 0.0, c_coef = 0.0, d_coef = 0.0, e_coef;
  ham_time h;
  hamiltonian_time_of(s->r, &h);
  z_series_coef(s, poly_order, z0, x_coef, vpar_coef);
  for (int i = 0; i < 3; i++) poly_multiplication_coef(x_coef[i], vpar_coef, n, x_vpar_coef[i]);
#define COLS(k) for (int i = 0; i < 3; i++) { col[i] = x_coef[i][k]; colv[i] = x_vpar_coef[i][k]; }
  COLS(0)
  e_coef = h.h1_in_curlA + vpar_coef[0] * cm_over_e * h.h1_in_curlh + dot3(col, h.vec_mismatch_der) +
           cm_over_e * dot3(colv, h.vec_parcurr_der);
  if (poly_order >= 1) {
    COLS(1)
    d_coef = 0.5 * (vpar_coef[1] * cm_over_e * h.h1_in_curlh + dot3(col, h.vec_mismatch_der) +
                    cm_over_e * dot3(colv, h.vec_parcurr_der));
  }
  if (poly_order >= 2) {
    COLS(2)
    c_coef = (1.0 / 3.0) * (vpar_coef[2] * cm_over_e * h.h1_in_curlh + dot3(col, h.vec_mismatch_der) +
                            cm_over_e * dot3(colv, h.vec_parcurr_der));
  }
  if (poly_order >= 3) {
    COLS(3)
    b_coef = (1.0 / 4.0) * (vpar_coef[3] * cm_over_e * h.h1_in_curlh + dot3(col, h.vec_mismatch_der) +
                            cm_over_e * dot3(colv, h.vec_parcurr_der));
  }
#undef COLS
  if (poly_order >= 1) d_coef = 2.0 * d_coef;
  if (poly_order >= 2) c_coef = 6.0 * c_coef;
  if (poly_order >= 3) b_coef = 24.0 * b_coef;
  e_coef = e_coef * (double)s->sign_rhs;
  if (poly_order >= 1) d_coef = d_coef * (double)s->sign_rhs;
  if (poly_order >= 2) c_coef = c_coef * (double)s->sign_rhs;
  if (poly_order >= 3) b_coef = b_coef * (double)s->sign_rhs;
  switch (poly_order) {
    case 1: *tau_root = Quadratic_Solver2(d_coef, e_coef, -t_remain, s->tr); break;
    case 2: *tau_root = Cubic_Solver(c_coef, d_coef, e_coef, -t_remain, s->tr); break;
    default: *tau_root = Quartic_Solver(0, b_coef, c_coef, d_coef, e_coef, -t_remain, s->tr); break;
  }
}
/* :2117-2130, 2134-2210 ; optq = {t_hamiltonian, gyrophase, vpar_int, vpar2_int} */
static void calc_optional_quantities(poly_state *s, int poly_order, const double z0[4], double tau, double optq[4])
{
  const gor_mesh *m = s->m;
  const double cm_over_e = m->cm_over_e;
  const int n = poly_order + 1;
  double x_coef[3][5], vpar_coef[5], x_vpar_coef[3][5], dtdtau_coef[5], col[3], colv[3], prod[5], prod2[5];
  ham_time h;
  hamiltonian_time_of(s->r, &h);
  if (s->number_of_integration_steps > 1) set_integration_coef_manually(s, poly_order, z0);
  z_series_coef(s, poly_order, z0, x_coef, vpar_coef);
  for (int i = 0; i < 3; i++) poly_multiplication_coef(vpar_coef, x_coef[i], n, x_vpar_coef[i]); /* poly_multiplication(vpar_coef,x_coef) */
  for (int k = 0; k < n; k++) {
    for (int i = 0; i < 3; i++) { col[i] = x_coef[i][k]; colv[i] = x_vpar_coef[i][k]; }
    dtdtau_coef[k] = dot3(h.vec_mismatch_der, col) + cm_over_e * h.h1_in_curlh * vpar_coef[k] +
                     cm_over_e * dot3(h.vec_parcurr_der, colv);
  }
  dtdtau_coef[0] = dtdtau_coef[0] + h.h1_in_curlA;
  if (m->boole_time_hamiltonian)
    optq[0] = optq[0] + moment_integration(poly_order, tau, dtdtau_coef) * (double)s->sign_rhs;
  if (m->boole_gyrophase) {
    double omega_coef[5];
    for (int k = 0; k < n; k++) {
      for (int i = 0; i < 3; i++) col[i] = x_coef[i][k];
      omega_coef[k] = 1.0 / cm_over_e * dot3(s->r + TP_GB, col);
    }
    omega_coef[0] = omega_coef[0] + 1.0 / cm_over_e * s->r[TP_BMOD1];
    poly_multiplication_coef(dtdtau_coef, omega_coef, n, prod);
    optq[1] = optq[1] - (double)s->sign_rhs * moment_integration(poly_order, tau, prod);
  }
  if (m->boole_vpar_int) {
    poly_multiplication_coef(dtdtau_coef, vpar_coef, n, prod);
    optq[2] = optq[2] + (double)s->sign_rhs * moment_integration(poly_order, tau, prod);
  }
  if (m->boole_vpar2_int) {
    poly_multiplication_coef(vpar_coef, vpar_coef, n, prod2);
    poly_multiplication_coef(dtdtau_coef, prod2, n, prod);
    optq[3] = optq[3] + (double)s->sign_rhs * moment_integration(poly_order, tau, prod);
  }
}

/* :2690-2705 */
static double normal_distance_func(const poly_state *s, const double z123[3], int iface)
{
  double dist1 = -s->r[TP_DIST_REF];
  double d = dot3(z123, anorm_col(s, iface));
  if (iface == 1) d = d - dist1;
  return d;
}
/* :2709-2739 (i_precomp = 0 branch; poly1 quantities of tetra_physics_poly_precomp_mod.f90:103-156
 * formed on the fly with the same matmul(n_vec, mat) accumulation order) */
static double normal_velocity_func(const poly_state *s, const double z[4], int iface)
{
  const gor_mesh *m = s->m;
  const double *r = s->r, *n = anorm_col(s, iface);
  if (m->i_precomp != 0) { /* :2734-2738; b is the module variable: set by i_precomp = 1, never by i_precomp = 2 (zero) */
    const double *p4 = p4rec(s);
    double sacc = 0.0;
    for (int i = 0; i < 4; i++)
      sacc = sacc + (p4[P4_AN_AMAT1_0 + i + 4 * (iface - 1)] + s->perpinv * p4[P4_AN_AMAT1_1 + i + 4 * (iface - 1)]) * z[i];
    return sacc * (double)s->sign_rhs + dot3(n, s->b);
  }
  double in_alp[3], in_bet[3], in_gam[3];
  for (int j = 0; j < 3; j++) {
    in_alp[j] = dot3(n, r + TP_ALPMAT + 3 * j); /* sum_i n(i)*alpmat(i,j) */
    in_bet[j] = dot3(n, r + TP_BETMAT + 3 * j);
  }
  double in_betvec = dot3(n, r + TP_CURLA);
  double t[3];
  for (int j = 0; j < 3; j++) t[j] = (-CLIGHT * in_bet[j] + s->perpinv * m->cm_over_e * in_alp[j]) * z[j];
  double v = (((0.0 + t[0]) + t[1]) + t[2] + in_betvec * z[3]) * (double)s->sign_rhs + dot3(n, s->b);
  if (m->boole_strong_electric_field) {
    for (int j = 0; j < 3; j++) in_gam[j] = dot3(n, r + TP_GAMMAT + 3 * j);
    double in_gamvec = dot3(n, r + TP_CURLVE);
    for (int j = 0; j < 3; j++) t[j] = -0.5 * m->cm_over_e * in_gam[j] * z[j];
    v = v + (((0.0 + t[0]) + t[1]) + t[2] + m->cm_over_e * in_gamvec * z[3]) * (double)s->sign_rhs;
  }
  return v;
}
/* :2741-2775 */
static double normal_v_func_from_trajectory(const poly_state *s, int poly_order, int iface, double tau)
{
  const double *n = anorm_col(s, iface);
  double v = 0.0, t[3];
  if (poly_order >= 1) {
    for (int i = 0; i < 3; i++) t[i] = n[i] * (s->b[i] + s->amat_in_z[i]);
    v = ((0.0 + t[0]) + t[1]) + t[2];
  }
  if (poly_order >= 2) {
    for (int i = 0; i < 3; i++) t[i] = n[i] * tau * (s->amat_in_b[i] + s->amat2_in_z[i]);
    v = v + (((0.0 + t[0]) + t[1]) + t[2]);
  }
  if (poly_order >= 3) {
    double tau2_half = tau * tau * 0.5;
    for (int i = 0; i < 3; i++) t[i] = n[i] * tau2_half * (s->amat2_in_b[i] + s->amat3_in_z[i]);
    v = v + (((0.0 + t[0]) + t[1]) + t[2]);
  }
  if (poly_order >= 4) {
    double tau3_sixth = (tau * tau) * tau / 6.0;
    for (int i = 0; i < 3; i++) t[i] = n[i] * tau3_sixth * (s->amat3_in_b[i] + s->amat4_in_z[i]);
    v = v + (((0.0 + t[0]) + t[1]) + t[2]);
  }
  return v;
}
/* :679-758 */
static void check_three_planes(const poly_state *s, const double z[4], int iface_new, bool *ok)
{
  for (int j = 1; j <= 3; j++) {
    int k = ((iface_new + j - 1) % 4) + 1;
    if (normal_distance_func(s, z, k) < 0.0) *ok = false;
  }
}
static void check_face_convergence(const poly_state *s, const double z[4], int iface_new, bool *ok)
{
  if (fabs(normal_distance_func(s, z, iface_new)) > 1.e-11) *ok = false;
}
static void check_velocity(const poly_state *s, const double z[4], int iface_new, bool *ok)
{
  if (normal_velocity_func(s, z, iface_new) > 0.0) *ok = false;
}
static void check_exit_time(double tau, double tau_max, bool *ok, int poly_order)
{
  if (poly_order > 2)
    if (tau > tau_max) *ok = false;
}
/* :2779-2831 */
static double physical_estimate_tau(const poly_state *s)
{
  const gor_mesh *m = s->m;
  const double *r = s->r;
  const double eps_modulation = 0.1;
  double tetra_dist_ref = fabs(r[TP_TETRA_DIST_REF]);
  double vperp2 = -2.0 * s->perpinv * s->bmod0;
  double vd_ExB;
  if (m->boole_strong_electric_field)
    vd_ExB = r[TP_VE_MOD_AVG];
  else
    vd_ExB = fabs(CLIGHT / s->bmod0 * r[TP_ER_MOD]);
  bool boole_vd_ExB = !(vd_ExB == 0.0), boole_vperp = !(vperp2 == 0.0);
  double c1 = fabs(tetra_dist_ref / s->z_init[3]);
  double tau_est;
  if (boole_vperp) {
    double c2 = sqrt(tetra_dist_ref * s->vmod0 * r[TP_R1] / (vperp2 * m->grid_size[1] * eps_modulation));
    tau_est = c1;
    if (c2 < tau_est) tau_est = c2;
    if (boole_vd_ExB) {
      double c3 = tetra_dist_ref / vd_ExB;
      if (c3 < tau_est) tau_est = c3;
    }
  } else if (boole_vd_ExB) {
    double c3 = tetra_dist_ref / vd_ExB;
    tau_est = c1;
    if (c3 < tau_est) tau_est = c3;
  } else {
    tau_est = c1;
  }
  return fabs(tau_est / s->dt_dtau_const);
}


/* ------------------------------------------------------------------------------------------------
 * Adaptive energy-controlled sub-stepping (boole_adaptive_time_steps) -- SRC/pusher_tetra_poly.f90:830-1254.
 * Where the reference reads uninitialised locals the restatement defines: eta_minimum / tau_minimum start as the unsplit
 * step (1, tau), and iface_out_adaptive keeps the incoming face when no exit-time solve was made.
 * ---------------------------------------------------------------------------------------------- */
static void analytic_approx(poly_state *s, int poly_order, const bool boole_faces[4], int i_scaling,
                            const double z[4], int *iface_inout, double *dtau, bool *boole_approx);
/* :1167-1213 */
static void adaptive_time_steps_update_eta(const gor_mesh *m, int poly_order, double delta_energy_current, int *eta)
{
  const double threshold = 1.0, min_step_error = (double)1E-15f, additive_increase = 1.0; /* 1E-15 is a default-real literal */
  double scale_factor = pow(delta_energy_current / m->desired_delta_energy, 1.0 / poly_order);
  double max_scale_factor = pow(delta_energy_current / (min_step_error * *eta), 1.0 / (poly_order + 1));
  if ((scale_factor > threshold) && (max_scale_factor > threshold)) {
    scale_factor = scale_factor < max_scale_factor ? scale_factor : max_scale_factor;
    int c = (int)ceil(*eta * scale_factor);
    *eta = c < m->max_n_intermediate_steps ? c : m->max_n_intermediate_steps;
  } else if (scale_factor > threshold) {
    int c = (int)ceil(*eta * scale_factor);
    *eta = c < m->max_n_intermediate_steps ? c : m->max_n_intermediate_steps;
  } else {
    *eta = (int)(*eta + additive_increase);
  }
}
/* :1217-1252 */
static void adaptive_time_steps_exit_time(poly_state *s, int poly_order, int i_scaling, const double z[4],
                                          bool boole_guess_adaptive, int *iface_new_adaptive, double *tau_exit,
                                          bool *boole_analytical_approx)
{
  bool boole_faces[4] = {true, true, true, true};
  if (boole_guess_adaptive && (poly_order > 2)) {
    analytic_approx(s, 2, boole_faces, i_scaling, z, iface_new_adaptive, tau_exit, boole_analytical_approx);
    if (*boole_analytical_approx) {
      for (int i = 0; i < 4; i++) boole_faces[i] = false;
      boole_faces[*iface_new_adaptive - 1] = true;
    }
    *iface_new_adaptive = 0;
  }
  analytic_approx(s, poly_order, boole_faces, i_scaling, z, iface_new_adaptive, tau_exit, boole_analytical_approx);
}
/* :937-1163 */
static void adaptive_time_steps_equidistant(poly_state *s, int poly_order, int i_scaling, bool boole_guess_adaptive,
                                            bool boole_passing, double *delta_energy_current, int *iface_out_adaptive,
                                            double *tau, double z[4], bool *boole_face_correct)
{
  const gor_mesh *m = s->m;
  const int max_n = m->max_n_intermediate_steps;
  double z_start_adaptive[4];
  memcpy(z_start_adaptive, s->intermediate_z0_list[s->number_of_integration_steps - 1], sizeof(z_start_adaptive));
  const int number_of_integration_steps_start_adaptive = s->number_of_integration_steps - 1;
  const double energy_start_adaptive = gor_energy_tot(m, z_start_adaptive, s->perpinv, s->ind_tetr);
  int eta = 1, eta_extended = 1, eta_limit, eta_minimum = 1, eta_buffer;
  bool boole_reached_minimum = false, boole_energy_check = false, boole_exit_tetrahedron, boole_analytical_approx;
  double delta_energy_minimum = *delta_energy_current;
  double tau_prime, tau_collected, tau_exit = 0.0, tau_minimum = *tau, tau_buffer, energy_current;
  int iface_new_adaptive = *iface_out_adaptive;
  /* Safety: the reference's loop has no iteration bound; when two partitions give EXACTLY the same energy error (it is a
   * multiple of 2^-53) neither branch below fires, eta is reset to the steps actually taken and the loop can cycle for
   * ever.  The restatement ends the loop after max_n + 64 partitions, keeping the last one. */
  int n_partitions = 0;
  while (eta < max_n) { /* PARTITION */
    if (++n_partitions > max_n + 64) break;
    adaptive_time_steps_update_eta(m, poly_order, *delta_energy_current, &eta);
    if (boole_reached_minimum) {
      eta = eta_minimum;
      *tau = tau_minimum;
    }
    memcpy(z, z_start_adaptive, 4 * sizeof(double));
    s->number_of_integration_steps = number_of_integration_steps_start_adaptive;
    tau_prime = *tau / eta;
    tau_collected = 0;
    *boole_face_correct = true;
    boole_exit_tetrahedron = false;
    boole_energy_check = false;
    if (boole_passing)
      eta_limit = (int)ceil(max_n * 1.1);
    else
      eta_limit = eta;
    for (int i = 1; i <= eta_limit - 1; i++) { /* STEPWISE */
      set_integration_coef_manually(s, poly_order, z);
      analytic_integration(s, poly_order, z, tau_prime);
      bool left = false;
      for (int k = 1; k <= 4; k++)
        if (normal_distance_func(s, z, k) < 0.0) {
          left = true;
          break;
        }
      if (left) {
        if (i == 1) {
          *boole_face_correct = false;
          return;
        }
        memcpy(z, s->intermediate_z0_list[s->number_of_integration_steps - 1], 4 * sizeof(double));
        s->number_of_integration_steps = s->number_of_integration_steps - 1;
        boole_exit_tetrahedron = true;
        break;
      }
      tau_collected = tau_collected + tau_prime;
      eta_extended = i;
      if (!boole_guess_adaptive) {
        iface_new_adaptive = 0;
        adaptive_time_steps_exit_time(s, poly_order, i_scaling, z, false, &iface_new_adaptive, &tau_exit,
                                      &boole_analytical_approx);
        if (!boole_analytical_approx) {
          *boole_face_correct = false;
          return;
        }
        if (tau_exit <= tau_prime) {
          tau_prime = tau_exit;
          break;
        }
      }
    }
    if (boole_exit_tetrahedron) {
      iface_new_adaptive = 0;
      adaptive_time_steps_exit_time(s, poly_order, i_scaling, z, boole_guess_adaptive, &iface_new_adaptive, &tau_exit,
                                    &boole_analytical_approx);
      if (!boole_analytical_approx) {
        *boole_face_correct = false;
        return;
      }
      tau_prime = tau_exit;
    } else {
      set_integration_coef_manually(s, poly_order, z);
    }
    analytic_integration(s, poly_order, z, tau_prime);
    tau_collected = tau_collected + tau_prime;
    energy_current = gor_energy_tot(m, z, s->perpinv, s->ind_tetr);
    *delta_energy_current = fabs(1 - energy_current / energy_start_adaptive);
    eta_buffer = eta;
    tau_buffer = *tau;
    eta = eta_extended + 1;
    *tau = tau_collected;
    if (boole_reached_minimum) {
      break;
    } else if (*delta_energy_current < delta_energy_minimum) {
      delta_energy_minimum = *delta_energy_current;
      eta_minimum = eta_buffer;
      tau_minimum = tau_buffer;
      if (*delta_energy_current <= m->desired_delta_energy) {
        boole_energy_check = true;
        break;
      }
    } else if (*delta_energy_current > delta_energy_minimum) {
      boole_reached_minimum = true;
      continue;
    }
  }
  *iface_out_adaptive = iface_new_adaptive;
  (void)boole_energy_check; /* the reference only prints a message when the energy goal was not met */
  if (s->tr) s->tr->n_adaptive++;
}
/* :830-933 */
static void overhead_adaptive_time_steps(poly_state *s, int poly_order, int i_scaling, bool boole_guess_adaptive,
                                         bool boole_passing, int *iface_inout_adaptive, double *tau, double z[4],
                                         bool *boole_face_correct)
{
  if (boole_passing) {
    check_three_planes(s, z, *iface_inout_adaptive, boole_face_correct);
    check_face_convergence(s, z, *iface_inout_adaptive, boole_face_correct);
  } else {
    check_three_planes(s, z, 0, boole_face_correct);
  }
  if (!*boole_face_correct) return;
  const double *z0 = s->intermediate_z0_list[s->number_of_integration_steps - 1];
  double energy_start = gor_energy_tot(s->m, z0, s->perpinv, s->ind_tetr);
  double energy_current = gor_energy_tot(s->m, z, s->perpinv, s->ind_tetr);
  double delta_energy_current = fabs(1 - energy_current / energy_start);
  if (delta_energy_current > s->m->desired_delta_energy)
    adaptive_time_steps_equidistant(s, poly_order, i_scaling, boole_guess_adaptive, boole_passing, &delta_energy_current,
                                    iface_inout_adaptive, tau, z, boole_face_correct);
}

/* :762-826 ; returns false when the particle has to be removed (ind_tetr=-1, iface=-1) */
static void prolonged_trajectory(poly_state *s, int poly_order, int i_scaling, double z[4], double *tau,
                                 int *iface_new, bool *boole_face_correct, bool *boole_analytical_approx)
{
  bool boole_faces[4] = {true, true, true, true};
  double tau_save = *tau, tau_max = 0.0;
  int iface_new_save = *iface_new;
  if (poly_order > 2) {
    analytic_approx(s, 2, boole_faces, i_scaling, z, iface_new, tau, boole_analytical_approx);
    tau_max = *tau * eps_tau;
  }
  *iface_new = iface_new_save;
  analytic_approx(s, poly_order, boole_faces, i_scaling, z, iface_new, tau, boole_analytical_approx);
  if (!*boole_analytical_approx) return;
  analytic_integration(s, poly_order, z, *tau);
  if (s->m->boole_adaptive_time_steps) /* :811-814 */
    overhead_adaptive_time_steps(s, poly_order, i_scaling, false, true, iface_new, tau, z, boole_face_correct);
  check_exit_time(*tau, tau_max, boole_face_correct, poly_order);
  check_three_planes(s, z, *iface_new, boole_face_correct);
  check_velocity(s, z, *iface_new, boole_face_correct);
  check_face_convergence(s, z, *iface_new, boole_face_correct);
  *tau = *tau + tau_save;
  if (s->tr) s->tr->n_fallback[2]++;
}

/* :2835-2998 */
static void trouble_shooting_polynomial_solver(poly_state *s, int poly_order, double z[4], double *tau,
                                               int *iface_new, bool *boole_trouble_shooting)
{
  bool boole_faces[4] = {true, true, true, true};
  bool boole_analytical_approx, boole_face_correct;
  int i_scaling = 0, poly_order_new = poly_order;
  double tau_max, tau_max_est;
  if (s->tr) s->tr->n_fallback[1]++;
  *boole_trouble_shooting = true;
  *iface_new = s->iface_init;
  memcpy(z, s->z_init, 4 * sizeof(double));
  analytic_approx(s, 2, boole_faces, i_scaling, z, iface_new, tau, &boole_analytical_approx);
  tau_max = *tau * eps_tau;
  boole_face_correct = false;
  int i = 0;
  while ((!boole_face_correct) && (poly_order == 4)) {
    i = i + 1;
    i_scaling = i;
    *iface_new = s->iface_init;
    memcpy(z, s->z_init, 4 * sizeof(double));
    s->number_of_integration_steps = 0;
    analytic_approx(s, poly_order, boole_faces, i_scaling, z, iface_new, tau, &boole_analytical_approx);
    if (!boole_analytical_approx) {
      *boole_trouble_shooting = false;
      return;
    }
    analytic_integration(s, poly_order, z, *tau);
    /* :2897-2902 -- called BEFORE boole_face_correct is reset, i.e. with .false.: it returns at once (:893) */
    if (s->m->boole_adaptive_time_steps)
      overhead_adaptive_time_steps(s, poly_order, i_scaling, false, true, iface_new, tau, z, &boole_face_correct);
    boole_face_correct = true;
    check_three_planes(s, z, *iface_new, &boole_face_correct);
    check_face_convergence(s, z, *iface_new, &boole_face_correct);
    check_velocity(s, z, *iface_new, &boole_face_correct);
    if (*tau > tau_max) {
      tau_max_est = physical_estimate_tau(s);
      tau_max_est = tau_max_est * eps_tau;
      if (*tau > tau_max_est) boole_face_correct = false;
    }
    if (i == 6) break;
  }
  if (!boole_face_correct) {
    *iface_new = s->iface_init;
    memcpy(z, s->z_init, 4 * sizeof(double));
    s->number_of_integration_steps = 0;
    switch (poly_order) {
      case 2: poly_order_new = 2; i_scaling = 1; break;
      case 3: poly_order_new = 3; i_scaling = 1; break;
      case 4: poly_order_new = 3; i_scaling = 0; break;
      default: /* no case(1) in the reference (:2936-2947): poly_order_new undefined; keep order, i_scaling */
        poly_order_new = poly_order;
        break;
    }
    analytic_approx(s, poly_order_new, boole_faces, i_scaling, z, iface_new, tau, &boole_analytical_approx);
    if (!boole_analytical_approx) {
      *boole_trouble_shooting = false;
      return;
    }
    analytic_integration(s, poly_order, z, *tau); /* original order, :2961 */
    boole_face_correct = true;
    if (s->m->boole_adaptive_time_steps) /* :2966-2971, with the REDUCED order */
      overhead_adaptive_time_steps(s, poly_order_new, i_scaling, false, true, iface_new, tau, z, &boole_face_correct);
    check_three_planes(s, z, *iface_new, &boole_face_correct);
    check_face_convergence(s, z, *iface_new, &boole_face_correct);
    check_velocity(s, z, *iface_new, &boole_face_correct);
    if (*tau > tau_max) {
      tau_max_est = physical_estimate_tau(s);
      tau_max_est = tau_max_est * eps_tau;
      if (*tau > tau_max_est) boole_face_correct = false;
    }
    if (!boole_face_correct) {
      *boole_trouble_shooting = false;
      return;
    }
  }
}

/* :182-675 */
static void pusher_tetra_poly(poly_state *s, int poly_order, int *ind_tetr_inout, int *iface, double x[3],
                              double *vpar, double z_save[3], double t_remain_in, double *t_pass,
                              bool *boole_t_finished, int *iper_phi, double optq[4] /* intent(out), may be NULL */)
{
  const gor_mesh *m = s->m;
  bool boole_faces[4] = {true, true, true, true};
  if (optq) optq[0] = optq[1] = optq[2] = optq[3] = 0.0; /* initialise_optional_quantities :2117-2130 */
  bool boole_analytical_approx = false, boole_face_correct, boole_trouble_shooting = true;
  double z[4], tau = 0.0, tau_max;
  int iface_new, i_scaling = 0;

  initialize_pusher_tetra_poly(s, *ind_tetr_inout, x, *iface, *vpar, t_remain_in);
  s->removed = false;
  s->number_of_integration_steps = 0;
  memcpy(z, s->z_init, sizeof(z));
  *iper_phi = 0;
  *boole_t_finished = false;
  *t_pass = 0.0; /* undefined in the reference on the "remove particle" returns */
  iface_new = s->iface_init;

  /* ---- first attempt with second order guess (:267-346) */
  analytic_approx(s, 2, boole_faces, i_scaling, z, &iface_new, &tau, &boole_analytical_approx);
  tau_max = tau * eps_tau;
  if (m->boole_guess && boole_analytical_approx && (poly_order > 2)) {
    for (int i = 0; i < 4; i++) boole_faces[i] = false;
    boole_faces[iface_new - 1] = true;
  }
  if (poly_order > 2) {
    iface_new = s->iface_init;
    analytic_approx(s, poly_order, boole_faces, i_scaling, z, &iface_new, &tau, &boole_analytical_approx);
  }
  boole_face_correct = true;
  if (!boole_analytical_approx) boole_face_correct = false;
  if (boole_face_correct) {
    analytic_integration(s, poly_order, z, tau);
    if (m->boole_adaptive_time_steps) /* :316-319 */
      overhead_adaptive_time_steps(s, poly_order, i_scaling, m->boole_guess != 0, true, &iface_new, &tau, z, &boole_face_correct);
    check_three_planes(s, z, iface_new, &boole_face_correct);
    check_face_convergence(s, z, iface_new, &boole_face_correct);
    check_exit_time(tau, tau_max, &boole_face_correct, poly_order);
    if (boole_face_correct) {
      double nv = (m->i_precomp == 0) ? normal_v_func_from_trajectory(s, poly_order, iface_new, tau)
                                      : normal_velocity_func(s, z, iface_new); /* :328-332 */
      if (nv > 0.0) {
        if (poly_order > 2) {
          boole_face_correct = false;
        } else {
          prolonged_trajectory(s, poly_order, i_scaling, z, &tau, &iface_new, &boole_face_correct,
                               &boole_analytical_approx);
          if (!boole_analytical_approx) {
            *ind_tetr_inout = -1;
            *iface = -1;
        s->removed = true;
            return;
          }
        }
      }
    }
  }
  /* ---- second attempt without guess (+ rescaling in 2nd order) (:361-418) */
  if (!boole_face_correct) {
    if (s->tr) s->tr->n_fallback[0]++;
    for (int i = 0; i < 4; i++) boole_faces[i] = true;
    boole_face_correct = true;
    iface_new = s->iface_init;
    memcpy(z, s->z_init, sizeof(z));
    s->number_of_integration_steps = 0;
    if (poly_order == 2)
      analytic_approx(s, poly_order, boole_faces, 1, z, &iface_new, &tau, &boole_analytical_approx);
    else
      analytic_approx(s, poly_order, boole_faces, i_scaling, z, &iface_new, &tau, &boole_analytical_approx);
    if (!boole_analytical_approx) {
      *ind_tetr_inout = -1;
      *iface = -1;
        s->removed = true;
      return;
    }
    analytic_integration(s, poly_order, z, tau);
    if (m->boole_adaptive_time_steps) /* :391-399 */
      overhead_adaptive_time_steps(s, poly_order, poly_order == 2 ? 1 : i_scaling, false, true, &iface_new, &tau, z,
                                   &boole_face_correct);
    check_exit_time(tau, tau_max, &boole_face_correct, poly_order);
    check_three_planes(s, z, iface_new, &boole_face_correct);
    check_face_convergence(s, z, iface_new, &boole_face_correct);
    if (boole_face_correct) {
      if (normal_velocity_func(s, z, iface_new) > 0.0) {
        prolonged_trajectory(s, poly_order, i_scaling, z, &tau, &iface_new, &boole_face_correct,
                             &boole_analytical_approx);
        if (!boole_analytical_approx) {
          *ind_tetr_inout = -1;
          *iface = -1;
        s->removed = true;
          return;
        }
      }
    }
    /* ---- third attempt: trouble shooting (:429-441) */
    if (!boole_face_correct) {
      trouble_shooting_polynomial_solver(s, poly_order, z, &tau, &iface_new, &boole_trouble_shooting);
      if (!boole_trouble_shooting) {
        *ind_tetr_inout = -1;
        *iface = -1;
        s->removed = true;
        return;
      }
    }
  }
  /* ---- final processing (:459-487) */
  const bool tt2 = (m->i_time_tracing_option == 2);
  double thl_static[3] = {0.0, 0.0, 0.0};
  double *t_hamiltonian_list = s->thl_heap ? s->thl_heap : thl_static; /* allocate(t_hamiltonian_list(number_of_integration_steps+1)) :471 */
  t_hamiltonian_list[0] = 0.0;
  for (int i = 0; i < 3; i++) x[i] = z[i] + s->r[TP_X1 + i];
  *vpar = z[3];
  if (!tt2) {
    *t_pass = tau * s->dt_dtau_const;
  } else { /* Hamiltonian time tracing with computation of polynomial (:468-486) */
    double t_hamiltonian = 0.0;
    const int nst = s->number_of_integration_steps;
    for (int i = 1; i <= nst; i++) {
      calc_t_hamiltonian(s, poly_order, s->intermediate_z0_list[i - 1], s->tau_steps_list[i - 1], &t_hamiltonian);
      t_hamiltonian_list[i] = t_hamiltonian;
    }
    *t_pass = t_hamiltonian;
  }

  if (fabs(*t_pass) >= fabs(s->t_remain)) {
    /* ---- fourth attempt: particle stops inside the cell (:497-645) */
    if (!tt2) {
      memcpy(z, s->z_init, sizeof(z));
      if (s->number_of_integration_steps > 1) {
        iface_new = s->iface_init;
        set_integration_coef_manually(s, poly_order, z);
      }
      s->number_of_integration_steps = 0;
      tau = s->t_remain / s->dt_dtau_const;
    } else { /* :523-541 */
      /* findloc(abs(t_hamiltonian_list) > abs(t_remain)); |t_pass| == |t_remain| exactly finds nothing in the
       * reference (index 0, out of bounds); the last step is taken here */
      int i_step_root = 0;
      for (int i = 1; i <= s->number_of_integration_steps + 1; i++)
        if (fabs(t_hamiltonian_list[i - 1]) > fabs(s->t_remain)) {
          i_step_root = i;
          break;
        }
      if (i_step_root == 0) i_step_root = s->number_of_integration_steps + 1;
      memcpy(z, s->intermediate_z0_list[i_step_root - 2], sizeof(z));
      set_integration_coef_manually(s, poly_order, z);
      s->number_of_integration_steps = i_step_root - 2;
      iface_new = s->iface_init;
      double t_remain_new = s->t_remain - t_hamiltonian_list[i_step_root - 2];
      get_t_hamiltonian_root(s, poly_order, z, t_remain_new, &tau);
      if (tau > s->tau_steps_list[i_step_root - 2]) tau = t_remain_new / s->dt_dtau_const;
    }
    analytic_integration(s, poly_order, z, tau);
    if (m->boole_adaptive_time_steps) /* :564-568 */
      overhead_adaptive_time_steps(s, poly_order, i_scaling, false, false, &iface_new, &tau, z, &boole_face_correct);
    *ind_tetr_inout = s->ind_tetr;
    *iface = 0;
    boole_face_correct = true;
    for (int i = 1; i <= 4; i++)
      if (normal_distance_func(s, z, i) < 0.0) boole_face_correct = false;
    if (boole_face_correct) {
      *boole_t_finished = true;
      for (int i = 0; i < 3; i++) z_save[i] = z[i];
      for (int i = 0; i < 3; i++) x[i] = z[i] + s->r[TP_X1 + i];
      *vpar = z[3];
      *t_pass = s->t_remain;
    } else {
      if (s->tr) s->tr->n_fallback[3]++;
      trouble_shooting_polynomial_solver(s, poly_order, z, &tau, &iface_new, &boole_trouble_shooting);
      if (!boole_trouble_shooting) {
        *ind_tetr_inout = -1;
        *iface = -1;
        s->removed = true;
        return;
      }
      for (int i = 0; i < 3; i++) x[i] = z[i] + s->r[TP_X1 + i];
      *vpar = z[3];
      if (!tt2) {
        *t_pass = tau * s->dt_dtau_const;
      } else { /* :621-631 */
        double t_hamiltonian = 0.0;
        const int nst = s->number_of_integration_steps;
        for (int i = 1; i <= nst; i++)
          calc_t_hamiltonian(s, poly_order, s->intermediate_z0_list[i - 1], s->tau_steps_list[i - 1], &t_hamiltonian);
        *t_pass = t_hamiltonian;
      }
      for (int i = 0; i < 3; i++) z_save[i] = z[i];
      *iface = iface_new;
      handover2neighbour(m, s->ind_tetr, ind_tetr_inout, iface, x, iper_phi);
    }
  } else {
    for (int i = 0; i < 3; i++) z_save[i] = z[i];
    *iface = iface_new;
    handover2neighbour(m, s->ind_tetr, ind_tetr_inout, iface, x, iper_phi);
  }
  /* optional quantities of this push (:662-667), summed over the pushes of the time step by the caller */
  if (optq && (m->boole_time_hamiltonian || m->boole_gyrophase || m->boole_vpar_int || m->boole_vpar2_int)) {
    const int nst = s->number_of_integration_steps;
    for (int i = 1; i <= nst; i++)
      calc_optional_quantities(s, poly_order, s->intermediate_z0_list[i - 1], s->tau_steps_list[i - 1], optq);
  }
}


/* ------------------------------------------------------------------------------------------------
 * Orbit events: parallel adiabatic invariant J_par / banana tips (v_par = 0) and toroidal (phi = 0) mappings.
 * module par_adiab_inv_poly_mod, SRC/pusher_tetra_poly.f90:3156-3429, and the event part of
 * gorilla_plot_orbit_integration, SRC/gorilla_plot_mod.f90:433-658.  Events go to a buffer instead of files.
 * Integer powers as libgcc __powidf2 forms them (see above): a**3 = a*(a*a), a**4 = (a*a)*(a*a), x**5 = x*((x*x)*(x*x)).
 * ---------------------------------------------------------------------------------------------- */
/* :3295-3322 (no case(1) in the reference: undefined for poly_order = 1, refused by the driver below) */
static double par_adiab_tau(int poly_order, double a44, double b4, double tau, double vpar_in)
{
  const double t2 = tau * tau, v2 = vpar_in * vpar_in, a2 = a44 * a44, bb = b4 * b4;
  const double a3 = a44 * a2, t3 = tau * t2, t4 = t2 * t2;
  double r = tau * v2 + 0.5 * t2 * (2.0 * b4 * vpar_in + 2.0 * a44 * v2);
  if (poly_order == 4) /* (sic) 2.d0*a44*2.d0*vpar_in**2 at :3316 */
    r = r + 1.0 / 3.0 * t3 * (bb + 3.0 * a44 * b4 * vpar_in + 2.0 * a44 * 2.0 * v2);
  else
    r = r + 1.0 / 3.0 * t3 * (bb + 3.0 * a44 * b4 * vpar_in + 2.0 * a2 * v2);
  if (poly_order >= 3) r = r + 1.0 / 4.0 * t4 * (a44 * bb + 7.0 / 3.0 * a2 * b4 * vpar_in + (4.0 * a3 * v2) / 3.0);
  if (poly_order >= 4) {
    const double a4 = a2 * a2, t5 = tau * t4;
    r = r + 1.0 / 60.0 * t5 * (7.0 * a2 * bb + 15.0 * a3 * b4 * vpar_in + 8.0 * a4 * v2);
  }
  return r;
}
/* :3326-3374 */
static double tau_vpar_root(int poly_order, double a44, double b4, double vpar_in, gor_trace *tr)
{
  const double a2 = a44 * a44, a3 = a44 * a2, a4 = a2 * a2;
  double c1 = vpar_in, c2 = b4 + a44 * vpar_in, c3 = a44 * b4 + a2 * vpar_in, c4 = 0.0, c5 = 0.0;
  if (poly_order >= 3) c4 = a2 * b4 + a3 * vpar_in;
  if (poly_order >= 4) c5 = a3 * b4 + a4 * vpar_in;
  switch (poly_order) {
    case 2: return Quadratic_Solver2(c3, c2, c1, tr);
    case 3: return Cubic_Solver(c4, c3, c2, c1, tr);
    default: return Quartic_Solver(0, c5, c4, c3, c2, c1, tr);
  }
}
typedef struct {
  const gor_event_settings *cfg;
  double par_adiab_inv;
  int32_t counter_banana_mappings, counter_phi_0_mappings;
  int64_t particle, push;
  gor_event *events;
  int64_t cap, n_events;
  double t; /* t_step - t_remain after the current push */
} event_state;
static void emit_event(event_state *es, int kind, int counter, const double x[3], double v0, double v1)
{
  if (es->n_events < es->cap) {
    gor_event *e = &es->events[es->n_events];
    e->particle = es->particle;
    e->kind = kind;
    e->counter = counter;
    e->push = es->push;
    for (int i = 0; i < 3; i++) e->x[i] = x[i];
    e->value[0] = v0;
    e->value[1] = v1;
    e->t = es->t;
  }
  es->n_events++;
}
/* par_adiab_inv_tetra_poly :3173-3291 (uses the pusher's module state after the push) */
static void par_adiab_inv_tetra_poly(poly_state *s, int poly_order, double vpar_in, double vpar_end, event_state *es)
{
  const double a44 = s->amat[3][3], b4 = s->b[3];
  const int nst = s->number_of_integration_steps;
  if ((vpar_end > 0.0) && (vpar_in < 0.0)) {
    int turning_index = -1; /* findloc(intermediate_z0_list(4,1:n) > 0) - 1 */
    for (int i = 1; i <= nst; i++)
      if (s->intermediate_z0_list[i - 1][3] > 0.0) {
        turning_index = i - 1;
        break;
      }
    if (turning_index == 0) return; /* reference: error stop (cannot happen: z0(4,1) = vpar_in < 0) */
    if (turning_index == -1) turning_index = nst;
    const double tau_part1 = tau_vpar_root(poly_order, a44, b4, s->intermediate_z0_list[turning_index - 1][3], s->tr);
    for (int i = 1; i <= turning_index - 1; i++)
      es->par_adiab_inv = es->par_adiab_inv +
                          par_adiab_tau(poly_order, a44, b4, s->tau_steps_list[i - 1], s->intermediate_z0_list[i - 1][3]) * s->dt_dtau_const;
    es->par_adiab_inv = es->par_adiab_inv +
                        par_adiab_tau(poly_order, a44, b4, tau_part1, s->intermediate_z0_list[turning_index - 1][3]) * s->dt_dtau_const;
    if (es->counter_banana_mappings > 1) {
      const int nskip = es->cfg->n_skip_vpar_0;
      if (es->counter_banana_mappings / nskip * nskip == es->counter_banana_mappings) {
        double z[4], x[3];
        memcpy(z, s->intermediate_z0_list[turning_index - 1], sizeof(z));
        set_integration_coef_manually(s, poly_order, z);
        { /* analytic_integration_external :3378-3408 */
          const double tau = tau_part1;
          if (poly_order >= 1)
            for (int i = 0; i < 4; i++) z[i] = z[i] + tau * (s->b[i] + s->amat_in_z[i]);
          if (poly_order >= 2) {
            double tau2_half = tau * tau * 0.5;
            for (int i = 0; i < 4; i++) z[i] = z[i] + tau2_half * (s->amat_in_b[i] + s->amat2_in_z[i]);
          }
          if (poly_order >= 3) {
            double tau3_sixth = (tau * tau) * tau / 6.0;
            for (int i = 0; i < 4; i++) z[i] = z[i] + tau3_sixth * (s->amat2_in_b[i] + s->amat3_in_z[i]);
          }
          if (poly_order >= 4) {
            double t2 = tau * tau;
            double tau4_twentyfourth = (t2 * t2) / 24.0;
            for (int i = 0; i < 4; i++) z[i] = z[i] + tau4_twentyfourth * (s->amat3_in_b[i] + s->amat4_in_z[i]);
          }
        }
        for (int i = 0; i < 3; i++) x[i] = z[i] + s->r[TP_X1 + i];
        emit_event(es, GOR_EVENT_VPAR_0, es->counter_banana_mappings, x, es->par_adiab_inv,
                   gor_energy_tot(s->m, z, s->perpinv, s->ind_tetr));
      }
    }
    es->counter_banana_mappings = es->counter_banana_mappings + 1;
    es->par_adiab_inv = 0.0;
    es->par_adiab_inv = es->par_adiab_inv +
                        par_adiab_tau(poly_order, a44, b4, s->tau_steps_list[turning_index - 1] - tau_part1, 0.0) * s->dt_dtau_const;
    for (int i = turning_index + 1; i <= nst; i++)
      es->par_adiab_inv = es->par_adiab_inv +
                          par_adiab_tau(poly_order, a44, b4, s->tau_steps_list[i - 1], s->intermediate_z0_list[i - 1][3]) * s->dt_dtau_const;
  } else {
    for (int i = 1; i <= nst; i++)
      es->par_adiab_inv = es->par_adiab_inv +
                          par_adiab_tau(poly_order, a44, b4, s->tau_steps_list[i - 1], s->intermediate_z0_list[i - 1][3]) * s->dt_dtau_const;
  }
}

/* ------------------------------------------------------------------------------------------------
 * RK-module pieces used by find_tetra -- SRC/pusher_tetra_rk.f90:50-193,810-896,2422-2467
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const gor_mesh *m;
  const double *r;
  int ind_tetr, iface_init, sign_rhs, sign_t_step_save;
  double perpinv, perpinv2, vmod_init, spamat, dt_dtau_const, dist_min, dist1, dist_max, t_remain;
  double Bvec[3], b[4], z_init[4], amat[3][3], anorm[4][3]; /* anorm[f][i] = anorm(i+1,f+1) */
  double dtau_ref, dtau_max, dtau_quad;
  int fb; /* per-push fall-back bits: 1 Newton failed, 2 last line of defence, 4 three-planes switch, 8 v_n>0 / stop outside */
  bool acc; /* boole_accuracy_ode45 of the routine that is running (= boole_pusher_ode45 except inside the Newton wrapper) */
} rk_state;

static void initialize_pusher_tetra_rk_mod(rk_state *s, int ind_tetr, const double x[3], int iface,
                                           double vpar, double t_remain_in)
{
  const gor_mesh *m = s->m;
  const double eps_distmin = 1.e-10, eps_distmax = 10.0, eps_modulation = 0.1, eps_dtau_max = 10.0,
               eps_dtau_quad = 1.5;
  s->t_remain = t_remain_in;
  s->ind_tetr = ind_tetr;
  const double *r = s->r = rec(m, ind_tetr);
  s->sign_t_step_save = isign1(s->t_remain);
  s->sign_rhs = m->sign_sqg * s->sign_t_step_save;
  for (int i = 0; i < 3; i++) s->z_init[i] = x[i] - r[TP_X1 + i];
  s->z_init[3] = vpar;
  s->iface_init = iface;
  double B0 = r[TP_BMOD1];
  for (int f = 0; f < 4; f++)
    for (int i = 0; i < 3; i++) s->anorm[f][i] = r[TP_ANORM + 3 * f + i];
  s->dt_dtau_const = r[TP_DT_DTAU_CONST];
  s->dt_dtau_const = s->dt_dtau_const * (double)s->sign_rhs;
  double bmod = gor_bmod(m, s->z_init, ind_tetr);
  double phi_elec = phi_elec_func(m, s->z_init, ind_tetr);
  double vperp2 = -2.0 * s->perpinv * bmod;
  double vpar2 = vpar * vpar;
  s->vmod_init = sqrt(vpar2 + vperp2);
  const double cm_over_e = m->cm_over_e, perpinv = s->perpinv;
  for (int i = 0; i < 3; i++)
    s->b[i] = (r[TP_CURLH + i] * (vperp2 + vpar2 + 2.0 * perpinv * B0) + perpinv * r[TP_GBXH1 + i]) * cm_over_e -
              CLIGHT * (2.0 * (r[TP_PHI1] - phi_elec) * r[TP_CURLH + i] + r[TP_GPHIXH1 + i]);
  s->b[3] = perpinv * r[TP_GBXCURLA] - CLIGHT / cm_over_e * r[TP_GPHIXCURLA];
  for (int i = 0; i < 3; i++) s->Bvec[i] = r[TP_CURLA + i];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      s->amat[i][j] = perpinv * cm_over_e * r[TP_ALPMAT + i + 3 * j] - CLIGHT * r[TP_BETMAT + i + 3 * j];
  s->spamat = perpinv * cm_over_e * r[TP_SPALPMAT] - CLIGHT * r[TP_SPBETMAT];
  if (m->boole_strong_electric_field) {
    double dv2 = v2_E_mod_func(m, s->z_init, ind_tetr) - r[TP_V2EMOD_1];
    for (int i = 0; i < 3; i++)
      s->b[i] = s->b[i] - 0.5 * cm_over_e * r[TP_GV2EMODXH1 + i] + cm_over_e * r[TP_CURLH + i] * dv2;
    s->b[3] = s->b[3] + cm_over_e * perpinv * r[TP_GBXCURLVE] - CLIGHT * r[TP_GPHIXCURLVE] -
              0.5 * cm_over_e * r[TP_GV2EMODXCURLVE] - 0.5 * r[TP_GV2EMODXCURLA];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) s->amat[i][j] = s->amat[i][j] - 0.5 * cm_over_e * r[TP_GAMMAT + i + 3 * j];
    s->spamat = s->spamat - 0.5 * cm_over_e * r[TP_SPGAMMAT];
    for (int i = 0; i < 3; i++) s->Bvec[i] = s->Bvec[i] + cm_over_e * r[TP_CURLVE + i];
  }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) s->amat[i][j] = s->amat[i][j] * (double)s->sign_rhs;
  for (int i = 0; i < 4; i++) s->b[i] = s->b[i] * (double)s->sign_rhs;
  for (int i = 0; i < 3; i++) s->Bvec[i] = s->Bvec[i] * (double)s->sign_rhs;
  s->spamat = s->spamat * (double)s->sign_rhs;
  s->dist1 = -r[TP_DIST_REF];
  s->dist_min = eps_distmin * fabs(s->dist1);
  s->dist_max = eps_distmax * fabs(s->dist1);
  double tetra_dist_ref = fabs(r[TP_TETRA_DIST_REF]);
  double vd_ExB;
  if (m->boole_strong_electric_field)
    vd_ExB = r[TP_VE_MOD_AVG];
  else
    vd_ExB = fabs(CLIGHT / bmod * r[TP_ER_MOD]);
  bool boole_vd_ExB = !(vd_ExB == 0.0), boole_vperp = !(vperp2 == 0.0);
  double c1 = fabs(tetra_dist_ref / s->z_init[3]);
  double dtau_ref = c1;
  if (boole_vperp) {
    double c2 = sqrt(tetra_dist_ref * s->vmod_init * r[TP_R1] / (vperp2 * m->grid_size[1] * eps_modulation));
    if (c2 < dtau_ref) dtau_ref = c2;
  }
  if (boole_vd_ExB) {
    double c3 = tetra_dist_ref / vd_ExB;
    if (c3 < dtau_ref) dtau_ref = c3;
  }
  /* boole_dt_dtau = .true. (default) */
  dtau_ref = fabs(dtau_ref / s->dt_dtau_const);
  s->dtau_ref = dtau_ref;
  s->dtau_max = eps_dtau_max * dtau_ref;
  s->dtau_quad = eps_dtau_quad * dtau_ref;
}
static void rhs_pusher_tetra_rk4(const rk_state *s, const double z[4], double dzdtau[4])
{
  for (int i = 0; i < 3; i++) {
    double mv = ((0.0 + s->amat[i][0] * z[0]) + s->amat[i][1] * z[1]) + s->amat[i][2] * z[2];
    dzdtau[i] = s->b[i] + mv + s->Bvec[i] * z[3];
  }
  dzdtau[3] = s->b[3] + s->spamat * z[3];
}
static void rk4_step(const rk_state *s, double y[4], double h, double dzdtau[4])
{
  double hh = h * 0.5, h6 = h / 6.0, dydx[4], yt[4], dyt[4], dym[4];
  rhs_pusher_tetra_rk4(s, y, dydx);
  for (int i = 0; i < 4; i++) yt[i] = y[i] + hh * dydx[i];
  rhs_pusher_tetra_rk4(s, yt, dyt);
  for (int i = 0; i < 4; i++) yt[i] = y[i] + hh * dyt[i];
  rhs_pusher_tetra_rk4(s, yt, dym);
  for (int i = 0; i < 4; i++) yt[i] = y[i] + h * dym[i];
  for (int i = 0; i < 4; i++) dym[i] = dyt[i] + dym[i];
  rhs_pusher_tetra_rk4(s, yt, dyt);
  for (int i = 0; i < 4; i++) y[i] = y[i] + h6 * (dydx[i] + dyt[i] + 2.0 * dym[i]);
  for (int i = 0; i < 4; i++) dzdtau[i] = dyt[i];
}
/* ------------------------------------------------------------------------------------------------
 * r8_fehl / r8_rkf45 (SRC/contrib/rkf45.f90:776-923, 925-1578) for neqn = 4 and the right-hand side
 * rhs_pusher_tetra_rk45 (SRC/pusher_tetra_rk.f90:900-910, the same as the RK4 one), and odeint_allroutines
 * (SRC/odeint_rkf45.f90).  The routine's SAVEd variables are the struct; x**0.2 is libm pow as gfortran calls it.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  double abserr_save, h, relerr_save, f1[5], f2[5], f3[5], f4[5], f5[5];
  int flag_save, init, kflag, kop, nfe;
  int neqn; /* 4: rhs_pusher_tetra_rk45 ; 5: rhs_par_adiab_ode45 (:2779-2790), z(5) = integral of v_par^2 */
} rkf45_state;
static void rkf45_rhs(const rk_state *s, int neqn, const double *z, double *dz)
{
  rhs_pusher_tetra_rk4(s, z, dz);
  if (neqn == 5) dz[4] = z[3] * z[3];
}
static void r8_fehl(const rk_state *s, int neqn, const double *y, double h, const double *yp, double *f1, double *f2,
                    double *f3, double *f4, double *f5, double *sout)
{
  double ch = h / 4.0, t1[5];
  for (int i = 0; i < neqn; i++) f5[i] = y[i] + ch * yp[i];
  rkf45_rhs(s, neqn, f5, f1);
  ch = 3.0 * h / 32.0;
  for (int i = 0; i < neqn; i++) f5[i] = y[i] + ch * (yp[i] + 3.0 * f1[i]);
  rkf45_rhs(s, neqn, f5, f2);
  ch = h / 2197.0;
  for (int i = 0; i < neqn; i++) f5[i] = y[i] + ch * (1932.0 * yp[i] + (7296.0 * f2[i] - 7200.0 * f1[i]));
  rkf45_rhs(s, neqn, f5, f3);
  ch = h / 4104.0;
  for (int i = 0; i < neqn; i++)
    f5[i] = y[i] + ch * ((8341.0 * yp[i] - 845.0 * f3[i]) + (29440.0 * f2[i] - 32832.0 * f1[i]));
  rkf45_rhs(s, neqn, f5, f4);
  ch = h / 20520.0;
  for (int i = 0; i < neqn; i++)
    t1[i] = y[i] + ch * ((-6080.0 * yp[i] + (9295.0 * f3[i] - 5643.0 * f4[i])) + (41040.0 * f1[i] - 28352.0 * f2[i]));
  memcpy(f1, t1, (size_t)neqn * sizeof(double));
  rkf45_rhs(s, neqn, f1, f5);
  ch = h / 7618050.0;
  /* the caller passes f1 as the output array s as well (call r8_fehl(..., f1, f2, f3, f4, f5, f1)) */
  for (int i = 0; i < neqn; i++)
    sout[i] = y[i] + ch * ((902880.0 * yp[i] + (3855735.0 * f3[i] - 1371249.0 * f4[i])) + (3953664.0 * f2[i] + 277020.0 * f5[i]));
}
/* returns the flag; only the branches reachable from odeint_allroutines are restated (flag = 1 first call, flag = 2 after a
 * return with 6 or 7); a fatal stop of the reference returns 8 */
static int r8_rkf45(const rk_state *s, rkf45_state *q, double *y, double *yp, double *t, double tout, double *relerr,
                    double abserr, int flag)
{
  const double remin = 1.0e-12, eps = DBL_EPSILON;
  const int maxnfe = 3000, neqn = q->neqn;
  if (*relerr < 0.0 || abserr < 0.0) return 8;
  if (flag == 0 || 8 < flag || flag < -2) return 8;
  int mflag = abs(flag);
  if (mflag != 1) {
    if (*t == tout && q->kflag != 3) return 8;
    if (mflag == 2) {
      if (q->kflag == 3) { flag = q->flag_save; mflag = abs(flag); }
      else if (q->init == 0) flag = q->flag_save;
      else if (q->kflag == 4) q->nfe = 0;
      else if (q->kflag == 5 && abserr == 0.0) return 8;
      else if (q->kflag == 6 && *relerr <= q->relerr_save && abserr <= q->abserr_save) return 8;
    } else {
      return 8;
    }
  }
  q->flag_save = flag;
  q->kflag = 0;
  q->relerr_save = *relerr;
  q->abserr_save = abserr;
  const double relerr_min = 2.0 * DBL_EPSILON + remin;
  if (*relerr < relerr_min) {
    *relerr = relerr_min;
    q->kflag = 3;
    return 3;
  }
  double dt = tout - *t;
  if (mflag == 1) {
    q->init = 0;
    q->kop = 0;
    rkf45_rhs(s, neqn, y, yp);
    q->nfe = 1;
    if (*t == tout) return 2;
  }
  if (q->init == 0) {
    q->init = 1;
    q->h = fabs(dt);
    double toln = 0.0;
    for (int k = 0; k < neqn; k++) {
      const double tol = *relerr * fabs(y[k]) + abserr;
      if (0.0 < tol) {
        toln = tol;
        const double ypk = fabs(yp[k]);
        const double h2 = q->h * q->h;
        if (tol < ypk * (q->h * (h2 * h2))) q->h = pow(tol / ypk, 0.2); /* h**5 = h*((h*h)*(h*h)) (__powidf2) */
      }
    }
    if (toln <= 0.0) q->h = 0.0;
    q->h = fmax(q->h, 26.0 * eps * fmax(fabs(*t), fabs(dt)));
    q->flag_save = flag < 0 ? -2 : 2;
  }
  q->h = copysign(q->h, dt);
  if (2.0 * fabs(dt) <= fabs(q->h)) q->kop = q->kop + 1;
  if (q->kop == 10000) {
    q->kop = 0;
    return 7;
  }
  if (fabs(dt) <= 26.0 * eps * fabs(*t)) {
    *t = tout;
    for (int i = 0; i < neqn; i++) y[i] = y[i] + dt * yp[i];
    rkf45_rhs(s, neqn, y, yp);
    q->nfe = q->nfe + 1;
    return 2;
  }
  bool output = false;
  const double scale = 2.0 / *relerr, ae = scale * abserr;
  for (;;) {
    bool hfaild = false;
    const double hmin = 26.0 * eps * fabs(*t);
    dt = tout - *t;
    if (!(2.0 * fabs(q->h) <= fabs(dt))) {
      if (fabs(dt) <= fabs(q->h)) { output = true; q->h = dt; }
      else q->h = 0.5 * dt;
    }
    double esttol;
    for (;;) {
      if (maxnfe < q->nfe) { q->kflag = 4; return 4; }
      r8_fehl(s, neqn, y, q->h, yp, q->f1, q->f2, q->f3, q->f4, q->f5, q->f1);
      q->nfe = q->nfe + 5;
      double eeoet = 0.0;
      for (int k = 0; k < neqn; k++) {
        const double et = fabs(y[k]) + fabs(q->f1[k]) + ae;
        if (et <= 0.0) return 5;
        const double ee = fabs((-2090.0 * yp[k] + (21970.0 * q->f3[k] - 15048.0 * q->f4[k])) +
                               (22528.0 * q->f2[k] - 27360.0 * q->f5[k]));
        eeoet = fmax(eeoet, ee / et);
      }
      esttol = fabs(q->h) * eeoet * scale / 752400.0;
      if (esttol <= 1.0) break;
      hfaild = true;
      output = false;
      double sf;
      if (esttol < 59049.0) sf = 0.9 / pow(esttol, 0.2);
      else sf = 0.1;
      q->h = sf * q->h;
      if (fabs(q->h) < hmin) { q->kflag = 6; return 6; }
    }
    *t = *t + q->h;
    memcpy(y, q->f1, (size_t)neqn * sizeof(double));
    rkf45_rhs(s, neqn, y, yp);
    q->nfe = q->nfe + 1;
    double sf;
    if (0.0001889568 < esttol) sf = 0.9 / pow(esttol, 0.2);
    else sf = 5.0;
    if (hfaild) sf = fmin(sf, 1.0);
    q->h = copysign(fmax(sf * fabs(q->h), hmin), q->h);
    if (output) {
      *t = tout;
      return 2;
    }
    if (flag <= 0) break;
  }
  return -2;
}
/* odeint_allroutines(y, nvar, 0, x2, eps, rhs) (SRC/odeint_rkf45.f90) */
static void odeint_allroutines_n(const rk_state *s, double *y, int neqn, double x2, double eps_rel)
{
  rkf45_state q;
  memset(&q, 0, sizeof(q));
  q.neqn = neqn;
  double yp[5], epsrel = eps_rel, epsabs = 1e-31, x1in = 0.0;
  int flag = r8_rkf45(s, &q, y, yp, &x1in, x2, &epsrel, epsabs, 1);
  if (flag == 6) {
    epsrel = 10 * epsrel;
    epsabs = 10 * epsabs;
    r8_rkf45(s, &q, y, yp, &x1in, x2, &epsrel, epsabs, 2);
  } else if (flag == 7) {
    r8_rkf45(s, &q, y, yp, &x1in, x2, &epsrel, epsabs, 2);
  }
}
static void odeint_allroutines(const rk_state *s, double y[4], double x2, double eps_rel)
{
  odeint_allroutines_n(s, y, 4, x2, eps_rel);
}
/* integration_step (:2549-2581): adaptive ODE45 over [0, dtau] followed by a zero-length RK4 step for dz/dtau, or one RK4 step */
static void integration_step(const rk_state *s, double z[4], double dtau, double dzdtau[4], bool boole_accuracy)
{
  if (boole_accuracy) {
    odeint_allroutines(s, z, dtau, s->m->rel_err_ode45);
    rk4_step(s, z, 0.0, dzdtau);
  } else {
    rk4_step(s, z, dtau, dzdtau);
  }
}
static void rk_normal_distances_func(const rk_state *s, const double z123[3], double out[4])
{
  for (int f = 0; f < 4; f++) out[f] = dot3(z123, s->anorm[f]);
  out[0] = out[0] - s->dist1;
}
/* the tetra_physics_poly4 record of the current tetrahedron (boole_newton_precalc) */
static inline const double *rk_p4rec(const rk_state *s)
{
  return s->m->tetra_physics_poly4 + ((int64_t)s->ind_tetr - 1) * P4_NDOUBLES;
}
/* sum((anorm_in_amat<k>_0 + perpinv*..._1 [+ perpinv2*..._2])(:,iface) * v) */
static double rk_p4_dot(const rk_state *s, int order, int iface, const double v[4])
{
  const double *p4 = rk_p4rec(s);
  const int base = order == 1 ? P4_AN_AMAT1_0 : P4_AN_AMAT2_0;
  double sacc = 0.0;
  for (int i = 0; i < 4; i++) {
    double e = p4[base + i + 4 * (iface - 1)] + s->perpinv * p4[base + 16 + i + 4 * (iface - 1)];
    if (order == 2) e = e + s->perpinv2 * p4[base + 32 + i + 4 * (iface - 1)];
    sacc = sacc + e * v[i];
  }
  return sacc;
}
/* :2451-2467 ; boole_newton_precalc: normal_velocity_analytic (:2487-2505) from the current position z */
static double rk_normal_velocity_func(const rk_state *s, int iface, const double dzdtau[4], const double z[4])
{
  if (s->m->boole_newton_precalc) return rk_p4_dot(s, 1, iface, z) * (double)s->sign_rhs + dot3(s->anorm[iface - 1], s->b);
  return dot3(dzdtau, s->anorm[iface - 1]);
}

/* ------------------------------------------------------------------------------------------------
 * pusher_tetra_rk -- SRC/pusher_tetra_rk.f90 (RK4 mode: boole_pusher_ode45 = .false., boole_dt_dtau = .true.,
 * boole_newton_precalc = .false.).  integration_step (:2549-2581) is then a single rk4_step.
 * ---------------------------------------------------------------------------------------------- */
#define RK_KITER 48
static double rk_normal_distance_func(const rk_state *s, const double z123[3], int iface)
{
  double d = dot3(z123, s->anorm[iface - 1]);
  if (iface == 1) d = d - s->dist1;
  return d;
}
/* :2469-2483  sum(matmul(anorm(:,iface),amat)*dzdtau(1:3)) + sum(anorm(:,iface)*Bvec)*dzdtau(4) */
static double rk_normal_acceleration_func(const rk_state *s, int iface, const double dzdtau[4], const double z[4])
{
  const double *n = s->anorm[iface - 1];
  if (s->m->boole_newton_precalc) /* normal_acceleration_analytic (:2507-2527) */
    return rk_p4_dot(s, 2, iface, z) + rk_p4_dot(s, 1, iface, s->b) * (double)s->sign_rhs;
  double t[3];
  for (int j = 0; j < 3; j++) t[j] = ((0.0 + n[0] * s->amat[0][j]) + n[1] * s->amat[1][j]) + n[2] * s->amat[2][j];
  return dot3(t, dzdtau) + dot3(n, s->Bvec) * dzdtau[3];
}
static bool rk_any_gt(const double d[4], double lim) { return d[0] > lim || d[1] > lim || d[2] > lim || d[3] > lim; }
static int rk_minloc(const double d[4])
{
  int k = 0;
  for (int i = 1; i < 4; i++)
    if (d[i] < d[k]) k = i;
  return k + 1;
}
static double rk_minval(const double d[4]) { return d[rk_minloc(d) - 1]; }

/* :636-809 */
static void quad_analytic_approx(const rk_state *s, const double z[4], const bool allowed_faces[4], int *iface_inout,
                                 double *dtau, bool *boole_quad_approx)
{
  double acoef[4], bcoef[4], ccoef[4], dtau_vec[4], discr, dummy;
  const double *r = s->r;
  if (s->m->boole_newton_precalc) { /* analytic_coeff(2, z, coef_mat) (:579-632) */
    for (int f = 0; f < 4; f++) {
      ccoef[f] = dot3(z, s->anorm[f]);
      bcoef[f] = rk_p4_dot(s, 1, f + 1, z) * (double)s->sign_rhs + dot3(s->anorm[f], s->b);
      acoef[f] = rk_p4_dot(s, 2, f + 1, z) + rk_p4_dot(s, 1, f + 1, s->b) * (double)s->sign_rhs;
    }
    ccoef[0] = ccoef[0] - s->dist1;
  } else {
  for (int f = 0; f < 4; f++) acoef[f] = r[TP_ACOEF_PRE + f] * (double)s->sign_rhs;
  if (s->m->boole_strong_electric_field) /* :672 */
    for (int f = 0; f < 4; f++)
      acoef[f] = acoef[f] + s->m->cm_over_e * r[TP_ACOEF_PRE_SE + f] * (double)s->sign_rhs;
  for (int f = 0; f < 4; f++) bcoef[f] = z[3] * acoef[f] + dot3(s->b, s->anorm[f]);
  for (int f = 0; f < 4; f++) acoef[f] = acoef[f] * (s->b[3] + s->spamat * z[3]);
  rk_normal_distances_func(s, z, ccoef);
  }
  const int iface = *iface_inout;
  for (int f = 0; f < 4; f++) dtau_vec[f] = s->dtau_max;
  for (int i = 0; i < 4; i++) {
    if (!allowed_faces[i]) continue;
    const double a = acoef[i], b = bcoef[i], c = ccoef[i];
    if (iface == i + 1) {
      if (a > 0.0) {
        if (b < 0.0) dtau_vec[i] = -2.0 * b / a;
      } else if (a < 0.0) {
        if (b > 0.0) dtau_vec[i] = -2.0 * b / a;
      }
    } else if (fabs(c) > s->dist_min) {
      if (c > 0.0) {
        if (a > 0.0) {
          if (b < 0.0) {
            discr = b * b - 2.0 * a * c;
            if (discr > 0.0) {
              dummy = (-b + sqrt(discr));
              if (fabs(dummy) > EPS) dtau_vec[i] = 2.0 * c / dummy;
              else dtau_vec[i] = (-sqrt(discr) - b) / a;
            } else if (discr == 0.0) {
              dtau_vec[i] = -b / a;
            }
          }
        } else if (a < 0.0) {
          discr = b * b - 2.0 * a * c;
          dummy = (-b + sqrt(discr));
          if (fabs(dummy) > EPS) dtau_vec[i] = 2.0 * c / dummy;
          else dtau_vec[i] = (-sqrt(discr) - b) / a;
        } else {
          if (b < 0.0) dtau_vec[i] = -c / b;
        }
      } else if (c < 0.0) {
        if (a < 0.0) {
          if (b > 0.0) {
            discr = b * b - 2.0 * a * c;
            if (discr > 0.0) dtau_vec[i] = (sqrt(discr) - b) / a;
            else if (discr == 0.0) dtau_vec[i] = -b / a;
          }
        } else if (a > 0.0) {
          discr = b * b - 2.0 * a * c;
          dtau_vec[i] = (sqrt(discr) - b) / a;
        } else {
          if (b > 0.0) dtau_vec[i] = -c / b;
        }
      }
      /* else (NaN): reference prints 'Should not happen' and stops */
    } else {
      if (((a > 0.0) && (b < 0.0)) || ((a < 0.0) && (b > 0.0))) dtau_vec[i] = -2.0 * b / a;
    }
  }
  int best = -1;
  for (int i = 0; i < 4; i++) {
    bool valid = (dtau_vec[i] < s->dtau_max) && (dtau_vec[i] > 0.0);
    if (valid && (best < 0 || dtau_vec[i] < dtau_vec[best])) best = i;
  }
  if (best >= 0) {
    *boole_quad_approx = true;
    *iface_inout = best + 1;
    *dtau = dtau_vec[best];
  } else {
    *boole_quad_approx = false;
  }
}

/* :1000-1188 (newton_face_convergence with RK4 accuracy == newton_face_convergence_wrapped + restore on failure) */
static void newton_face_convergence_wrapped(rk_state *s, double z[4], double *tau, int iface, double dzdtau[4],
                                            bool *converged, bool start_quadratic_in)
{
  double z_start[4], dzdtau_start[4], z_save[4], dzdtau_save[4];
  double tau_start = *tau, dtau = 0, tau_save = 0, dist, dist_new = 0.0, discr, nvel, nacc, nd[4];
  bool start_quadratic = start_quadratic_in;
  memcpy(z_start, z, sizeof(z_start));
  memcpy(dzdtau_start, dzdtau, sizeof(dzdtau_start));
  *converged = false;
  dist = rk_normal_distance_func(s, z, iface);
  int k = 0;
  while (fabs(dist) > s->dist_min) {
    k++;
    memcpy(z_save, z, sizeof(z_save));
    memcpy(dzdtau_save, dzdtau, sizeof(dzdtau_save));
    nvel = rk_normal_velocity_func(s, iface, dzdtau, z);
    if (nvel != 0.0) dtau = -dist / nvel;
    else return;
    tau_save = *tau;
    rk_normal_distances_func(s, z, nd);
    if (rk_any_gt(nd, s->dist_max)) {
      dtau = *tau + dtau;
      *tau = 0.0;
      memcpy(z, s->z_init, 4 * sizeof(double));
    }
    if (fabs(dtau) > s->dtau_max) {
      start_quadratic = true;
    } else {
      integration_step(s, z, dtau, dzdtau, s->acc);
      dist_new = rk_normal_distance_func(s, z, iface);
    }
    if ((fabs(dist_new) >= fabs(dist)) || start_quadratic) {
      start_quadratic = false;
      memcpy(z, z_save, sizeof(z_save));
      memcpy(dzdtau, dzdtau_save, sizeof(dzdtau_save));
      *tau = tau_save;
      nacc = 0.5 * rk_normal_acceleration_func(s, iface, dzdtau, z);
      discr = nvel * nvel - 4.0 * nacc * dist;
      if (discr > 0.0) {
        if (nacc < 0.0) dtau = (-nvel - sqrt(discr)) / (2.0 * nacc);
        else if (nacc > 0.0) dtau = (-nvel + sqrt(discr)) / (2.0 * nacc);
        else dtau = -dist / nvel;
        rk_normal_distances_func(s, z, nd);
        if (rk_any_gt(nd, s->dist_max)) {
          dtau = *tau + dtau;
          *tau = 0.0;
          memcpy(z, s->z_init, 4 * sizeof(double));
        }
        if (fabs(dtau) > s->dtau_max) {
          memcpy(z, z_start, sizeof(z_start));
          *tau = tau_start;
          memcpy(dzdtau, dzdtau_start, sizeof(dzdtau_start));
          return;
        }
        integration_step(s, z, dtau, dzdtau, s->acc);
        *tau = *tau + dtau;
        dist = rk_normal_distance_func(s, z, iface);
      } else {
        return;
      }
    } else {
      *tau = *tau + dtau;
      dist = dist_new;
    }
    if (k > RK_KITER) return;
  }
  if (*tau <= 0.0) {
    memcpy(z, z_start, sizeof(z_start));
    *tau = tau_start;
    memcpy(dzdtau, dzdtau_start, sizeof(dzdtau_start));
    return;
  }
  *converged = true;
}
/* :914-996 */
static void newton_face_convergence(rk_state *s, double z[4], double *tau, int iface, double dzdtau[4], bool *converged,
                                    bool start_quadratic)
{
  double z_save[4], dzdtau_save[4], tau_save = *tau;
  const bool acc_in = s->acc; /* boole_accuracy_ode45_in */
  memcpy(z_save, z, sizeof(z_save));
  memcpy(dzdtau_save, dzdtau, sizeof(dzdtau_save));
  s->acc = false; /* Newton with the RK4 method first (:938-941) */
  newton_face_convergence_wrapped(s, z, tau, iface, dzdtau, converged, start_quadratic);
  s->acc = acc_in;
  if (!*converged) {
    memcpy(z, z_save, sizeof(z_save));
    *tau = tau_save;
    memcpy(dzdtau, dzdtau_save, sizeof(dzdtau_save));
    return;
  }
  if (acc_in) { /* repeat the RK4-Newton step with ODE45 and converge with the ODE45-Newton (:950-979) */
    memcpy(z, z_save, sizeof(z_save));
    const double dtau = *tau - tau_save;
    integration_step(s, z, dtau, dzdtau, acc_in);
    newton_face_convergence_wrapped(s, z, tau, iface, dzdtau, converged, false);
  }
}

/* :1583-1623 ; streaming form of the 1000-step scan (z_mat/normal_distances_mat are not materialised) */
static void bisection_search_start(rk_state *s, double tau_in, int n_steps, double z_start[4], double *dtau, double *tau_out)
{
  double z_run[4], dzdtau[4], nd[4], tau_run = 0.0;
  memcpy(z_run, s->z_init, sizeof(z_run));
  *dtau = tau_in / (double)n_steps;
  /* start_index = (last index whose point is strictly inside) + 1; 0 + 1 = 1 when none is inside */
  int last_inside = 0;
  bool take_next = false;
  rk_normal_distances_func(s, z_run, nd);
  memcpy(z_start, z_run, 4 * sizeof(double));
  *tau_out = 0.0;
  if (nd[0] > 0.0 && nd[1] > 0.0 && nd[2] > 0.0 && nd[3] > 0.0) { last_inside = 1; take_next = true; }
  for (int i = 2; i <= n_steps; i++) {
    integration_step(s, z_run, *dtau, dzdtau, s->acc);
    tau_run = tau_run + *dtau;
    rk_normal_distances_func(s, z_run, nd);
    if (take_next) {
      memcpy(z_start, z_run, 4 * sizeof(double));
      *tau_out = tau_run;
      take_next = false;
    }
    if (nd[0] > 0.0 && nd[1] > 0.0 && nd[2] > 0.0 && nd[3] > 0.0) { last_inside = i; take_next = true; }
  }
  /* last point inside: the reference reads z_mat(:,n_steps+1) (out of bounds); keep the last point instead */
  if (last_inside == n_steps) {
    memcpy(z_start, z_run, 4 * sizeof(double));
    *tau_out = tau_run;
  }
}

/* :1409-1581 */
static void bisection_face_convergence(rk_state *s, double z[4], double *tau_inout, double dtau_in, int *iface,
                                       double dzdtau[4], bool *converged)
{
  double tau = *tau_inout, dtau = dtau_in, z_save[4], nd[4];
  memcpy(z_save, z, sizeof(z_save));
  *converged = false;
  for (int l = 1; l <= 2; l++) {
    int k = 0;
    while (!*converged) {
      rk_normal_distances_func(s, z, nd);
      double mn = rk_minval(nd);
      if (mn < -s->dist_min) {
        dtau = -fabs(dtau / 2.0);
        if (rk_any_gt(nd, s->dist_max)) {
          double dtau_save = dtau;
          dtau = tau + dtau;
          tau = 0.0;
          memcpy(z, s->z_init, 4 * sizeof(double));
          integration_step(s, z, dtau, dzdtau, s->acc);
          tau = tau + dtau;
          dtau = dtau_save;
        } else {
          integration_step(s, z, dtau, dzdtau, s->acc);
          tau = tau + dtau;
        }
      } else if (mn > s->dist_min) {
        dtau = +fabs(dtau / 2.0);
        integration_step(s, z, dtau, dzdtau, s->acc);
        tau = tau + dtau;
      }
      rk_normal_distances_func(s, z, nd);
      if (fabs(rk_minval(nd)) < s->dist_min) {
        if (rk_normal_velocity_func(s, rk_minloc(nd), dzdtau, z) > 0.0) {
          dtau = +fabs(dtau / 2.0);
          integration_step(s, z, dtau, dzdtau, s->acc);
          tau = tau + dtau;
        } else {
          int j = 0;
          for (int i = 0; i < 4; i++)
            if (nd[i] < 0.0) j++;
          if (j <= 1) {
            *iface = rk_minloc(nd);
            *converged = true;
          } else {
            dtau = -fabs(dtau / 2.0);
            integration_step(s, z, dtau, dzdtau, s->acc);
            tau = tau + dtau;
          }
        }
      }
      k++;
      if (k > RK_KITER) {
        if (l == 1) {
          int nfc = 0;
          for (int i = 0; i < 4; i++)
            if (nd[i] < 0.0 && fabs(nd[i]) < s->dist_min) nfc++;
          if (nfc > 1) {
            s->dist_min = 2.0 * s->dist_min;
            tau = *tau_inout;
            dtau = dtau_in;
            memcpy(z, z_save, sizeof(z_save));
          } else {
            bisection_search_start(s, *tau_inout, 1000, z, &dtau, &tau);
          }
          break;
        } else {
          goto done;
        }
      }
    }
    if (*converged) break;
  }
done:
  *tau_inout = tau;
}

/* :1696-2071 */
static void last_line_defense(rk_state *s, double z[4], double *tau, int *iface, double dzdtau[4], bool *llod_converged)
{
  bool allowed_faces[4] = {true, true, true, true};
  bool turned_tangential = false, converged = false, distance_bisection = false, quad_ok, bis_ok, newton_ok;
  double dtau, nd[4], z_save[4], dtau_save, tau_save;
  int iface_new = s->iface_init, iface_init_outside = 0, k;
  s->fb |= 2;
  *llod_converged = true;
  *tau = 0.0;
  memcpy(z, s->z_init, 4 * sizeof(double));
  if (s->iface_init != 0)
    if (rk_normal_distance_func(s, z, s->iface_init) < 0.0) iface_init_outside = s->iface_init;
  quad_analytic_approx(s, z, allowed_faces, &iface_new, &dtau, &quad_ok);
  if (quad_ok) {
    integration_step(s, z, dtau, dzdtau, s->acc);
    *tau = *tau + dtau;
  } else {
    dtau = s->dtau_ref;
    integration_step(s, z, dtau, dzdtau, s->acc);
    *tau = *tau + dtau;
    rk_normal_distances_func(s, z, nd);
    iface_new = rk_minloc(nd);
  }
  k = 0;
  for (;;) {
    k++;
    rk_normal_distances_func(s, z, nd);
    if (rk_any_gt(nd, s->dist_max)) {
      distance_bisection = true;
      dtau = *tau - 0.5 * fabs(dtau);
      *tau = 0.0;
      memcpy(z, s->z_init, 4 * sizeof(double));
      rk4_step(s, z, dtau, dzdtau); /* integration_step(..., .false.): "Set accuracy to FALSE, always!" (:1798) */
      *tau = *tau + dtau;
    } else {
      if (distance_bisection && s->acc) { /* :1802-1808 */
        memcpy(z, s->z_init, 4 * sizeof(double));
        dtau = *tau;
        integration_step(s, z, dtau, dzdtau, s->acc);
        distance_bisection = false;
      } else {
        break;
      }
    }
    if (k > RK_KITER) {
      *llod_converged = false;
      return;
    }
  }
  (void)distance_bisection;
  if (iface_init_outside != 0) {
    if (rk_normal_distance_func(s, z, iface_init_outside) < 0.0) {
      if (rk_normal_velocity_func(s, iface_init_outside, dzdtau, z) < 0.0) {
        for (int i = 1; i <= 3; i++) {
          int j = ((iface_init_outside + i - 1) % 4) + 1;
          if (rk_normal_distance_func(s, z, j) < 0.0) turned_tangential = true;
        }
        iface_new = iface_init_outside;
        if (fabs(rk_normal_distance_func(s, z, iface_new)) < s->dist_min) converged = true;
      }
    }
  }
  if (!converged) {
    if (turned_tangential) {
      bisection_face_convergence(s, z, tau, dtau, &iface_new, dzdtau, &bis_ok);
      if (!bis_ok) {
        *llod_converged = false;
        return;
      }
    } else {
      bool dtau_decreased = false;
      k = 0;
      for (;;) {
        k++;
        rk_normal_distances_func(s, z, nd);
        bool out[4];
        int n_out = 0;
        for (int i = 0; i < 4; i++) {
          out[i] = nd[i] < 0.0;
          if (out[i]) n_out++;
        }
        if (n_out == 0) {
          dtau = dtau_decreased ? 0.5 * fabs(dtau) : 2.0 * fabs(dtau);
        } else if (n_out == 1) {
          iface_new = 1;
          for (int i = 0; i < 4; i++)
            if (out[i]) { iface_new = i + 1; break; }
          if (rk_normal_velocity_func(s, iface_new, dzdtau, z) >= 0.0) {
            if (iface_init_outside != iface_new) {
              dtau = -0.5 * fabs(dtau);
              dtau_decreased = true;
            } else {
              dtau = dtau_decreased ? 0.5 * fabs(dtau) : 2.0 * fabs(dtau);
            }
          } else {
            break;
          }
        } else {
          int l = 0;
          for (int i = 0; i < 4; i++)
            if (out[i] && fabs(nd[i]) < s->dist_min) l++;
          if (l == n_out) {
            int j = 0;
            for (int i = 0; i < 4; i++) {
              if (!out[i]) continue;
              if (fabs(nd[i]) >= s->dist_min) continue;
              if (rk_normal_velocity_func(s, i + 1, dzdtau, z) > 0.0) j++;
            }
            if (j > 0) {
              dtau = 2.0 * fabs(dtau);
            } else {
              int best = -1;
              for (int i = 0; i < 4; i++)
                if (out[i] && (best < 0 || fabs(nd[i]) < fabs(nd[best]))) best = i;
              iface_new = best + 1;
              break;
            }
          } else {
            dtau = -0.5 * fabs(dtau);
            dtau_decreased = true;
          }
        }
        if (rk_any_gt(nd, s->dist_max)) {
          dtau = *tau + dtau;
          *tau = 0.0;
          memcpy(z, s->z_init, 4 * sizeof(double));
        }
        integration_step(s, z, dtau, dzdtau, s->acc);
        *tau = *tau + dtau;
        if (k > RK_KITER) {
          *llod_converged = false;
          return;
        }
      }
      memcpy(z_save, z, sizeof(z_save));
      dtau_save = dtau;
      tau_save = *tau;
      newton_face_convergence(s, z, tau, iface_new, dzdtau, &newton_ok, false);
      for (int i = 1; i <= 3; i++) {
        int j = ((iface_new + i - 1) % 4) + 1;
        if (rk_normal_distance_func(s, z, j) < 0.0) newton_ok = false;
      }
      if ((!newton_ok) || (rk_normal_velocity_func(s, iface_new, dzdtau, z) >= 0.0)) {
        memcpy(z, z_save, sizeof(z_save));
        *tau = tau_save;
        dtau = dtau_save;
        bisection_face_convergence(s, z, tau, dtau, &iface_new, dzdtau, &bis_ok);
        if (!bis_ok) {
          *llod_converged = false;
          return;
        }
      }
    }
  }
  *iface = iface_new;
}

/* :2075-2418 (boole_dt_dtau = .true.) ; returns false when final processing did not converge */
static bool rk_final_processing(rk_state *s, double z[4], double *tau, int *iface_inout, int *ind_tetr_out, int *iper_phi,
                                double x[3], double *vpar, double *t_pass, bool *boole_t_finished)
{
  const gor_mesh *m = s->m;
  int iface_new = *iface_inout;
  double dzdtau[4], nd[4], nd_save[4], z_save[4], dtau, tau_save;
  bool llod_ok, newton_ok, bis_ok;
  for (int i = 0; i < 3; i++) x[i] = z[i] + s->r[TP_X1 + i];
  *t_pass = *tau * s->dt_dtau_const;
  if (fabs(s->t_remain) < fabs(*t_pass)) {
    memcpy(z, s->z_init, 4 * sizeof(double));
    *tau = 0.0;
    dtau = s->t_remain / s->dt_dtau_const;
    memcpy(z_save, z, sizeof(z_save));
    rk_normal_distances_func(s, z, nd_save);
    integration_step(s, z, dtau, dzdtau, s->acc);
    if (rk_any_gt(nd_save, s->dist_max)) {
      memcpy(z, z_save, sizeof(z_save));
      last_line_defense(s, z, tau, &iface_new, dzdtau, &llod_ok);
      for (int j = 1; j <= 3; j++) {
        int k = ((iface_new + j - 1) % 4) + 1;
        if (rk_normal_distance_func(s, z, k) < 0.0) return false;
      }
      if (rk_normal_velocity_func(s, iface_new, dzdtau, z) > 0.0) return false;
      for (int i = 0; i < 3; i++) x[i] = z[i] + s->r[TP_X1 + i];
      *t_pass = *tau * s->dt_dtau_const;
      if (fabs(*t_pass) <= fabs(s->t_remain)) {
        *boole_t_finished = false;
        *vpar = z[3];
        *iface_inout = iface_new;
        handover2neighbour(m, s->ind_tetr, ind_tetr_out, iface_inout, x, iper_phi);
        return true; /* `exit` of the loop, then falls to the classification below in the reference: see note */
      }
      return false;
    }
    *tau = *tau + dtau;
    for (int i = 0; i < 3; i++) x[i] = z[i] + s->r[TP_X1 + i];
    rk_normal_distances_func(s, z, nd);
    *t_pass = *tau * s->dt_dtau_const;
    int i_outside_plane = 0, iface_outside = 0;
    for (int i = 0; i < 4; i++)
      if (nd[i] < 0.0) { i_outside_plane++; iface_outside = i + 1; }
    bool any_conv = false;
    for (int i = 0; i < 4; i++)
      if (fabs(nd[i]) < s->dist_min) any_conv = true;
    if (any_conv) {
      if (i_outside_plane == 1) iface_new = iface_outside;
      else
        for (int i = 0; i < 4; i++)
          if (fabs(nd[i]) < s->dist_min) iface_new = i + 1;
      for (int i = 0; i < 3; i++) x[i] = z[i] + s->r[TP_X1 + i];
      *t_pass = *tau * s->dt_dtau_const;
      *boole_t_finished = true;
      *vpar = z[3];
      if (rk_normal_velocity_func(s, iface_new, dzdtau, z) < 0.0) {
        *iface_inout = iface_new;
        handover2neighbour(m, s->ind_tetr, ind_tetr_out, iface_inout, x, iper_phi);
      } else {
        *ind_tetr_out = s->ind_tetr;
        *iface_inout = iface_new;
      }
    } else if (i_outside_plane != 0) {
      s->fb |= 8;
      if (i_outside_plane == 1) {
        iface_new = iface_outside;
        tau_save = *tau;
        memcpy(z_save, z, sizeof(z_save));
        newton_face_convergence(s, z, tau, iface_new, dzdtau, &newton_ok, true);
        if (!newton_ok) {
          memcpy(z, z_save, sizeof(z_save));
          *tau = tau_save;
          rk4_step(s, z, 0.0, dzdtau);
          newton_face_convergence(s, z, tau, iface_new, dzdtau, &newton_ok, false);
          if (!newton_ok) {
            memcpy(z, z_save, sizeof(z_save));
            *tau = tau_save;
            bisection_face_convergence(s, z, tau, *tau, &iface_new, dzdtau, &bis_ok);
            if (!bis_ok) return false;
          }
        }
        if (rk_normal_velocity_func(s, iface_new, dzdtau, z) > 0.0) {
          memcpy(z, z_save, sizeof(z_save));
          *tau = tau_save;
          bisection_face_convergence(s, z, tau, *tau, &iface_new, dzdtau, &bis_ok);
          if (!bis_ok) return false;
        }
      } else {
        bisection_face_convergence(s, z, tau, *tau, &iface_new, dzdtau, &bis_ok);
        if (!bis_ok) return false;
      }
      for (int i = 0; i < 3; i++) x[i] = z[i] + s->r[TP_X1 + i];
      *t_pass = *tau * s->dt_dtau_const;
      *boole_t_finished = false;
      *vpar = z[3];
      *iface_inout = iface_new;
      handover2neighbour(m, s->ind_tetr, ind_tetr_out, iface_inout, x, iper_phi);
    } else {
      *vpar = z[3];
      *boole_t_finished = true;
      *ind_tetr_out = s->ind_tetr;
      *iface_inout = 0;
      *iper_phi = 0;
    }
  } else {
    *boole_t_finished = false;
    *vpar = z[3];
    *iface_inout = iface_new;
    handover2neighbour(m, s->ind_tetr, ind_tetr_out, iface_inout, x, iper_phi);
  }
  return true;
}

/* :197-575 */
static void pusher_tetra_rk(rk_state *s, int *ind_tetr_inout, int *iface, double x[3], double *vpar, double z_final[3],
                            double t_remain_in, double *t_pass, bool *boole_t_finished, int *iper_phi, gor_trace *tr)
{
  bool allowed_faces[4] = {true, true, true, true};
  bool quad_ok, newton_ok, llod_ok, boole_converged = false;
  double z[4], dzdtau[4], tau = 0.0, dtau, nd[4];
  int iface_new;
  initialize_pusher_tetra_rk_mod(s, *ind_tetr_inout, x, *iface, *vpar, t_remain_in);
  s->fb = 0;
  memcpy(z, s->z_init, sizeof(z));
  iface_new = *iface;
  *boole_t_finished = false;
  *t_pass = 0.0;
  *iper_phi = 0;
  s->acc = s->m->boole_pusher_ode45 != 0; /* :261 */
  quad_analytic_approx(s, z, allowed_faces, &iface_new, &dtau, &quad_ok);
  if (quad_ok) {
    integration_step(s, z, dtau, dzdtau, s->acc);
    tau = tau + dtau;
  } else {
    dtau = s->dtau_ref;
    rk4_step(s, z, dtau, dzdtau); /* a plain rk4_step in the reference (:315) */
    tau = tau + dtau;
    rk_normal_distances_func(s, z, nd);
    iface_new = 1;
    for (int i = 1; i < 4; i++)
      if (fabs(nd[i]) < fabs(nd[iface_new - 1])) iface_new = i + 1;
  }
  rk_normal_distances_func(s, z, nd);
  if (rk_any_gt(nd, s->dist_max)) {
    last_line_defense(s, z, &tau, &iface_new, dzdtau, &llod_ok);
    if (!llod_ok) {
      *ind_tetr_inout = -1;
      *iface = -1;
      goto count;
    }
  }
#define RK_LLOD_CYCLE()                                                 \
  do {                                                                  \
    last_line_defense(s, z, &tau, &iface_new, dzdtau, &llod_ok);        \
    boole_converged = false;                                            \
  } while (0)
  for (int i = 1; i <= 5; i++) {
    boole_converged = true;
    newton_face_convergence(s, z, &tau, iface_new, dzdtau, &newton_ok, false);
    if (!newton_ok) {
      s->fb |= 1;
      allowed_faces[iface_new - 1] = false;
      if (!allowed_faces[0] && !allowed_faces[1] && !allowed_faces[2] && !allowed_faces[3]) { RK_LLOD_CYCLE(); continue; }
      if (tau > s->dtau_quad) {
        memcpy(z, s->z_init, sizeof(z));
        tau = 0.0;
      }
      quad_analytic_approx(s, z, allowed_faces, &iface_new, &dtau, &quad_ok);
      if (!quad_ok) { RK_LLOD_CYCLE(); continue; }
      rk_normal_distances_func(s, z, nd);
      if (rk_any_gt(nd, s->dist_max)) {
        dtau = tau + dtau;
        tau = 0.0;
        memcpy(z, s->z_init, sizeof(z));
      }
      integration_step(s, z, dtau, dzdtau, s->acc);
      tau = tau + dtau;
      boole_converged = false;
      continue;
    }
    bool cycled = false;
    for (int j = 1; j <= 3; j++) {
      int k = ((iface_new + j - 1) % 4) + 1;
      if (rk_normal_distance_func(s, z, k) < 0.0) {
        s->fb |= 4;
        allowed_faces[iface_new - 1] = false;
        if (!allowed_faces[0] && !allowed_faces[1] && !allowed_faces[2] && !allowed_faces[3]) { RK_LLOD_CYCLE(); cycled = true; break; }
        if (allowed_faces[k - 1]) {
          iface_new = k;
          boole_converged = false;
          cycled = true;
          break;
        } else {
          RK_LLOD_CYCLE();
          cycled = true;
          break;
        }
      }
    }
    if (cycled) continue;
    if (rk_normal_velocity_func(s, iface_new, dzdtau, z) > 0.0) {
      s->fb |= 8;
      allowed_faces[iface_new - 1] = false;
      if (!allowed_faces[0] && !allowed_faces[1] && !allowed_faces[2] && !allowed_faces[3]) { RK_LLOD_CYCLE(); continue; }
      if (tau > s->dtau_quad) {
        memcpy(z, s->z_init, sizeof(z));
        tau = 0.0;
      }
      quad_analytic_approx(s, z, allowed_faces, &iface_new, &dtau, &quad_ok);
      if (!quad_ok) { RK_LLOD_CYCLE(); continue; }
      rk_normal_distances_func(s, z, nd);
      if (rk_any_gt(nd, s->dist_max)) {
        dtau = tau + dtau;
        tau = 0.0;
        memcpy(z, s->z_init, sizeof(z));
      }
      integration_step(s, z, dtau, dzdtau, s->acc);
      tau = tau + dtau;
      boole_converged = false;
      continue;
    }
    if (tau <= 0.0) {
      allowed_faces[iface_new - 1] = false;
      memcpy(z, s->z_init, sizeof(z));
      tau = 0.0;
      if (!allowed_faces[0] && !allowed_faces[1] && !allowed_faces[2] && !allowed_faces[3]) { RK_LLOD_CYCLE(); continue; }
      quad_analytic_approx(s, z, allowed_faces, &iface_new, &dtau, &quad_ok);
      if (!quad_ok) { RK_LLOD_CYCLE(); continue; }
      integration_step(s, z, dtau, dzdtau, s->acc);
      tau = tau + dtau;
      boole_converged = false;
      continue;
    }
    break;
  }
#undef RK_LLOD_CYCLE
  if (!boole_converged) {
    *ind_tetr_inout = -1;
    *iface = -1;
    goto count;
  }
  *iface = iface_new;
  int ind_out = *ind_tetr_inout;
  if (!rk_final_processing(s, z, &tau, iface, &ind_out, iper_phi, x, vpar, t_pass, boole_t_finished)) {
    *ind_tetr_inout = -1;
    *iface = -1;
    goto count;
  }
  *ind_tetr_inout = ind_out;
  for (int i = 0; i < 3; i++) z_final[i] = z[i];
  if ((fabs(*t_pass) >= fabs(s->t_remain)) && (!*boole_t_finished)) {
    *ind_tetr_inout = -1;
    *iface = -1;
    goto count;
  }
  if ((*t_pass * (double)s->sign_t_step_save) <= 0.0) {
    *ind_tetr_inout = -1;
    *iface = -1;
    goto count;
  }
count:
  if (tr)
    for (int q = 0; q < 4; q++)
      if (s->fb & (1 << q)) tr->n_fallback[q]++;
}

/* SRC/tetra_physics_mod.f90:1038-1072 */
static bool isinside(const gor_mesh *m, int ind_tetr, const double x[3], double cur_dist_value[4])
{
  const double *r = rec(m, ind_tetr);
  double dist_min = EPS * fabs(r[TP_DIST_REF]);
  double d[3] = {x[0] - r[TP_X1], x[1] - r[TP_X1 + 1], x[2] - r[TP_X1 + 2]};
  bool all_ok = true;
  for (int f = 0; f < 4; f++) {
    double v = dot3(r + TP_ANORM + 3 * f, d);
    if (f == 0) v = v + r[TP_DIST_REF];
    cur_dist_value[f] = v;
    if (!(v >= -dist_min)) all_ok = false;
  }
  return all_ok;
}

/* SRC/find_tetra_mod.f90:283-600 (boole_grid_for_find_tetra = .false.) */
void gor_find_tetra(const gor_mesh *m, double x[3], double vpar, double vperp, int32_t *ind_tetr_out,
                    int32_t *iface, int sign_t_step)
{
  int64_t indtetr_start = 0, ntetr_searched = 0, ntetr = m->ntetr;
  int numerical_corr_plus = 0, numerical_corr_minus = 0, nphi = m->grid_size[1];
  *ind_tetr_out = -1;
  *iface = -1;
  if (m->grid_kind == 1 || m->grid_kind == 5) {
    int nr = m->grid_size[0], nz = m->grid_size[2];
    double hr = (m->Rmax - m->Rmin) / nr, hphi = (2.0 * PI) / nphi, hz = (m->Zmax - m->Zmin) / nz;
    int ir = (int)((x[0] - m->Rmin) / hr) + 1, iphi = (int)(x[1] / hphi) + 1,
        iz = (int)((x[2] - m->Zmin) / hz) + 1;
    if (ir < 1 || ir > nr || iphi < 1 || iphi > nphi || iz < 1 || iz > nz) return; /* reference: stop */
    indtetr_start = (int64_t)(((double)iz - 1.0) * 6.0 + 6.0 * (double)nz * ((double)ir - 1.0) +
                              6.0 * ((double)iphi - 1.0) * (double)nr * (double)nz + 1.0);
    ntetr_searched = 6;
  } else {
    int ind_b = (m->coord_system == 2) ? 2 : 1; /* 0-based slot of phi */
    int64_t ntetr_in_plane = ntetr / nphi;
    double q = x[ind_b] * nphi / (2.0 * PI / m->n_field_periods);
    int ind_plane_tetra_start = (int)q;
    if (fabs(q - (double)ind_plane_tetra_start) > (1.0 - EPS)) numerical_corr_plus = 1;
    if (fabs(q - (double)ind_plane_tetra_start) < EPS) numerical_corr_minus = 1;
    indtetr_start = (int64_t)ind_plane_tetra_start * ntetr_in_plane + 1;
    ntetr_searched = ntetr_in_plane * (1 + numerical_corr_plus + numerical_corr_minus);
  }
  for (int64_t i = 1; i <= ntetr_searched; i++) {
    int64_t ind_search = indtetr_start + i - 1 - (int64_t)numerical_corr_minus * (ntetr / nphi);
    if (ind_search > ntetr) ind_search -= ntetr;
    if (ind_search <= 0) ind_search += ntetr;
    double cur_dist_value[4];
    if (isinside(m, (int)ind_search, x, cur_dist_value)) {
      *ind_tetr_out = (int)ind_search;
      *iface = 0;
      const double *r = rec(m, (int)ind_search);
      bool conv[4];
      int n_plane_conv = 0;
      for (int f = 0; f < 4; f++) {
        conv[f] = fabs(cur_dist_value[f]) <= (EPS * fabs(r[TP_DIST_REF]));
        if (conv[f]) n_plane_conv++;
      }
      if (n_plane_conv > 0) {
        rk_state rk;
        memset(&rk, 0, sizeof(rk));
        rk.m = m;
        double vperp2 = vperp * vperp, z[4], dzdtau[4];
        for (int k = 0; k < 3; k++) z[k] = x[k] - r[TP_X1 + k];
        z[3] = vpar;
        int iface_new = 1;
        for (int f = 1; f < 4; f++)
          if (fabs(cur_dist_value[f]) < fabs(cur_dist_value[iface_new - 1])) iface_new = f + 1;
        int ind_tetr_tried[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        rk.perpinv = -0.5 * vperp2 / (r[TP_BMOD1] + dot3(r + TP_GB, z));
        rk.perpinv2 = rk.perpinv * rk.perpinv;
        for (int i_try = 1; i_try <= 2 * n_plane_conv; i_try++) {
          if (*ind_tetr_out == -1) break;
          ind_tetr_tried[i_try - 1] = *ind_tetr_out;
          const double *rr = rec(m, *ind_tetr_out);
          for (int k = 0; k < 3; k++) z[k] = x[k] - rr[TP_X1 + k];
          initialize_pusher_tetra_rk_mod(&rk, *ind_tetr_out, x, iface_new, vpar, (double)sign_t_step);
          rk_normal_distances_func(&rk, z, cur_dist_value);
          /* reference stops if any(cur_dist_value < -dist_min); unreachable for isinside-accepted starts */
          bool conv_t[4];
          for (int f = 0; f < 4; f++) conv_t[f] = fabs(cur_dist_value[f]) <= (EPS * fabs(rr[TP_DIST_REF]));
          rk4_step(&rk, z, 0.0, dzdtau);
          int counter_vnorm_pos = 0;
          for (int l = 1; l <= 4; l++) {
            if (!conv_t[l - 1]) continue;
            if (rk_normal_velocity_func(&rk, l, dzdtau, z) > 0.0) counter_vnorm_pos++;
          }
          if (counter_vnorm_pos == n_plane_conv) {
            *iface = iface_new;
            break;
          } else {
            int ind_tetr_save = *ind_tetr_out, iface_new_save = iface_new, iper_phi;
            double x_save[3] = {x[0], x[1], x[2]};
            for (int l = 1; l <= 4; l++) {
              if (!conv_t[l - 1]) continue;
              if (rk_normal_velocity_func(&rk, l, dzdtau, z) > 0.0) continue;
              iface_new = l;
              int out;
              handover2neighbour(m, ind_tetr_save, &out, &iface_new, x, &iper_phi);
              *ind_tetr_out = out;
              bool tried = false;
              for (int t = 0; t < 2 * n_plane_conv; t++)
                if (ind_tetr_tried[t] == out) tried = true;
              if (tried || out == -1) {
                x[0] = x_save[0]; x[1] = x_save[1]; x[2] = x_save[2];
                iface_new = iface_new_save;
              } else {
                break;
              }
            }
          }
        }
      }
      if (*ind_tetr_out == -1)
        *iface = -1;
      else
        break;
    } else {
      *ind_tetr_out = -1;
      *iface = -1;
    }
  }
}

/* SRC/orbit_timestep_gorilla.f90:278-358 ; Fortran modulo(a,p) for reals as gfortran expands it (trans-intrinsic.cc,
 * gfc_conv_intrinsic_mod): r = fmod(a,p) -- exact --, r += p when r != 0 and its sign differs from p's, and a zero result
 * takes the sign of p.  (a - floor(a/p)*p rounds k*p and differs in the last bits once |a| >= 2p.) */
static double f_modulo(double a, double p)
{
  double r = fmod(a, p);
  if (r != 0.0 && ((r < 0.0) != (p < 0.0))) r += p;
  if (r == 0.0) r = copysign(0.0, p);
  return r;
}
int gor_check_coordinate_domain(const gor_mesh *m, double x[3])
{
  double per = 2.0 * PI / m->n_field_periods;
  if (m->coord_system == 1) {
    if (m->boole_periodic_relocation)
      x[1] = f_modulo(x[1], per);
    else if (x[1] < 0.0 || x[1] > per)
      return GOR_ERR_DOMAIN;
  } else {
    if (x[0] < m->sfc_s_min || x[0] > 1.0) return GOR_ERR_DOMAIN;
    if (m->boole_periodic_relocation) {
      x[1] = f_modulo(x[1], 2.0 * PI);
      x[2] = f_modulo(x[2], per);
    } else if (x[1] < 0.0 || x[1] > 2.0 * PI || x[2] < 0.0 || x[2] > per) {
      return GOR_ERR_DOMAIN;
    }
  }
  return GOR_OK;
}

/* SRC/orbit_timestep_gorilla.f90:19-147 (ipusher = 2) */
static int orbit_timestep_core(const gor_mesh *m, double x[3], double *vpar, double *vperp, double t_step,
                               int32_t *boole_initialized, int32_t *ind_tetr, int32_t *iface, double *t_remain_out,
                               gor_trace *tr, event_state *es);
int gor_orbit_timestep(const gor_mesh *m, double x[3], double *vpar, double *vperp, double t_step,
                       int32_t *boole_initialized, int32_t *ind_tetr, int32_t *iface, double *t_remain_out,
                       gor_trace *tr)
{
  return orbit_timestep_core(m, x, vpar, vperp, t_step, boole_initialized, ind_tetr, iface, t_remain_out, tr, NULL);
}
/* orbit_timestep_gorilla with the event capture of gorilla_plot_orbit_integration (gorilla_plot_mod.f90:520-638):
 * after every push that does not end the time step, J_par / banana tips (:585-596) and toroidal mappings (:601-638). */
/* ------------------------------------------------------------------------------------------------
 * module par_adiab_inv_rk_mod (SRC/pusher_tetra_rk.f90:2589-2798): J_par and banana tips for the RK pusher.  v_par^2 is
 * integrated along the orbit as a fifth equation of the RKF45 integration from z_init over tau = t_pass / dt_dtau_const; the
 * bounce point is bracketed by halving the step with alternating sign until |v_par| <= 10 cm/s.
 * ---------------------------------------------------------------------------------------------- */
static void calc_par_adiab_tau(const rk_state *s, double dtau, double z_inout[4], double *par_adiab_tau)
{
  double z[5] = {z_inout[0], z_inout[1], z_inout[2], z_inout[3], 0.0};
  odeint_allroutines_n(s, z, 5, dtau, s->m->rel_err_ode45);
  *par_adiab_tau = z[4];
  memcpy(z_inout, z, 4 * sizeof(double));
}
static void calc_par_adiab_until_root(const rk_state *s, double tau_in, double z_inout[4], double *par_adiab_tau, double *tau_out)
{
  double z[5] = {z_inout[0], z_inout[1], z_inout[2], z_inout[3], 0.0};
  const double vpar_min = 1.e1;
  double dtau = tau_in;
  *tau_out = 0.0;
  int i = 0;
  while (fabs(z[3]) > vpar_min) {
    i++;
    const double vpar_save = z[3];
    odeint_allroutines_n(s, z, 5, dtau, s->m->rel_err_ode45);
    *tau_out = *tau_out + dtau;
    /* same sign of v_par as before the step: keep the direction, else turn around; half the length either way */
    const bool same_side = (vpar_save > 0.0) == (z[3] > 0.0);
    if (same_side) dtau = (dtau > 0.0) ? fabs(dtau / 2) : -fabs(dtau / 2);
    else dtau = (dtau > 0.0) ? -fabs(dtau / 2) : fabs(dtau / 2);
    if (i > 100) break; /* reference: print + stop */
  }
  *par_adiab_tau = z[4];
  memcpy(z_inout, z, 4 * sizeof(double));
}
static void par_adiab_inv_tetra_rk(const rk_state *s, double t_pass, double vpar_in, double vpar_end, event_state *es)
{
  const double tau = t_pass / s->dt_dtau_const;
  double z[4], par_adiab_tau, tau_part1;
  memcpy(z, s->z_init, sizeof(z));
  if ((vpar_end > 0.0) && (vpar_in < 0.0)) {
    calc_par_adiab_until_root(s, tau, z, &par_adiab_tau, &tau_part1);
    es->par_adiab_inv = es->par_adiab_inv + par_adiab_tau * s->dt_dtau_const;
    if (es->counter_banana_mappings > 1) {
      const int nskip = es->cfg->n_skip_vpar_0;
      if (es->counter_banana_mappings / nskip * nskip == es->counter_banana_mappings) {
        double x[3];
        for (int i = 0; i < 3; i++) x[i] = z[i] + s->r[TP_X1 + i];
        emit_event(es, GOR_EVENT_VPAR_0, es->counter_banana_mappings, x, es->par_adiab_inv,
                   gor_energy_tot(s->m, z, s->perpinv, s->ind_tetr));
      }
    }
    es->counter_banana_mappings = es->counter_banana_mappings + 1;
    es->par_adiab_inv = 0.0;
    calc_par_adiab_tau(s, tau - tau_part1, z, &par_adiab_tau);
    es->par_adiab_inv = es->par_adiab_inv + par_adiab_tau * s->dt_dtau_const;
  } else {
    calc_par_adiab_tau(s, tau, z, &par_adiab_tau);
    es->par_adiab_inv = es->par_adiab_inv + par_adiab_tau * s->dt_dtau_const;
  }
}

int gor_orbit_timestep_events(const gor_mesh *m, double x[3], double *vpar, double *vperp, double t_step,
                              int32_t *boole_initialized, int32_t *ind_tetr, int32_t *iface, double *t_remain_out,
                              gor_trace *tr, const gor_event_settings *cfg, double *par_adiab_inv,
                              int32_t *counter_vpar_0, int32_t *counter_phi_0, int64_t particle, gor_event *events,
                              int64_t cap, int64_t *n_events)
{
  /* polynomial orders 2..4 (par_adiab_tau :3302-3320) or the RK pusher (par_adiab_inv_rk_mod) */
  if (m->ipusher == 2 && m->poly_order < 2) return GOR_ERR_CONFIG;
  if ((cfg->boole_J_par || cfg->boole_poincare_vpar_0) && cfg->n_skip_vpar_0 < 1) return GOR_ERR_CONFIG;
  if (cfg->boole_poincare_phi_0 && cfg->n_skip_phi_0 < 1) return GOR_ERR_CONFIG;
  if (cfg->boole_full_orbit && cfg->n_skip_full_orbit < 1) return GOR_ERR_CONFIG;
  event_state es;
  es.cfg = cfg;
  es.par_adiab_inv = *par_adiab_inv;
  es.counter_banana_mappings = *counter_vpar_0;
  es.counter_phi_0_mappings = *counter_phi_0;
  es.particle = particle;
  es.push = 0;
  es.t = 0.0;
  es.events = events;
  es.cap = cap;
  es.n_events = *n_events;
  int rc = orbit_timestep_core(m, x, vpar, vperp, t_step, boole_initialized, ind_tetr, iface, t_remain_out, tr, &es);
  *par_adiab_inv = es.par_adiab_inv;
  *counter_vpar_0 = es.counter_banana_mappings;
  *counter_phi_0 = es.counter_phi_0_mappings;
  *n_events = es.n_events;
  return rc;
}
static int orbit_timestep_core(const gor_mesh *m, double x[3], double *vpar, double *vperp, double t_step,
                               int32_t *boole_initialized, int32_t *ind_tetr, int32_t *iface, double *t_remain_out,
                               gor_trace *tr, event_state *es)
{
  if (m->ipusher != 2 && m->ipusher != 1) return GOR_ERR_CONFIG;
  if (!*boole_initialized) {
    int rcd = gor_check_coordinate_domain(m, x);
    if (rcd != GOR_OK) return rcd;
    int sign_t_step = isign1(t_step);
    gor_find_tetra(m, x, *vpar, *vperp, ind_tetr, iface, sign_t_step);
    if (*ind_tetr == -1) return GOR_OK;
    *boole_initialized = 1;
  }
  if (t_step == 0.0) return GOR_OK;
  /* A particle that is already lost (ind_tetr = -1 from an earlier call) is left untouched.  The reference
   * would index tetra_physics(-1) at :74 (out of bounds; its callers never re-enter with a lost particle). */
  if (*ind_tetr < 1) {
    if (t_remain_out) *t_remain_out = t_step;
    return GOR_OK;
  }
  double vperp2 = (*vperp) * (*vperp);
  double z_save[3];
  {
    const double *r = rec(m, *ind_tetr);
    for (int i = 0; i < 3; i++) z_save[i] = x[i] - r[TP_X1 + i];
  }
  poly_state s;
  memset(&s, 0, sizeof(s));
  s.m = m;
  s.tr = tr;
  s.tau_steps_list = s.list_tau_static;
  s.intermediate_z0_list = s.list_z0_static;
  s.list_cap = 2;
  void *list_heap = NULL;
  if (m->boole_adaptive_time_steps && m->ipusher == 2) { /* manage_intermediate_steps_arrays :98-101 */
    if (m->desired_delta_energy <= 0.0 || m->max_n_intermediate_steps < 2) return GOR_ERR_CONFIG; /* :868-874 */
    s.list_cap = 3 * m->max_n_intermediate_steps;
    list_heap = malloc(((size_t)s.list_cap * 6 + 1) * sizeof(double));
    s.tau_steps_list = (double *)list_heap;
    s.intermediate_z0_list = (double(*)[4])((double *)list_heap + s.list_cap);
    s.thl_heap = (double *)list_heap + (size_t)s.list_cap * 5;
  }
  s.perpinv = -0.5 * vperp2 / gor_bmod(m, z_save, *ind_tetr);
  s.perpinv2 = s.perpinv * s.perpinv;
  rk_state rk;
  memset(&rk, 0, sizeof(rk));
  rk.m = m;
  rk.perpinv = s.perpinv;
  rk.perpinv2 = s.perpinv2;
  double t_remain = t_step, t_pass;
  bool boole_t_finished = false;
  int ind_tetr_save = *ind_tetr, iper;
  for (;;) {
    if (*ind_tetr == -1) {
      if (t_remain_out) *t_remain_out = t_remain;
      break;
    }
    ind_tetr_save = *ind_tetr;
    const double vpar_save = *vpar;
    int it = *ind_tetr, ifc = *iface;
    if (m->ipusher == 1)
      pusher_tetra_rk(&rk, &it, &ifc, x, vpar, z_save, t_remain, &t_pass, &boole_t_finished, &iper, tr);
    else
    {
      double optq[4];
      pusher_tetra_poly(&s, m->poly_order, &it, &ifc, x, vpar, z_save, t_remain, &t_pass, &boole_t_finished, &iper, optq);
      if (tr) /* a caller of the pusher sums the per-push values along the orbit */
        for (int q = 0; q < 4; q++) tr->optional_quantities[q] = tr->optional_quantities[q] + optq[q];
    }
    *ind_tetr = it;
    *iface = ifc;
    if (tr) {
      if (tr->n_pushes < tr->cap) {
        tr->ind_tetr[tr->n_pushes] = it;
        tr->iface[tr->n_pushes] = ifc;
      }
      tr->n_pushes++;
    }
    t_remain = t_remain - t_pass;
    if (es) {
      es->t = t_step - t_remain;
      if (es->cfg->boole_full_orbit) { /* gorilla_plot_mod.f90:553-579, before the exit on boole_t_finished */
        const int64_t cnt = es->push + 1, nskip = es->cfg->n_skip_full_orbit; /* counter_tetrahedron_passes */
        if (cnt / nskip * nskip == cnt) {
          double zv[4] = {z_save[0], z_save[1], z_save[2], *vpar};
          emit_event(es, GOR_EVENT_FULL_ORBIT, (int)cnt, x, gor_p_phi(m, *vpar, z_save, ind_tetr_save),
                     gor_energy_tot(m, zv, s.perpinv, ind_tetr_save));
        }
      }
    }
    if (boole_t_finished) {
      if (t_remain_out) *t_remain_out = t_remain;
      break;
    }
    if (es) { /* gorilla_plot_mod.f90:585-638 */
      /* a removed particle (unrecoverable push) is skipped: the reference would evaluate J_par on stale module state */
      if (es->cfg->boole_J_par || es->cfg->boole_poincare_vpar_0) {
        if (m->ipusher == 1) par_adiab_inv_tetra_rk(&rk, t_pass, vpar_save, *vpar, es);
        else if (!s.removed) par_adiab_inv_tetra_poly(&s, m->poly_order, vpar_save, *vpar, es);
      }
      if (iper != 0) {
        es->counter_phi_0_mappings = es->counter_phi_0_mappings + iper;
        if (es->cfg->boole_poincare_phi_0) {
          const int nskip = es->cfg->n_skip_phi_0;
          if (es->counter_phi_0_mappings / nskip * nskip == es->counter_phi_0_mappings) {
            double zv[4] = {z_save[0], z_save[1], z_save[2], *vpar};
            emit_event(es, GOR_EVENT_PHI_0, es->counter_phi_0_mappings, x, gor_p_phi(m, *vpar, z_save, ind_tetr_save),
                       gor_energy_tot(m, zv, s.perpinv, ind_tetr_save));
          }
        }
      }
      es->push++;
    }
  }
  *vperp = vperp_func(m, z_save, s.perpinv, ind_tetr_save);
  free(list_heap);
  return GOR_OK;
}

int64_t gor_orbit_timestep_batch(const gor_mesh *m, int64_t n, double *x, double *vpar, double *vperp,
                                 double t_step, int32_t *boole_initialized, int32_t *ind_tetr, int32_t *iface,
                                 double *t_remain_out, int64_t *n_pushes, int nthreads)
{
  return gor_orbit_timestep_batch_opt(m, n, x, vpar, vperp, t_step, boole_initialized, ind_tetr, iface, t_remain_out,
                                      n_pushes, NULL, nthreads);
}
int64_t gor_orbit_timestep_batch_opt(const gor_mesh *m, int64_t n, double *x, double *vpar, double *vperp,
                                     double t_step, int32_t *boole_initialized, int32_t *ind_tetr, int32_t *iface,
                                     double *t_remain_out, int64_t *n_pushes, double *optional_quantities, int nthreads)
{
  int64_t total = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : total)
  for (int64_t i = 0; i < n; i++) {
    gor_trace tr;
    memset(&tr, 0, sizeof(tr));
    double tro = 0.0;
    gor_orbit_timestep(m, x + 3 * i, vpar + i, vperp + i, t_step, boole_initialized + i, ind_tetr + i,
                       iface + i, &tro, &tr);
    if (t_remain_out) t_remain_out[i] = tro;
    if (n_pushes) n_pushes[i] = tr.n_pushes;
    if (optional_quantities)
      for (int q = 0; q < 4; q++) optional_quantities[4 * i + q] = tr.optional_quantities[q];
    total += tr.n_pushes;
  }
  return total;
}

/* ------------------------------------------------------------------------------------------------
 * Probes of the real gcc (-fcx-fortran-rules) complex lowering and of glibc cabs/csqrt, used to pin
 * the device's explicit restatements (gorilla_b200/csrc/gb_math.cuh) bit for bit.
 * ---------------------------------------------------------------------------------------------- */
void gor_probe_csqrt(double re, double im, double out[2])
{
  cplx r = csqrt(CMPLX(re, im));
  out[0] = creal(r);
  out[1] = cimag(r);
}
double gor_probe_cabs(double re, double im) { return cabs(CMPLX(re, im)); }
void gor_probe_cdiv(double ar, double ai, double br, double bi, double out[2])
{
  volatile cplx a = CMPLX(ar, ai), b = CMPLX(br, bi);
  cplx r = a / b;
  out[0] = creal(r);
  out[1] = cimag(r);
}
void gor_probe_cmul(double ar, double ai, double br, double bi, double out[2])
{
  volatile cplx a = CMPLX(ar, ai), b = CMPLX(br, bi);
  cplx r = a * b;
  out[0] = creal(r);
  out[1] = cimag(r);
}
void gor_probe_rmul(double r0, double br, double bi, double out[2])
{
  volatile double rr = r0;
  volatile cplx b = CMPLX(br, bi);
  cplx r = rmul(rr, b);
  out[0] = creal(r);
  out[1] = cimag(r);
}
