// gb_roots.cuh -- exit-time root solvers of the polynomial pusher, FP64, fixed operation order.
//
// Replaces (reference file:line):
//   Linear_Solver / Quadratic_Solver1 / Quadratic_Solver2 / Cubic_Solver / Quartic_Solver
//                                   SRC/pusher_tetra_poly.f90:1767-2021
//   quadraticRoots/cubicRoots/quarticRoots + pack_roots
//                                   SRC/contrib/Polynomial234RootSolvers.f90:36-142
//   cmplx_roots_gen (polish=.true., start=.false.), cmplx_laguerre, cmplx_laguerre2newton,
//   solve_quadratic_eq              SRC/contrib/cmplx_roots_sg.f90:83-201,558-731,906-1305,1311-1362
//
// The Laguerre -> SG -> Newton iteration is data dependent; each lane runs its own state machine and
// the warp reconverges after the solve.  Degree is a template parameter so that the Horner loops and
// the deflation are fully unrolled and the polynomial lives in registers.
#pragma once
#include "gb_math.cuh"

namespace gb {

#define GB_HUGE DBL_MAX

// exp(cmplx(0, FRAC_JUMPS(k+1)*2*pi)) exactly as glibc cexp() returns it (generated with
// oracle/gor_frac_jump_phase; checked again at test time, tests/test_device_math_host.py)
GB_HD cd frac_jump_phase(int k)
{
  switch (k) {
    case 0: return mk(-0x1.43a4ec067471ap-1, -0x1.8cbc14adec55cp-1);
    case 1: return mk(0x1.b9f8520125e44p-1, -0x1.027835a80b001p-1);
    case 2: return mk(-0x1.d9f09317910bbp-5, 0x1.ff24763655658p-1);
    case 3: return mk(-0x1.ffc29786c1b25p-1, -0x1.f5776cee4e213p-6);
    case 4: return mk(0x1.bde5c12533376p-1, 0x1.f7445f6e48663p-2);
    case 5: return mk(0x1.4ee0d0497a649p-1, 0x1.834c92e8b2a3ep-1);
    case 6: return mk(-0x1.61e482927f6dbp-2, 0x1.e073ae0e62aa1p-1);
    case 7: return mk(-0x1.70abece2c21e4p-1, 0x1.63482dd7978b3p-1);
    case 8: return mk(0x1.ea9723c588b18p-1, 0x1.2504ed506e352p-2);
    default: return mk(0x1.100ce2673d75fp-7, -0x1.fffb7b8d5ef39p-1);
  }
}
GB_HD double frac_jump(int k)
{
  switch (k) {
    case 0: return 0.64109297;
    case 1: return 0.91577881;
    case 2: return 0.25921289;
    case 3: return 0.50487203;
    case 4: return 0.08177045;
    case 5: return 0.13653241;
    case 6: return 0.306162;
    case 7: return 0.37794326;
    case 8: return 0.04618805;
    default: return 0.75132137;
  }
}

#define GB_FRAC_ERR 2.0e-15

// Horner evaluation of p, p', p''/2 (+ Adams' error bound) -- the common body of every mode
template <int DEG, bool WITH_D2, bool WITH_EK>
GB_HD void horner(const cd *poly, cd root, cd &p, cd &dp, cd &d2p_half, double &ek)
{
  double absroot = 0.0;
  if (WITH_EK) {
    ek = cabs_glibc(poly[DEG]);
    absroot = cabs_glibc(root);
  }
  p = poly[DEG];
  dp = mk(0.0, 0.0);
  d2p_half = mk(0.0, 0.0);
#pragma unroll
  for (int k = DEG; k >= 1; k--) {
    if (WITH_D2) d2p_half = cadd(dp, cmul(d2p_half, root));
    dp = cadd(p, cmul(dp, root));
    p = cadd(poly[k - 1], cmul(p, root));
    if (WITH_EK) ek = absroot * ek + cabs_glibc(p);
  }
}

// Laguerre step denominator (shared by cmplx_laguerre and mode 2 of cmplx_laguerre2newton)
template <int DEG>
GB_HD cd laguerre_denom(cd F_half)
{
  const double one_nth = 1.0 / DEG;
  const double n_1_nth = (DEG - 1.0) * one_nth;
  const double two_n_div_n_1 = 2.0 / n_1_nth;
  const cd c_one = mk(1.0, 0.0), c_one_nth = mk(one_nth, 0.0);
  cd denom_sqrt = csqrt_glibc(csub(c_one, rmul(two_n_div_n_1, F_half)));
  if (denom_sqrt.re >= 0.0) return cadd(c_one_nth, rmul(n_1_nth, denom_sqrt));
  return csub(c_one_nth, rmul(n_1_nth, denom_sqrt));
}

// cmplx_roots_sg.f90:558-731
template <int DEG>
GB_HD_NOINLINE bool cmplx_laguerre(const cd *poly, cd &root, int &iters)
{
  const int MAX_ITERS = 200;
  bool good_to_go = false;
  for (int i = 1; i <= MAX_ITERS; i++) {
    cd p, dp, d2p_half, fac_netwon = mk(0.0, 0.0);
    double ek;
    double absroot = cabs_glibc(root);
    horner<DEG, true, true>(poly, root, p, dp, d2p_half, ek);
    iters++;
    double abs2p = cabs2(p);
    if (abs2p == 0.0) return true;
    double sc = GB_FRAC_ERR * ek;
    double stopping_crit2 = sc * sc;
    if (abs2p < stopping_crit2) {
      if (abs2p < 0.01 * stopping_crit2) return true;
      good_to_go = true;
    } else {
      good_to_go = false;
    }
    cd denom = mk(0.0, 0.0);
    if (!cis0(dp)) {
      fac_netwon = cdiv(p, dp);
      cd fac_extra = cdiv(d2p_half, dp);
      cd F_half = cmul(fac_netwon, fac_extra);
      denom = laguerre_denom<DEG>(F_half);
    }
    cd dx;
    if (cis0(denom))
      dx = rmul(absroot + 1.0, frac_jump_phase(i % 10));
    else
      dx = cdiv(fac_netwon, denom);
    cd newroot = csub(root, dx);
    if (ceq(newroot, root)) return true;
    if (good_to_go) {
      root = newroot;
      return true;
    }
    if (i % 10 == 0) {
      double faq = frac_jump((i / 10 - 1) % 10);
      newroot = csub(root, rmul(faq, dx));
    }
    root = newroot;
  }
  return false;
}

// cmplx_roots_sg.f90:906-1305, starting_mode = 2
template <int DEG>
GB_HD_NOINLINE bool cmplx_laguerre2newton(const cd *poly, cd &root, int &iters)
{
  const int MAX_ITERS = 50;
  const cd c_one = mk(1.0, 0.0);
  int mode = 2, i, j = 1, iter = 0;
  bool good_to_go = false;
  double stopping_crit2 = 0.0;
  for (;;) {
    if (mode >= 2) {
      for (i = 1; i <= MAX_ITERS; i++) {
        cd p, dp, d2p_half, fac_netwon = mk(0.0, 0.0);
        double ek;
        horner<DEG, true, true>(poly, root, p, dp, d2p_half, ek);
        double abs2p = cabs2(p);
        iter++;
        if (abs2p == 0.0) { iters += iter; return true; }
        double sc = GB_FRAC_ERR * ek;
        stopping_crit2 = sc * sc;
        if (abs2p < stopping_crit2) {
          if (abs2p < 0.01 * stopping_crit2) { iters += iter; return true; }
          good_to_go = true;
        } else {
          good_to_go = false;
        }
        cd denom = mk(0.0, 0.0);
        if (!cis0(dp)) {
          fac_netwon = cdiv(p, dp);
          cd fac_extra = cdiv(d2p_half, dp);
          cd F_half = cmul(fac_netwon, fac_extra);
          double abs2_F_half = cabs2(F_half);
          if (abs2_F_half <= 0.0625) {
            if (abs2_F_half <= 0.000625)
              mode = 0;
            else
              mode = 1;
          }
          denom = laguerre_denom<DEG>(F_half);
        }
        cd dx;
        if (cis0(denom))
          dx = rmul(cabs_glibc(root) + 1.0, frac_jump_phase(i % 10));
        else
          dx = cdiv(fac_netwon, denom);
        cd newroot = csub(root, dx);
        if (ceq(newroot, root)) { iters += iter; return true; }
        if (good_to_go) {
          root = newroot;
          iters += iter;
          return true;
        }
        if (mode != 2) {
          root = newroot;
          j = i + 1;
          break;
        }
        if (i % 10 == 0) {
          double faq = frac_jump((i / 10 - 1) % 10);
          newroot = csub(root, rmul(faq, dx));
        }
        root = newroot;
      }
      if (i >= MAX_ITERS) { iters += iter; return false; }
    }
    if (mode == 1) {
      for (i = j; i <= MAX_ITERS; i++) {
        cd p, dp, d2p_half;
        double ek;
        if ((i - j) % 10 == 0) {
          horner<DEG, true, true>(poly, root, p, dp, d2p_half, ek);
          double sc = GB_FRAC_ERR * ek;
          stopping_crit2 = sc * sc;
        } else {
          horner<DEG, true, false>(poly, root, p, dp, d2p_half, ek);
        }
        double abs2p = cabs2(p);
        iter++;
        if (abs2p == 0.0) { iters += iter; return true; }
        if (abs2p < stopping_crit2) {
          if (cis0(dp)) { iters += iter; return true; }
          if (abs2p < 0.01 * stopping_crit2) { iters += iter; return true; }
          good_to_go = true;
        } else {
          good_to_go = false;
        }
        cd dx;
        if (cis0(dp)) {
          dx = rmul(cabs_glibc(root) + 1.0, frac_jump_phase(i % 10));
        } else {
          cd fac_netwon = cdiv(p, dp);
          cd fac_extra = cdiv(d2p_half, dp);
          cd F_half = cmul(fac_netwon, fac_extra);
          double abs2_F_half = cabs2(F_half);
          if (abs2_F_half <= 0.000625) mode = 0;
          dx = cmul(fac_netwon, cadd(c_one, F_half));
        }
        cd newroot = csub(root, dx);
        if (ceq(newroot, root)) { iters += iter; return true; }
        if (good_to_go) {
          root = newroot;
          iters += iter;
          return true;
        }
        if (mode != 1) {
          root = newroot;
          j = i + 1;
          break;
        }
        if (i % 10 == 0) {
          double faq = frac_jump((i / 10 - 1) % 10);
          newroot = csub(root, rmul(faq, dx));
        }
        root = newroot;
      }
      if (i >= MAX_ITERS) { iters += iter; return false; }
    }
    if (mode == 0) {
      for (i = j; i <= j + 10; i++) {
        cd p, dp, d2p_half;
        double ek;
        if (i == j) {
          horner<DEG, false, true>(poly, root, p, dp, d2p_half, ek);
          double sc = GB_FRAC_ERR * ek;
          stopping_crit2 = sc * sc;
        } else {
          horner<DEG, false, false>(poly, root, p, dp, d2p_half, ek);
        }
        double abs2p = cabs2(p);
        iter++;
        if (abs2p == 0.0) { iters += iter; return true; }
        if (abs2p < stopping_crit2) {
          if (cis0(dp)) { iters += iter; return true; }
          if (abs2p < 0.01 * stopping_crit2) { iters += iter; return true; }
          good_to_go = true;
        } else {
          good_to_go = false;
        }
        cd dx;
        if (cis0(dp))
          dx = rmul(cabs_glibc(root) + 1.0, frac_jump_phase(i % 10));
        else
          dx = cdiv(p, dp);
        cd newroot = csub(root, dx);
        if (ceq(newroot, root)) { iters += iter; return true; }
        if (good_to_go) {
          root = newroot;
          iters += iter;
          return true;
        }
        root = newroot;
      }
      if (iter >= MAX_ITERS) { iters += iter; return false; }
      mode = 2;
    }
  }
}

// cmplx_roots_sg.f90:1311-1362
GB_HD void solve_quadratic_eq(cd &x0, cd &x1, const cd *poly)
{
  cd a = poly[2], b = poly[1], c = poly[0];
  cd b2 = cmul(b, b);
  cd delta = csqrt_glibc(csub(b2, rmul(4.0, cmul(a, c))));
  cd conjb = mk(b.re, -b.im);
  if (cmul(conjb, delta).re >= 0.0)
    x0 = rmul(-0.5, cadd(b, delta));
  else
    x0 = rmul(-0.5, csub(b, delta));
  if (cis0(x0)) {
    x1 = mk(0.0, 0.0);
  } else {
    x1 = cdiv(c, x0);
    x0 = cdiv(x0, a);
  }
}

// One deflation stage: find a root of the degree-N working polynomial, divide it out.
template <int N>
GB_HD void find_and_deflate(cd *poly2, cd *roots, int &iters)
{
  roots[N - 1] = mk(0.0, 0.0);
  if (!cmplx_laguerre2newton<N>(poly2, roots[N - 1], iters)) {
    roots[N - 1] = mk(0.0, 0.0);
    cmplx_laguerre<N>(poly2, roots[N - 1], iters);
  }
  cd coef = poly2[N];
#pragma unroll
  for (int i = N; i >= 1; i--) {
    cd prev = poly2[i - 1];
    poly2[i - 1] = coef;
    coef = cadd(prev, cmul(roots[N - 1], coef));
  }
}

// cmplx_roots_gen(roots, poly, DEG, .true., .false.)   cmplx_roots_sg.f90:83-201
template <int DEG>
GB_HD void cmplx_roots_gen(cd *roots, const cd *poly, int &iters)
{
  cd poly2[DEG + 1];
#pragma unroll
  for (int i = 0; i <= DEG; i++) poly2[i] = poly[i];
  if constexpr (DEG >= 4) find_and_deflate<4>(poly2, roots, iters);
  if constexpr (DEG >= 3) find_and_deflate<3>(poly2, roots, iters);
  roots[1] = mk(0.0, 0.0);
  roots[0] = mk(0.0, 0.0);
  if (!cmplx_laguerre2newton<2>(poly2, roots[1], iters)) {
    solve_quadratic_eq(roots[1], roots[0], poly2);
  } else {
    roots[0] = cneg(cadd(roots[1], cdiv(poly2[1], poly2[2])));
  }
#pragma unroll
  for (int n = 0; n < DEG; n++) cmplx_laguerre<DEG>(poly, roots[n], iters);
}

// pack_roots + the callers' "smallest positive real root / lambda" reduction, fused:
// a root counts as real when |Im| <= 1e-12*max(1,|Re|) (then Im := 0 exactly); the callers divide by
// lambda and take minval over {Im == 0, Re > 0}; empty mask -> huge.  The descending sort of
// pack_roots does not change a minimum, so it is not materialised.
template <int DEG>
GB_HD double min_positive_real_root(const cd *croots, double lambda)
{
  double best = GB_HUGE;
#pragma unroll
  for (int i = 0; i < DEG; i++) {
    double re = croots[i].re, im = croots[i].im;
    double tol_i = 1.0e-12 * fmax(1.0, fabs(re));
    if (fabs(im) <= tol_i) im = 0.0;
    re = re / lambda;
    im = im / lambda;
    if (fabs(im) == 0.0 && re > 0.0 && re < best) best = re;
  }
  return best;
}

template <int DEG>
GB_HD double solve_monic_min_positive(const double *q /* q[0]=const .. q[DEG-1] */, double lambda, int &iters)
{
  cd poly[DEG + 1], roots[DEG];
#pragma unroll
  for (int i = 0; i < DEG; i++) poly[i] = mk(q[i], 0.0);
  poly[DEG] = mk(1.0, 0.0);
  cmplx_roots_gen<DEG>(roots, poly, iters);
  return min_positive_real_root<DEG>(roots, lambda);
}

// ---- SRC/pusher_tetra_poly.f90:1767-2021 -----------------------------------------------------------
GB_HD double linear_solver(double a, double b)
{
  if (a == 0.0) return GB_HUGE;
  return -b / a;
}

// f(tau) = a/2 tau^2 + b tau + c, smallest positive root by sign-case analysis (:1809-1893)
GB_HD double quadratic_solver1(double acoef, double bcoef, double ccoef)
{
  const double eps = 1.e-10;
  double dtau = GB_HUGE;
  if (ccoef > 0.0) {
    if (acoef > 0.0) {
      if (bcoef < 0.0) {
        double discr = bcoef * bcoef - 2.0 * acoef * ccoef;
        if (discr > 0.0) {
          double dummy = (-bcoef + sqrt(discr));
          if (fabs(dummy) > eps)
            dtau = 2.0 * ccoef / dummy;
          else
            dtau = (-sqrt(discr) - bcoef) / acoef;
        } else if (discr == 0.0) {
          dtau = -bcoef / acoef;
        }
      }
    } else if (acoef < 0.0) {
      double discr = bcoef * bcoef - 2.0 * acoef * ccoef;
      double dummy = (-bcoef + sqrt(discr));
      if (fabs(dummy) > eps)
        dtau = 2.0 * ccoef / dummy;
      else
        dtau = (-sqrt(discr) - bcoef) / acoef;
    } else {
      if (bcoef < 0.0) dtau = -ccoef / bcoef;
    }
  } else if (ccoef < 0.0) {
    if (acoef < 0.0) {
      if (bcoef > 0.0) {
        double discr = bcoef * bcoef - 2.0 * acoef * ccoef;
        if (discr > 0.0)
          dtau = (sqrt(discr) - bcoef) / acoef;
        else if (discr == 0.0)
          dtau = -bcoef / acoef;
      }
    } else if (acoef > 0.0) {
      double discr = bcoef * bcoef - 2.0 * acoef * ccoef;
      dtau = (sqrt(discr) - bcoef) / acoef;
    } else {
      if (bcoef > 0.0) dtau = -ccoef / bcoef;
    }
  } else {
    if (((acoef > 0.0) && (bcoef < 0.0)) || ((acoef < 0.0) && (bcoef > 0.0))) dtau = -2.0 * bcoef / acoef;
  }
  return dtau;
}

static GB_HD_NOINLINE double quadratic_solver2(double a, double b, double c, int &iters)
{
  double lambda = b / c;
  double q[2];
  q[0] = 2.0 * (b * b) / (a * c);
  q[1] = q[0];
  return solve_monic_min_positive<2>(q, lambda, iters);
}

static GB_HD_NOINLINE double cubic_solver(double a, double b, double c, double d, int &iters)
{
  double lambda = b / (2.0 * c);
  double l2 = lambda * lambda;
  double q[3];
  q[2] = 3.0 * lambda * b / a;
  q[1] = 6.0 * c * l2 / a;
  q[0] = 6.0 * d * (l2 * lambda) / a;
  return solve_monic_min_positive<3>(q, lambda, iters);
}

static GB_HD_NOINLINE double quartic_solver(int i_scaling, double a, double b, double c, double d, double e, int &iters)
{
  double lambda;
  switch (i_scaling) {
    case 0: lambda = sqrt(fabs(b / (6.0 * d))); break;
    case 1: lambda = b / (3.0 * c); break;
    case 2: lambda = pow(fabs(b / (6.0 * e)), 1.0 / 3.0); break; // libm pow: last-bit parity not guaranteed
    case 3: lambda = c / (2.0 * d); break;
    case 4: lambda = sqrt(fabs(c / (2.0 * e))); break;
    case 5: lambda = d / e; break;
    default: lambda = pow(fabs(a / (24.0 * e)), 1.0 / 4.0); break;
  }
  double l2 = lambda * lambda;
  double q[4];
  q[3] = 4.0 * b * lambda / a;
  q[2] = 12.0 * c * l2 / a;
  q[1] = 24.0 * d * (l2 * lambda) / a;
  q[0] = 24.0 * e * (l2 * l2) / a;
  return solve_monic_min_positive<4>(q, lambda, iters);
}

} // namespace gb
