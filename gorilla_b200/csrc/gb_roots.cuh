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

// ----------------------------------------------------------------------------------------------------
// cmplx_roots_gen(roots, poly, deg, polish=.true., start=.false.) as ONE per-lane state machine.
//
// The reference runs, for a degree-d polynomial, d-1 root searches on successively deflated polynomials
// (cmplx_laguerre2newton: Laguerre -> SG -> Newton modes, with plain Laguerre as a fall-back) and then d
// Laguerre polishes on the original polynomial (cmplx_roots_sg.f90:146-198).  Written as nested loops, SIMT
// lanes would wait for each other at the end of every search, every mode and every polish.  Here each lane
// carries (stage, method, mode, i, j, ...) explicitly and every trip of the single loop below performs one
// iteration of whatever that lane is doing: all lanes share the same Horner evaluation (degree and
// coefficients are per-lane data, indexed statically with predicates), so a warp needs max-over-lanes of the
// TOTAL iteration count instead of the sum over stages of the per-stage maxima.  The arithmetic each lane
// performs is exactly the reference's sequence of operations (verified bit-for-bit against the oracle).
// ----------------------------------------------------------------------------------------------------
struct SgLaguerreConsts {
  double one_nth, n_1_nth, two_n_div_n_1;
};
GB_HD SgLaguerreConsts sg_consts(int degree)
{
  SgLaguerreConsts c;
  c.one_nth = 1.0 / degree;
  c.n_1_nth = (degree - 1.0) * c.one_nth;
  c.two_n_div_n_1 = 2.0 / c.n_1_nth;
  return c;
}

// One lane's solve as a resumable object: start() loads the polynomial, every step() performs one trip of the
// loop described above and returns true once all roots are final.  Keeping it resumable lets a kernel hand a lane
// the next polynomial the moment it finishes one (work queue), so lanes never idle behind slower neighbours.
struct SgSolver {
  cd poly[5], work[5], roots[4], root;
  int deg, n, phase, pol, mode, i, j, iter;
  bool good_to_go, done;
  double stopping_crit2;

  // poly_in: deg+1 coefficients (poly_in[0] constant term), deg in {2,3,4}
  GB_HD void start(int deg_, const cd *poly_in)
  {
    const cd zero = mk(0.0, 0.0);
    deg = deg_;
#pragma unroll
    for (int k = 0; k < 5; k++) {
      poly[k] = (k <= deg) ? poly_in[k <= deg ? k : 0] : zero;
      work[k] = poly[k];
    }
#pragma unroll
    for (int k = 0; k < 4; k++) roots[k] = zero;
    n = deg;      // degree of the working polynomial during the search stages
    phase = 0;    // 0: cmplx_laguerre2newton search, 1: cmplx_laguerre fall-back search, 2: polish
    pol = 0;      // root being polished
    root = zero;
    mode = 2; i = 1; j = 1; iter = 0;
    good_to_go = false;
    stopping_crit2 = 0.0;
    done = false;
  }

  GB_HD bool step()
  {
    const cd zero = mk(0.0, 0.0), c_one = mk(1.0, 0.0);
    const bool lag = (phase != 0);
    const int cdeg = (phase == 2) ? deg : n;
    cd c[5];
#pragma unroll
    for (int k = 0; k < 5; k++) c[k] = (phase == 2) ? poly[k] : work[k];
    const int mode0 = mode;  // mode at the start of this iteration
    const bool need_ek = lag || mode0 == 2 || (mode0 == 1 && ((i - j) % 10 == 0)) || (mode0 == 0 && i == j);

    // ---- Horner: p, p', p''/2 and Adams' bound
    cd p = (cdeg == 4) ? c[4] : (cdeg == 3) ? c[3] : c[2];
    cd dp = zero, d2p_half = zero;
    double ek = 0.0, absroot = 0.0;
    if (need_ek) {
      ek = cabs_glibc(p);
      absroot = cabs_glibc(root);
    }
#pragma unroll
    for (int k = 4; k >= 1; k--) {
      if (k <= cdeg) {
        d2p_half = cadd(dp, cmul(d2p_half, root));
        dp = cadd(p, cmul(dp, root));
        p = cadd(c[k - 1], cmul(p, root));
        if (need_ek) ek = absroot * ek + cabs_glibc(p);
      }
    }
    iter++;
        int ret = 0;  // 0 continue, 1 routine returns success, 2 routine returns failure
    const double abs2p = cabs2(p);
    if (abs2p == 0.0) {
      ret = 1;
    } else {
      if (need_ek) {
        const double sc = GB_FRAC_ERR * ek;
        stopping_crit2 = sc * sc;
      }
      if (abs2p < stopping_crit2) {
        if (!lag && mode0 != 2 && cis0(dp)) ret = 1;
        else if (abs2p < 0.01 * stopping_crit2) ret = 1;
        else good_to_go = true;
      } else {
        good_to_go = false;
      }
    }
    if (ret == 0) {
      // The step dx.  All modes divide p by p'; Laguerre and SG also need F/2 = (p/p')(p''/2p'); only
      // Laguerre takes the complex square root.  The common parts are issued once for the whole warp.
      cd dx;
      const bool dp0 = cis0(dp);
      const bool laguerre_step = lag || mode0 == 2;
      cd fac_netwon = zero, F_half = zero;
      if (!dp0) {
        fac_netwon = cdiv(p, dp);
        if (laguerre_step || mode0 == 1) {
          const cd fac_extra = cdiv(d2p_half, dp);
          F_half = cmul(fac_netwon, fac_extra);
          if (!lag) {
            const double abs2_F_half = cabs2(F_half);
            if (mode0 == 2) {
              if (abs2_F_half <= 0.0625) mode = (abs2_F_half <= 0.000625) ? 0 : 1;
            } else {
              if (abs2_F_half <= 0.000625) mode = 0;
            }
          }
        }
      }
      bool jump = dp0;  // random jump of length |root|+1 (denominator vanishes)
      if (laguerre_step) {
        cd denom = zero;
        if (!dp0) {
          const SgLaguerreConsts k = sg_consts(cdeg);
          const cd denom_sqrt = csqrt_glibc(csub(c_one, rmul(k.two_n_div_n_1, F_half)));
          const cd c_one_nth = mk(k.one_nth, 0.0);
          if (denom_sqrt.re >= 0.0) denom = cadd(c_one_nth, rmul(k.n_1_nth, denom_sqrt));
          else denom = csub(c_one_nth, rmul(k.n_1_nth, denom_sqrt));
        }
        jump = cis0(denom);
        if (!jump) dx = cdiv(fac_netwon, denom);
      } else if (mode0 == 1) {
        if (!jump) dx = cmul(fac_netwon, cadd(c_one, F_half));
      } else {
        if (!jump) dx = fac_netwon;
      }
      if (jump) dx = rmul(cabs_glibc(root) + 1.0, frac_jump_phase(i % 10));
      cd newroot = csub(root, dx);
      if (ceq(newroot, root)) {
        ret = 1;
      } else if (good_to_go) {
        root = newroot;
        ret = 1;
      } else if (lag) {
        if (i % 10 == 0) newroot = csub(root, rmul(frac_jump((i / 10 - 1) % 10), dx));
        root = newroot;
        i++;
        if (i > 200) ret = 2;
      } else if (mode0 == 0) {
        root = newroot;
        i++;
        if (i > j + 10) {
          if (iter >= 50) ret = 2;
          mode = 2;
          i = 1;
        }
      } else if (mode != mode0) {  // leaving Laguerre (2) or SG (1) for a faster mode
        root = newroot;
        j = i + 1;
        if (i >= 50) ret = 2;
        i = j;
      } else {
        if (i % 10 == 0) newroot = csub(root, rmul(frac_jump((i / 10 - 1) % 10), dx));
        root = newroot;
        i++;
        if (i > 50) ret = 2;
      }
    }
    if (ret != 0) {
      // ---- the routine this lane was in has returned: advance the stage machine
      bool start_search = false, start_polish = false;
      if (phase == 2) {
        if (pol == 0) roots[0] = root;
        else if (pol == 1) roots[1] = root;
        else if (pol == 2) roots[2] = root;
        else roots[3] = root;
        pol++;
        if (pol == deg) done = true;
        else start_polish = true;
      } else if (phase == 0 && ret == 2 && n >= 3) {
        // cmplx_laguerre2newton failed: plain Laguerre from (0,0) on the same polynomial (:163-166)
        root = zero;
        phase = 1;
        i = 1;
        good_to_go = false;
      } else if (n >= 3) {
        // root of the working polynomial found: store, divide it out (:169-176)
        if (n == 4) roots[3] = root;
        else roots[2] = root;
        cd coef = (n == 4) ? work[4] : work[3];
#pragma unroll
        for (int k = 4; k >= 1; k--) {
          if (k <= n) {
            const cd prev = work[k - 1];
            work[k - 1] = coef;
            coef = cadd(prev, cmul(root, coef));
          }
        }
        n--;
        start_search = true;
      } else {
        // n == 2: last search (:184-190)
        if (ret == 2) {
          solve_quadratic_eq(roots[1], roots[0], work);
        } else {
          roots[1] = root;
          roots[0] = cneg(cadd(roots[1], cdiv(work[1], work[2])));
        }
        pol = 0;
        start_polish = true;
      }
      if (start_search) {
        phase = 0;
        root = zero;
        mode = 2; i = 1; j = 1; iter = 0;
        good_to_go = false;
        stopping_crit2 = 0.0;
      }
      if (start_polish) {
        phase = 2;
        root = (pol == 0) ? roots[0] : (pol == 1) ? roots[1] : (pol == 2) ? roots[2] : roots[3];
        i = 1;
        good_to_go = false;
      }
    }
    return done;
  }
};

// poly: deg+1 coefficients (poly[0] constant term), roots: deg outputs.  deg in {2,3,4}.
GB_HD void sg_roots(int deg, const cd *poly_in, cd *roots_out, int &iters)
{
  SgSolver s;
  s.start(deg, poly_in);
  while (!s.step()) iters++;
#pragma unroll
  for (int k = 0; k < 4; k++)
    if (k < deg) roots_out[k] = s.roots[k];
}

// pack_roots + the callers' "smallest positive real root / lambda" reduction, fused:
// a root counts as real when |Im| <= 1e-12*max(1,|Re|) (then Im := 0 exactly); the callers divide by
// lambda and take minval over {Im == 0, Re > 0}; empty mask -> huge.  The descending sort of
// pack_roots does not change a minimum, so it is not materialised.
GB_HD double min_positive_real_root(int deg, const cd *croots, double lambda)
{
  double best = GB_HUGE;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    if (i < deg) {
      double re = croots[i].re, im = croots[i].im;
      const double tol_i = 1.0e-12 * fmax(1.0, fabs(re));
      if (fabs(im) <= tol_i) im = 0.0;
      re = re / lambda;
      im = div_z(im, lambda);
      if (fabs(im) == 0.0 && re > 0.0 && re < best) best = re;
    }
  }
  return best;
}

// monic real polynomial x^deg + q[deg-1] x^(deg-1) + ... + q[0]; one shared, non-inlined instance
static GB_HD_NOINLINE double solve_monic_min_positive(int deg, double q0, double q1, double q2, double q3, double lambda,
                                                      int &iters)
{
  cd poly[5], roots[4];
  poly[0] = mk(q0, 0.0);
  poly[1] = mk(deg == 1 ? 1.0 : q1, 0.0);
  poly[2] = mk(deg == 2 ? 1.0 : q2, 0.0);
  poly[3] = mk(deg == 3 ? 1.0 : q3, 0.0);
  poly[4] = mk(1.0, 0.0);
  sg_roots(deg, poly, roots, iters);
  return min_positive_real_root(deg, roots, lambda);
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

// Quadratic_Solver1 rewritten as "choose numerator and denominator, divide once": every branch of the
// reference's sign-case tree (:1827-1891) ends in exactly one division, so the tree only has to SELECT the
// operands.  sqrt and the division are then issued unconditionally by all lanes (no divergence, and the four
// faces of a tetrahedron give four independent chains).  Returns false when no root exists (dtau = huge).
GB_HD bool quadratic_solver1_numden(double a, double b, double c, double &num, double &den)
{
  const double eps = 1.e-10;
  const double discr = b * b - 2.0 * a * c;
  // sqrt of a negative discriminant is never used; feeding 1.0 keeps those lanes on the fast sqrt path
  const double sq = sqrt(discr >= 0.0 ? discr : 1.0);
  const double dummy = (-b + sq);
  const bool big = fabs(dummy) > eps;
  // the sign-case tree as predicates (else-branches of the reference catch NaN: cz, az are "neither > nor <")
  const bool cp = c > 0.0, cn = !cp && (c < 0.0), cz = !cp && !cn;
  const bool ap = a > 0.0, an = !ap && (a < 0.0), az = !ap && !an;
  const bool bn = b < 0.0, bp = b > 0.0;
  const bool dpos = discr > 0.0, dzero = !dpos && (discr == 0.0);
  const bool twoc = (cp && ap && bn && dpos) || (cp && an);          // 2c/dummy or (-sq-b)/a
  const bool psq = (cn && an && bp && dpos) || (cn && ap);           // (sq-b)/a
  const bool mb = (cp && ap && bn && dzero) || (cn && an && bp && dzero);  // -b/a
  const bool mc = (cp && az && bn) || (cn && az && bp);              // -c/b
  const bool m2b = cz && ((ap && bn) || (an && bp));                 // -2b/a
  const bool has = twoc || psq || mb || mc || m2b;
  // operands of the single division; lanes without a root divide 1/1 (fast path, result unused)
  num = twoc ? (big ? 2.0 * c : (-sq - b)) : psq ? (sq - b) : mb ? -b : mc ? -c : m2b ? -2.0 * b : 1.0;
  den = twoc ? (big ? dummy : a) : mc ? b : has ? a : 1.0;
  return has;
}

// Rescaling to a monic polynomial (the part of Quadratic_Solver2 / Cubic_Solver / Quartic_Solver before the root
// solve): q[0..deg-1] coefficients (constant term first), lambda the scale that the roots are divided by.
GB_HD void quadratic2_prepare(double a, double b, double c, double *q, double &lambda)
{
  lambda = b / c;
  q[0] = 2.0 * (b * b) / (a * c);
  q[1] = q[0];
  q[2] = 0.0;
  q[3] = 0.0;
}
GB_HD void cubic_prepare(double a, double b, double c, double d, double *q, double &lambda)
{
  lambda = b / (2.0 * c);
  const double l2 = lambda * lambda;
  q[2] = 3.0 * lambda * b / a;
  q[1] = 6.0 * c * l2 / a;
  q[0] = 6.0 * d * (l2 * lambda) / a;
  q[3] = 0.0;
}
GB_HD void quartic_prepare(int i_scaling, double a, double b, double c, double d, double e, double *q, double &lambda)
{
  switch (i_scaling) {
    case 0: lambda = sqrt(fabs(b / (6.0 * d))); break;
    case 1: lambda = b / (3.0 * c); break;
    case 2: lambda = pow(fabs(b / (6.0 * e)), 1.0 / 3.0); break; // libm pow: last-bit parity not guaranteed
    case 3: lambda = c / (2.0 * d); break;
    case 4: lambda = sqrt(fabs(c / (2.0 * e))); break;
    case 5: lambda = d / e; break;
    default: lambda = pow(fabs(a / (24.0 * e)), 1.0 / 4.0); break;
  }
  const double l2 = lambda * lambda;
  q[3] = 4.0 * b * lambda / a;
  q[2] = 12.0 * c * l2 / a;
  q[1] = 24.0 * d * (l2 * lambda) / a;
  q[0] = 24.0 * e * (l2 * l2) / a;
}

GB_HD double quadratic_solver2(double a, double b, double c, int &iters)
{
  double q[4], lambda;
  quadratic2_prepare(a, b, c, q, lambda);
  return solve_monic_min_positive(2, q[0], q[1], 0.0, 0.0, lambda, iters);
}
GB_HD double cubic_solver(double a, double b, double c, double d, int &iters)
{
  double q[4], lambda;
  cubic_prepare(a, b, c, d, q, lambda);
  return solve_monic_min_positive(3, q[0], q[1], q[2], 0.0, lambda, iters);
}
GB_HD double quartic_solver(int i_scaling, double a, double b, double c, double d, double e, int &iters)
{
  double q[4], lambda;
  quartic_prepare(i_scaling, a, b, c, d, e, q, lambda);
  return solve_monic_min_positive(4, q[0], q[1], q[2], q[3], lambda, iters);
}

} // namespace gb
