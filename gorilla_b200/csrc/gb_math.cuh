// gb_math.cuh -- FP64 complex arithmetic with a fixed, explicit operation order.
//
// The reference's root solver (SRC/contrib/cmplx_roots_sg.f90) is written with Fortran COMPLEX(8)
// arithmetic.  gfortran lowers it as follows, and the tetra sequence is only reproducible if the
// device does exactly the same operations in the same order:
//   * complex*complex : (ar*br - ai*bi, ar*bi + ai*br)                   (-fcx-fortran-rules)
//   * complex/complex : Smith-type "wide" division, branch on |br| < |bi| (tree-complex.cc)
//   * real*complex    : the real is promoted to (r, +0.0) and a FULL complex product is formed
//                       (signed zeros are honoured, so the 0.0*x terms are not folded away)
//   * abs(complex)    : libm cabs  -> glibc __hypot  (sysdeps/ieee754/dbl-64/e_hypot.c, non-FMA kernel)
//   * sqrt(complex)   : libm csqrt -> glibc __csqrt  (math/s_csqrt_template.c)
// tests/test_device_math_host.py compiles this header for the host and checks every routine
// bit-for-bit against gcc -fcx-fortran-rules / glibc on 10^6 random operands.
//
// Everything here must be compiled without FMA contraction (nvcc --fmad=false, g++ -ffp-contract=off).
#pragma once
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define GB_HD __host__ __device__ __forceinline__
#define GB_HD_NOINLINE __host__ __device__ __noinline__
#else
#define GB_HD inline
#define GB_HD_NOINLINE
#endif

namespace gb {

struct cd {
  double re, im;
};

GB_HD cd mk(double re, double im) { cd r; r.re = re; r.im = im; return r; }
GB_HD cd cadd(cd a, cd b) { return mk(a.re + b.re, a.im + b.im); }
GB_HD cd csub(cd a, cd b) { return mk(a.re - b.re, a.im - b.im); }
GB_HD cd cneg(cd a) { return mk(-a.re, -a.im); }
GB_HD cd cmul(cd a, cd b) { return mk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
// Fortran real*complex: promote r to (r, 0.0), full product
GB_HD cd rmul(double r, cd b) { return mk(r * b.re - 0.0 * b.im, r * b.im + 0.0 * b.re); }
GB_HD bool ceq(cd a, cd b) { return a.re == b.re && a.im == b.im; }
GB_HD bool cis0(cd a) { return a.re == 0.0 && a.im == 0.0; }
// real(conjg(p)*p)
GB_HD double cabs2(cd p) { return p.re * p.re - (-p.im) * p.im; }

// gcc expand_complex_div_wide (flag_complex_method == 1)
GB_HD cd cdiv(cd a, cd b)
{
  cd q;
  if (fabs(b.re) < fabs(b.im)) {
    double ratio = b.re / b.im;
    double div = (b.re * ratio) + b.im;
    double tr = (a.re * ratio) + a.im;
    double ti = (a.im * ratio) - a.re;
    q.re = tr / div;
    q.im = ti / div;
  } else {
    double ratio = b.im / b.re;
    double div = (b.im * ratio) + b.re;
    double tr = (a.im * ratio) + a.re;
    double ti = a.im - (a.re * ratio);
    q.re = tr / div;
    q.im = ti / div;
  }
  return q;
}

// glibc 2.35+ __hypot, generic (non-FMA) kernel
GB_HD double hypot_kernel(double ax, double ay)
{
  double t1, t2;
  double h = sqrt(ax * ax + ay * ay);
  if (h <= 2.0 * ay) {
    double delta = h - ay;
    t1 = ax * (2.0 * delta - ax);
    t2 = (delta - 2.0 * (ax - ay)) * delta;
  } else {
    double delta = h - ax;
    t1 = 2.0 * delta * (ax - 2.0 * ay);
    t2 = (4.0 * delta - ay) * ay + delta * delta;
  }
  h -= (t1 + t2) / (2.0 * h);
  return h;
}
GB_HD double hypot_glibc(double x, double y)
{
  const double SCALE = 0x1p-600, LARGE_VAL = 0x1p+511, TINY_VAL = 0x1p-459, HEPS = 0x1p-54;
  if (!(fabs(x) <= DBL_MAX) || !(fabs(y) <= DBL_MAX)) {
    if (isinf(x) || isinf(y)) return INFINITY;
    return x + y; // NaN
  }
  x = fabs(x);
  y = fabs(y);
  double ax = x < y ? y : x;
  double ay = x < y ? x : y;
  if (ax > LARGE_VAL) {
    if (ay <= ax * HEPS) return ax + ay;
    return hypot_kernel(ax * SCALE, ay * SCALE) / SCALE;
  }
  if (ay < TINY_VAL) {
    if (ax >= ay / HEPS) return ax + ay;
    ax = hypot_kernel(ax / SCALE, ay / SCALE) * SCALE;
    return ax;
  }
  if (ay <= ax * HEPS) return ax + ay;
  return hypot_kernel(ax, ay);
}
GB_HD double cabs_glibc(cd z) { return hypot_glibc(z.re, z.im); }

// glibc __csqrt for finite arguments with |re|,|im| in [2*DBL_MIN, DBL_MAX/4] (the scaling branches
// for the extreme ranges are not restated: the reference build traps on overflow long before).
GB_HD cd csqrt_glibc(cd x)
{
  cd res;
  if (!(fabs(x.re) <= DBL_MAX) || !(fabs(x.im) <= DBL_MAX)) {
    double n = x.re - x.re + (x.im - x.im); // NaN
    return mk(n, n);
  }
  if (x.im == 0.0) {
    if (x.re < 0.0) {
      res.re = 0.0;
      res.im = copysign(sqrt(-x.re), x.im);
    } else {
      res.re = fabs(sqrt(x.re));
      res.im = copysign(0.0, x.im);
    }
  } else if (x.re == 0.0) {
    double r;
    if (fabs(x.im) >= 2.0 * DBL_MIN)
      r = sqrt(0.5 * fabs(x.im));
    else
      r = 0.5 * sqrt(2.0 * fabs(x.im));
    res.re = r;
    res.im = copysign(r, x.im);
  } else {
    double d = hypot_glibc(x.re, x.im), r, s;
    if (x.re > 0.0) {
      r = sqrt(0.5 * (d + x.re));
      s = 0.5 * (x.im / r);
    } else {
      s = sqrt(0.5 * (d - x.re));
      r = fabs(0.5 * (x.im / s));
    }
    res.re = r;
    res.im = copysign(s, x.im);
  }
  return res;
}

} // namespace gb
