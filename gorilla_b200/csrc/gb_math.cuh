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

// The three routines below are the expensive ones (FP64 divisions and square roots).  They are written
// branch free -- both variants of every data-dependent case are formed with cheap multiplies and the operands
// of the single expensive division / square root are SELECTED -- so that lanes whose data fall into different
// cases still execute the expensive instructions together, and they are compiled as out-of-line functions so
// that the root-solver loop stays small enough for the instruction cache.
#if defined(__CUDACC__)
#define GB_MATH_FN static __host__ __device__ __noinline__
#else
#define GB_MATH_FN static inline
#endif

// IEEE division whose zero-numerator case is handled by selection.  Root iterations on the real axis carry
// exactly-zero imaginary parts, so two of the three divisions of every complex division are 0/x; the GPU's
// double-precision division takes its slow path for a zero (or subnormal) numerator, which would serialise
// those lanes.  (+-0)/x = +-0 with the sign product, for finite non-zero x: exactly what is returned here.
GB_HD double div_z(double num, double den)
{
  const bool zero_num = (num == 0.0) && (fabs(den) <= DBL_MAX) && (den != 0.0);
  const double q = (zero_num ? 1.0 : num) / den;
  return zero_num ? copysign(0.0, num) * copysign(1.0, den) : q;
}

// gcc expand_complex_div_wide (flag_complex_method == 1): branch on |br| < |bi|
//   true : ratio = br/bi; div = br*ratio + bi; tr = ar*ratio + ai; ti = ai*ratio - ar
//   false: ratio = bi/br; div = bi*ratio + br; tr = ai*ratio + ar; ti = ai - ar*ratio
GB_MATH_FN cd cdiv(cd a, cd b)
{
  const bool sw = fabs(b.re) < fabs(b.im);
  const double num = sw ? b.re : b.im, den = sw ? b.im : b.re;
  const double ratio = div_z(num, den);
  const double div = (num * ratio) + den;
  const double u = sw ? a.re : a.im, v = sw ? a.im : a.re;
  const double tr = (u * ratio) + v;
  const double w = sw ? a.im : a.re;  // the factor multiplied by ratio in ti
  const double pw = w * ratio;
  const double ti = sw ? (pw - a.re) : (a.im - pw);
  cd q;
  q.re = div_z(tr, div);
  q.im = div_z(ti, div);
  return q;
}

// glibc 2.35+ __hypot, generic (non-FMA) kernel:  h = sqrt(ax^2+ay^2) followed by one correction step
GB_HD double hypot_kernel(double ax, double ay)
{
  double h = sqrt(ax * ax + ay * ay);
  const bool near = h <= 2.0 * ay;
  const double delta = h - (near ? ay : ax);
  // near: t1 = ax*(2 delta - ax),      t2 = (delta - 2(ax-ay))*delta
  // far : t1 = 2 delta*(ax - 2 ay),    t2 = (4 delta - ay)*ay + delta*delta
  const double t1 = near ? ax * (2.0 * delta - ax) : 2.0 * delta * (ax - 2.0 * ay);
  const double t2 = near ? (delta - 2.0 * (ax - ay)) * delta : (4.0 * delta - ay) * ay + delta * delta;
  h -= (t1 + t2) / (2.0 * h);
  return h;
}
GB_MATH_FN double hypot_glibc(double x, double y)
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
    if (ax >= ay * 0x1p+54) return ax + ay;  // ay / HEPS, exact (power of two); avoids a 0/x division
    ax = hypot_kernel(ax / SCALE, ay / SCALE) * SCALE;
    return ax;
  }
  if (ay <= ax * HEPS) return ax + ay;
  return hypot_kernel(ax, ay);
}
GB_HD double cabs_glibc(cd z) { return hypot_glibc(z.re, z.im); }

// glibc __csqrt for finite arguments with |re|,|im| in [2*DBL_MIN, DBL_MAX/4] (the scaling branches
// for the extreme ranges are not restated: the reference build traps on overflow long before).
GB_MATH_FN cd csqrt_glibc(cd x)
{
  if (!(fabs(x.re) <= DBL_MAX) || !(fabs(x.im) <= DBL_MAX)) {
    double n = x.re - x.re + (x.im - x.im); // NaN
    return mk(n, n);
  }
  // glibc distinguishes: Im == 0 (result on an axis), Re == 0, and the general case
  //   r = sqrt(0.5*(|z| + Re)), s = 0.5*Im/r   (Re > 0)      s = sqrt(0.5*(|z| - Re)), r = |0.5*Im/s|   (Re < 0)
  // All cases take ONE square root of a selected argument; |z| - Re == |z| + |Re| bit for bit when Re < 0.
  const bool im0 = x.im == 0.0, re0 = x.re == 0.0;
  const bool general = !im0 && !re0;
  const double are = fabs(x.re), aim = fabs(x.im);
  const bool tiny = aim < 2.0 * DBL_MIN;
  double d = 0.0;
  if (general) d = hypot_glibc(x.re, x.im);
  const double arg = im0 ? are : (re0 ? (tiny ? 2.0 * aim : 0.5 * aim) : 0.5 * (d + are));
  const double t = sqrt(arg);
  cd res;
  if (im0) {
    res.re = (x.re < 0.0) ? 0.0 : fabs(t);
    res.im = (x.re < 0.0) ? copysign(t, x.im) : copysign(0.0, x.im);
  } else if (re0) {
    const double r = tiny ? 0.5 * t : t;
    res.re = r;
    res.im = copysign(r, x.im);
  } else {
    const double u = 0.5 * (x.im / t);
    res.re = (x.re > 0.0) ? t : fabs(u);
    res.im = copysign((x.re > 0.0) ? u : t, x.im);
  }
  return res;
}

} // namespace gb
