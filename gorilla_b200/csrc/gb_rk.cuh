// gb_rk.cuh -- Runge-Kutta (RK4) tetrahedron pusher, FP64, one particle per lane.
//
// Replaces (reference file:line), for boole_pusher_ode45 = .false., boole_dt_dtau = .true.,
// boole_newton_precalc = .false., handover_processing_kind = 1 (2 in the EXT = 2 variant):
//   initialize_pusher_tetra_rk_mod        SRC/pusher_tetra_rk.f90:50-193
//   pusher_tetra_rk                       :197-575
//   quad_analytic_approx                  :636-809
//   rhs_pusher_tetra_rk4 / rk4_step       :810-896      (integration_step :2549-2581 is one rk4_step here)
//   newton_face_convergence(_wrapped)     :914-1188
//   bisection_face_convergence            :1409-1581
//   bisection_search_start                :1583-1623    (streamed: the 1000-step history is not materialised)
//   last_line_defense                     :1696-2071
//   final_processing                      :2075-2418
//   normal_distance(s)_func / normal_velocity_func / normal_acceleration_func   :2422-2483
//
// Structure: RkPusher::push<FAST> is written once.  With FAST = true every branch that would enter the
// last-line-of-defence / bisection ladders returns false instead ("not decided"); the kernel then re-runs the
// push through push_rk_full_call (FAST = false), a non-inlined by-value function, exactly as the polynomial
// pusher does.  The ODE coefficients b, amat, Bvec, spamat and the step estimate dtau_ref are the polynomial
// pusher's b, A and physical_estimate_tau (same formulas in the reference, :104-116 vs pusher_tetra_poly.f90
// :1504-1517 and :170-186 vs :2779-2831), so that object is reused.
#pragma once
#include "gb_poly.cuh"

namespace gb {

#define GB_RK_KITER 48

// dz/dtau = b + A z of the RK module (rhs_pusher_tetra_rk4): (b + matmul(amat, z(1:3))) + Bvec*z(4) ; b4 + spamat*z4
template <class PP>
GB_HD void bm_vec_rk(double *o, const PP &P, const double *z)
{
#pragma unroll
  for (int i = 0; i < 3; i++) {
    double mv = (P.A.m[i][0] * z[0] + P.A.m[i][1] * z[1]) + P.A.m[i][2] * z[2];
    o[i] = P.b[i] + mv + P.A.c[i] * z[3];
  }
  o[3] = P.b[3] + P.A.s * z[3];
}

// ---- boole_pusher_ode45 (EXT = 2 kernels of the RK pusher): r8_fehl / r8_rkf45 (SRC/contrib/rkf45.f90:776-923, 925-1578) for the four
// equations dz/dtau = b + a z + Bvec v_par, and odeint_allroutines (SRC/odeint_rkf45.f90).  Only the paths reachable from
// odeint_allroutines are restated (first call with flag = 1, continuation with flag = 2 after a return of 6 or 7).
// x**0.2 is pow() of the platform's libm: the CUDA one on the device, which is not glibc's to the last bit -- the one
// place where this mode can leave the oracle's rounding (parity of this mode is asserted at 1e-10, not bit for bit).
// NEQ = 4: the orbit (rhs_pusher_tetra_rk45); NEQ = 5: orbit + integral of v_par^2 (rhs_par_adiab_ode45, :2779-2790)
template <int NEQ>
struct Rkf45T {
  double abserr_save, h, relerr_save, f1[NEQ], f2[NEQ], f3[NEQ], f4[NEQ], f5[NEQ];
  int flag_save, init, kflag, kop, nfe;
};
// the linear right-hand side of one tetrahedron, by value: b, A = (amat | Bvec | spamat)
struct OdeLin {
double b[4];
BlockMat A;
};
template <int NEQ>
GB_HD void ode_rhs(const OdeLin &L, const double *z, double *dz)
{
  bm_vec_rk(dz, L, z);
  if (NEQ == 5) dz[4] = z[3] * z[3];
}
template <int NEQ>
GB_HD void rkf45_fehl(const OdeLin &L, const double *y, double h, const double *yp, Rkf45T<NEQ> &q)
{
  double ch = h / 4.0, t1[NEQ];
#pragma unroll
  for (int i = 0; i < NEQ; i++) q.f5[i] = y[i] + ch * yp[i];
  ode_rhs<NEQ>(L, q.f5, q.f1);
  ch = 3.0 * h / 32.0;
#pragma unroll
  for (int i = 0; i < NEQ; i++) q.f5[i] = y[i] + ch * (yp[i] + 3.0 * q.f1[i]);
  ode_rhs<NEQ>(L, q.f5, q.f2);
  ch = h / 2197.0;
#pragma unroll
  for (int i = 0; i < NEQ; i++) q.f5[i] = y[i] + ch * (1932.0 * yp[i] + (7296.0 * q.f2[i] - 7200.0 * q.f1[i]));
  ode_rhs<NEQ>(L, q.f5, q.f3);
  ch = h / 4104.0;
#pragma unroll
  for (int i = 0; i < NEQ; i++)
    q.f5[i] = y[i] + ch * ((8341.0 * yp[i] - 845.0 * q.f3[i]) + (29440.0 * q.f2[i] - 32832.0 * q.f1[i]));
  ode_rhs<NEQ>(L, q.f5, q.f4);
  ch = h / 20520.0;
#pragma unroll
  for (int i = 0; i < NEQ; i++)
    t1[i] = y[i] + ch * ((-6080.0 * yp[i] + (9295.0 * q.f3[i] - 5643.0 * q.f4[i])) + (41040.0 * q.f1[i] - 28352.0 * q.f2[i]));
#pragma unroll
  for (int i = 0; i < NEQ; i++) q.f1[i] = t1[i];
  ode_rhs<NEQ>(L, q.f1, q.f5);
  ch = h / 7618050.0;
#pragma unroll
  for (int i = 0; i < NEQ; i++)   // the solution estimate goes to f1 (the caller passes f1 as s)
    q.f1[i] = y[i] + ch * ((902880.0 * yp[i] + (3855735.0 * q.f3[i] - 1371249.0 * q.f4[i])) +
                           (3953664.0 * q.f2[i] + 277020.0 * q.f5[i]));
}
template <int NEQ>
GB_HD int rkf45_run(const OdeLin &L, Rkf45T<NEQ> &q, double *y, double *yp, double &t, double tout, double &relerr, double abserr, int flag)
{
  const double remin = 1.0e-12, eps = DBL_EPSILON;
  const int maxnfe = 3000;
  if (relerr < 0.0 || abserr < 0.0) return 8;
  if (flag == 0 || 8 < flag || flag < -2) return 8;
  int mflag = flag < 0 ? -flag : flag;
  if (mflag != 1) {
    if (t == tout && q.kflag != 3) return 8;
    if (mflag == 2) {
      if (q.kflag == 3) { flag = q.flag_save; mflag = flag < 0 ? -flag : flag; }
      else if (q.init == 0) flag = q.flag_save;
      else if (q.kflag == 4) q.nfe = 0;
      else if (q.kflag == 5 && abserr == 0.0) return 8;
      else if (q.kflag == 6 && relerr <= q.relerr_save && abserr <= q.abserr_save) return 8;
    } else {
      return 8;
    }
  }
  q.flag_save = flag;
  q.kflag = 0;
  q.relerr_save = relerr;
  q.abserr_save = abserr;
  const double relerr_min = 2.0 * DBL_EPSILON + remin;
  if (relerr < relerr_min) {
    relerr = relerr_min;
    q.kflag = 3;
    return 3;
  }
  double dt = tout - t;
  if (mflag == 1) {
    q.init = 0;
    q.kop = 0;
    ode_rhs<NEQ>(L, y, yp);
    q.nfe = 1;
    if (t == tout) return 2;
  }
  if (q.init == 0) {
    q.init = 1;
    q.h = fabs(dt);
    double toln = 0.0;
#pragma unroll
    for (int k = 0; k < NEQ; k++) {
      const double tol = relerr * fabs(y[k]) + abserr;
      if (0.0 < tol) {
        toln = tol;
        const double ypk = fabs(yp[k]);
        const double h2 = q.h * q.h;
        if (tol < ypk * (q.h * (h2 * h2))) q.h = pow(tol / ypk, 0.2);
      }
    }
    if (toln <= 0.0) q.h = 0.0;
    q.h = fmax(q.h, 26.0 * eps * fmax(fabs(t), fabs(dt)));
    q.flag_save = flag < 0 ? -2 : 2;
  }
  q.h = copysign(q.h, dt);
  if (2.0 * fabs(dt) <= fabs(q.h)) q.kop = q.kop + 1;
  if (q.kop == 10000) {
    q.kop = 0;
    return 7;
  }
  if (fabs(dt) <= 26.0 * eps * fabs(t)) {
    t = tout;
#pragma unroll
    for (int i = 0; i < NEQ; i++) y[i] = y[i] + dt * yp[i];
    ode_rhs<NEQ>(L, y, yp);
    q.nfe = q.nfe + 1;
    return 2;
  }
  bool output = false;
  const double scale = 2.0 / relerr, ae = scale * abserr;
  for (;;) {
    bool hfaild = false;
    const double hmin = 26.0 * eps * fabs(t);
    dt = tout - t;
    if (!(2.0 * fabs(q.h) <= fabs(dt))) {
      if (fabs(dt) <= fabs(q.h)) { output = true; q.h = dt; }
      else q.h = 0.5 * dt;
    }
    double esttol;
    for (;;) {
      if (maxnfe < q.nfe) { q.kflag = 4; return 4; }
      rkf45_fehl<NEQ>(L, y, q.h, yp, q);
      q.nfe = q.nfe + 5;
      double eeoet = 0.0;
#pragma unroll
      for (int k = 0; k < NEQ; k++) {
        const double et = fabs(y[k]) + fabs(q.f1[k]) + ae;
        if (et <= 0.0) return 5;
        const double ee = fabs((-2090.0 * yp[k] + (21970.0 * q.f3[k] - 15048.0 * q.f4[k])) +
                               (22528.0 * q.f2[k] - 27360.0 * q.f5[k]));
        eeoet = fmax(eeoet, ee / et);
      }
      esttol = fabs(q.h) * eeoet * scale / 752400.0;
      if (esttol <= 1.0) break;
      hfaild = true;
      output = false;
      double sf;
      if (esttol < 59049.0) sf = 0.9 / pow(esttol, 0.2);
      else sf = 0.1;
      q.h = sf * q.h;
      if (fabs(q.h) < hmin) { q.kflag = 6; return 6; }
    }
    t = t + q.h;
#pragma unroll
    for (int i = 0; i < NEQ; i++) y[i] = q.f1[i];
    ode_rhs<NEQ>(L, y, yp);
    q.nfe = q.nfe + 1;
    double sf;
    if (0.0001889568 < esttol) sf = 0.9 / pow(esttol, 0.2);
    else sf = 5.0;
    if (hfaild) sf = fmin(sf, 1.0);
    q.h = copysign(fmax(sf * fabs(q.h), hmin), q.h);
    if (output) {
      t = tout;
      return 2;
    }
    if (flag <= 0) break;
  }
  return -2;
}
template <int NEQ>
struct VecN {
double v[NEQ];
};
typedef VecN<4> Vec4;
// odeint_allroutines(y, 4, 0, x2, eps, rhs) (SRC/odeint_rkf45.f90).  One shared, non-inlined instance: the RK pusher calls
// it from more than a dozen integration_step sites.
template <int NEQ>
GB_HD_NOINLINE VecN<NEQ> odeint_rkf45_n(OdeLin L, VecN<NEQ> y0, double x2, double eps_rel)
{
Rkf45T<NEQ> q;
q.abserr_save = q.h = q.relerr_save = 0.0;
q.flag_save = q.init = q.kflag = q.kop = q.nfe = 0;
double yp[NEQ], epsrel = eps_rel, epsabs = 1e-31, x1in = 0.0;
double *y = y0.v;
int flag = rkf45_run<NEQ>(L, q, y, yp, x1in, x2, epsrel, epsabs, 1);
if (flag == 6) {
  epsrel = 10 * epsrel;
  epsabs = 10 * epsabs;
  rkf45_run<NEQ>(L, q, y, yp, x1in, x2, epsrel, epsabs, 2);
} else if (flag == 7) {
  rkf45_run<NEQ>(L, q, y, yp, x1in, x2, epsrel, epsabs, 2);
}
return y0;
}
GB_HD Vec4 odeint_rkf45(const OdeLin &L, const Vec4 &y0, double x2, double eps_rel) { return odeint_rkf45_n<4>(L, y0, x2, eps_rel); }

// EXT = 2: hand-over via Cartesian skew coordinates when the mesh carries them (handover_processing_kind = 2)
template <int PHI, int EXT = 0>
struct RkPusher {
  PolyPusher<1, PHI, EXT> P;  // record, z_init, sign_rhs, dt_dtau_const, b, A (= amat | Bvec | spamat)
  double dist_min, dist_max, dtau_ref, dtau_max, dtau_quad, t_remain;
  int iface_init, sign_t_step, fallback;
  bool acc;   // boole_accuracy_ode45 of the routine that is running (EXT = 2 kernels; = boole_pusher_ode45 except in the Newton wrapper)

  GB_HD void init(const MeshDev *mp, double perpinv, int ind_tetr, const double *x, int iface, double vpar, double t_remain_in)
  {
    P.mp = mp;
    P.perpinv = perpinv;
    P.init(ind_tetr, x, iface, vpar, t_remain_in);
    P.template build_ode<true>();
    t_remain = t_remain_in;
    iface_init = iface;
    sign_t_step = signbit(t_remain_in) ? -1 : 1;
    const double dist1 = -P.r.dist_ref;
    dist_min = 1.e-10 * fabs(dist1);
    dist_max = 10.0 * fabs(dist1);
    dtau_ref = P.physical_estimate_tau();
    dtau_max = 10.0 * dtau_ref;
    dtau_quad = 1.5 * dtau_ref;
    fallback = 0;
    acc = (EXT == 2) && (mp->ode45 != 0);   // :261
  }

  GB_HD void distances(const double *z, double *d) const { P.normal_distances(z, d); }
  GB_HD double distance(const double *z, int iface) const { return P.normal_distance(z, iface); }
  // boole_newton_precalc (EXT = 2 kernels): sum((anorm_in_amat<k>_0 + perpinv*..._1 [+ perpinv^2*..._2])(:,iface) * v)
  GB_HD double p4_dot(int order, int iface, const double *v) const
  {
    const double *q = P.p4() + P4_AN_AMAT + (order == 1 ? 0 : 32) + 4 * (iface - 1);
    const double perpinv2 = P.perpinv * P.perpinv;   // analytic_coeff / initialize_const_motion_rk
    double sacc = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      double e = ldg(q + i) + P.perpinv * ldg(q + 16 + i);
      if (order == 2) e = e + perpinv2 * ldg(q + 32 + i);
      sacc = sacc + e * v[i];
    }
    return sacc;
  }
  GB_HD bool precalc() const { return EXT == 2 && P.mp->newton_precalc; }
  // normal_velocity_func (:2451-2467) / normal_velocity_analytic (:2487-2505)
  GB_HD double nvel(int iface, const double *dzdtau, const double *z) const
  {
    double n[3];
    P.face_normal(iface, n);
    if (precalc()) return p4_dot(1, iface, z) * (double)P.sign_rhs + dot3(n, P.b);
    return dot3(dzdtau, n);
  }
  // normal_acceleration_func (:2469-2483) / normal_acceleration_analytic (:2507-2527)
  GB_HD double nacc(int iface, const double *dzdtau, const double *z) const
  {
    if (precalc()) return p4_dot(2, iface, z) + p4_dot(1, iface, P.b) * (double)P.sign_rhs;
    double n[3], t[3];
    P.face_normal(iface, n);
#pragma unroll
    for (int j = 0; j < 3; j++) t[j] = (n[0] * P.A.m[0][j] + n[1] * P.A.m[1][j]) + n[2] * P.A.m[2][j];
    return dot3(t, dzdtau) + dot3(n, P.A.c) * dzdtau[3];
  }
  GB_HD static bool any_gt(const double *d, double lim) { return d[0] > lim || d[1] > lim || d[2] > lim || d[3] > lim; }
  GB_HD static int minloc4(const double *d)  // minloc(d,1): first minimum, 1-based
  {
    int k = 0;
    double cur = d[0];
    if (d[1] < cur) { k = 1; cur = d[1]; }
    if (d[2] < cur) { k = 2; cur = d[2]; }
    if (d[3] < cur) { k = 3; }
    return k + 1;
  }
  GB_HD static double sel4(const double *d, int k1) { return k1 == 1 ? d[0] : k1 == 2 ? d[1] : k1 == 3 ? d[2] : d[3]; }

  GB_HD void rk4_step(double *y, double h, double *dzdtau) const
  {
    const double hh = h * 0.5, h6 = h / 6.0;
    double dydx[4], yt[4], dyt[4], dym[4];
    bm_vec_rk(dydx, P, y);
#pragma unroll
    for (int i = 0; i < 4; i++) yt[i] = y[i] + hh * dydx[i];
    bm_vec_rk(dyt, P, yt);
#pragma unroll
    for (int i = 0; i < 4; i++) yt[i] = y[i] + hh * dyt[i];
    bm_vec_rk(dym, P, yt);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      yt[i] = y[i] + h * dym[i];
      dym[i] = dyt[i] + dym[i];
    }
    bm_vec_rk(dyt, P, yt);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      y[i] = y[i] + h6 * (dydx[i] + dyt[i] + 2.0 * dym[i]);
      dzdtau[i] = dyt[i];
    }
  }

  // integration_step (:2549-2581)
  GB_HD void integration_step(double *z, double dtau, double *dzdtau) const
  {
    if (EXT == 2 && acc) {
      OdeLin L;
      Vec4 y;
#pragma unroll
      for (int i = 0; i < 4; i++) { L.b[i] = P.b[i]; y.v[i] = z[i]; }
      L.A = P.A;
      y = odeint_rkf45(L, y, dtau, P.mp->rel_err_ode45);
#pragma unroll
      for (int i = 0; i < 4; i++) z[i] = y.v[i];
      rk4_step(z, 0.0, dzdtau);
    } else {
      rk4_step(z, dtau, dzdtau);
    }
  }

  // :636-809.  allowed: bit f set = face f+1 allowed.  The per-face sign tree is the polynomial pusher's
  // closed-form quadratic with two differences: the start face uses the reduced (linear) form and the
  // "c == 0" branch is taken for |c| <= dist_min.
  GB_HD bool quad_analytic_approx(const double *z, unsigned allowed, int &iface_inout, double &dtau) const
  {
    double cc[4];
    distances(z, cc);
    const int iface = iface_inout;
    const double fac = P.b[3] + P.A.s * z[3];
    double best = 0.0;
    int ibest = 0;
#pragma unroll
    for (int f = 0; f < 4; f++) {
      if (!(allowed & (1u << f))) continue;
      // acoef_pre = matmul(curlA, anorm) re-formed (tetra_physics_mod.f90:857), times sign_rhs
      double apre = dot3(P.r.curlA, P.r.an[f]) * (double)P.sign_rhs;
      // strong electric field (:672): + cm_over_e * matmul(curlvE, anorm) * sign_rhs
      if (PHI == 2) apre = apre + P.mp->cm_over_e * dot3(P.r.curlvE, P.r.an[f]) * (double)P.sign_rhs;
      double b = z[3] * apre + dot3(P.b, P.r.an[f]);
      double a = apre * fac;
      double c = cc[f];
      if (precalc()) {   // analytic_coeff(2, z, coef_mat) (:579-632): all three coefficients from the poly4 record
        c = dot3(z, P.r.an[f]);
        if (f == 0) c = c + P.r.dist_ref;
        b = p4_dot(1, f + 1, z) * (double)P.sign_rhs + dot3(P.r.an[f], P.b);
        a = p4_dot(2, f + 1, z) + p4_dot(1, f + 1, P.b) * (double)P.sign_rhs;
      }
      double num = 1.0, den = 1.0;
      bool has;
      if (iface == f + 1) {
        has = ((a > 0.0) && (b < 0.0)) || (!(a > 0.0) && (a < 0.0) && (b > 0.0));
        if (has) { num = -2.0 * b; den = a; }
      } else if (fabs(c) > dist_min) {
        has = quadratic_solver1_numden(a, b, c, num, den);
        // quadratic_solver1's "c == 0" leaf (neither c > 0 nor c < 0) only triggers for NaN here: the reference
        // stops ('Should not happen'); treat as no root
        if (!(c > 0.0) && !(c < 0.0)) has = false;
        if (!has) { num = 1.0; den = 1.0; }
      } else {
        has = ((a > 0.0) && (b < 0.0)) || ((a < 0.0) && (b > 0.0));
        if (has) { num = -2.0 * b; den = a; }
      }
      const double d = num / den;
      if (has && (d < dtau_max) && (d > 0.0) && (ibest == 0 || d < best)) {
        best = d;
        ibest = f + 1;
      }
    }
    if (ibest == 0) return false;
    iface_inout = ibest;
    dtau = best;
    return true;
  }

  // :1000-1188
  GB_HD bool newton_wrapped(double *z, double &tau, int iface, double *dzdtau, bool start_quadratic)
  {
    double z_start[4], dz_start[4], z_save[4], dz_save[4], nd[4];
    const double tau_start = tau;
    double dtau = 0.0, tau_save = 0.0, dist_new = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++) { z_start[i] = z[i]; dz_start[i] = dzdtau[i]; }
    double dist = distance(z, iface);
    int k = 0;
    while (fabs(dist) > dist_min) {
      k++;
#pragma unroll
      for (int i = 0; i < 4; i++) { z_save[i] = z[i]; dz_save[i] = dzdtau[i]; }
      const double nv = nvel(iface, dzdtau, z);
      if (nv != 0.0) dtau = -dist / nv;
      else return false;
      tau_save = tau;
      distances(z, nd);
      if (any_gt(nd, dist_max)) {
        dtau = tau + dtau;
        tau = 0.0;
#pragma unroll
        for (int i = 0; i < 4; i++) z[i] = P.z_init[i];
      }
      if (fabs(dtau) > dtau_max) {
        start_quadratic = true;
      } else {
        integration_step(z, dtau, dzdtau);
        dist_new = distance(z, iface);
      }
      if ((fabs(dist_new) >= fabs(dist)) || start_quadratic) {
        start_quadratic = false;
#pragma unroll
        for (int i = 0; i < 4; i++) { z[i] = z_save[i]; dzdtau[i] = dz_save[i]; }
        tau = tau_save;
        const double na = 0.5 * nacc(iface, dzdtau, z);
        const double discr = nv * nv - 4.0 * na * dist;
        if (discr > 0.0) {
          if (na < 0.0) dtau = (-nv - sqrt(discr)) / (2.0 * na);
          else if (na > 0.0) dtau = (-nv + sqrt(discr)) / (2.0 * na);
          else dtau = -dist / nv;
          distances(z, nd);
          if (any_gt(nd, dist_max)) {
            dtau = tau + dtau;
            tau = 0.0;
#pragma unroll
            for (int i = 0; i < 4; i++) z[i] = P.z_init[i];
          }
          if (fabs(dtau) > dtau_max) {
#pragma unroll
            for (int i = 0; i < 4; i++) { z[i] = z_start[i]; dzdtau[i] = dz_start[i]; }
            tau = tau_start;
            return false;
          }
          integration_step(z, dtau, dzdtau);
          tau = tau + dtau;
          dist = distance(z, iface);
        } else {
          return false;
        }
      } else {
        tau = tau + dtau;
        dist = dist_new;
      }
      if (k > GB_RK_KITER) return false;
    }
    if (tau <= 0.0) {
#pragma unroll
      for (int i = 0; i < 4; i++) { z[i] = z_start[i]; dzdtau[i] = dz_start[i]; }
      tau = tau_start;
      return false;
    }
    return true;
  }
  // :914-996 (RK4 accuracy): state restored when Newton did not converge
  GB_HD bool newton(double *z, double &tau, int iface, double *dzdtau, bool start_quadratic)
  {
    double z_save[4], dz_save[4];
    const double tau_save = tau;
#pragma unroll
    for (int i = 0; i < 4; i++) { z_save[i] = z[i]; dz_save[i] = dzdtau[i]; }
    const bool acc_in = acc;   // boole_accuracy_ode45_in
    acc = false;               // Newton with the RK4 method first (:938-941)
    bool ok = newton_wrapped(z, tau, iface, dzdtau, start_quadratic);
    acc = acc_in;
    if (!ok) {
#pragma unroll
      for (int i = 0; i < 4; i++) { z[i] = z_save[i]; dzdtau[i] = dz_save[i]; }
      tau = tau_save;
      return false;
    }
    if (EXT == 2 && acc_in) {   // repeat the RK4-Newton step with ODE45, then converge with the ODE45-Newton (:950-979)
#pragma unroll
      for (int i = 0; i < 4; i++) z[i] = z_save[i];
      const double dtau = tau - tau_save;
      integration_step(z, dtau, dzdtau);
      ok = newton_wrapped(z, tau, iface, dzdtau, false);
    }
    return ok;
  }

  // :1583-1623
  GB_HD void bisection_search_start(double tau_in, int n_steps, double *z_start, double &dtau, double &tau_out)
  {
    double zr[4], dz[4], nd[4], tau_run = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++) { zr[i] = P.z_init[i]; z_start[i] = zr[i]; }
    dtau = tau_in / (double)n_steps;
    int last_inside = 0;
    bool take_next = false;
    distances(zr, nd);
    tau_out = 0.0;
    if (nd[0] > 0.0 && nd[1] > 0.0 && nd[2] > 0.0 && nd[3] > 0.0) { last_inside = 1; take_next = true; }
    for (int i = 2; i <= n_steps; i++) {
      integration_step(zr, dtau, dz);
      tau_run = tau_run + dtau;
      distances(zr, nd);
      if (take_next) {
#pragma unroll
        for (int q = 0; q < 4; q++) z_start[q] = zr[q];
        tau_out = tau_run;
        take_next = false;
      }
      if (nd[0] > 0.0 && nd[1] > 0.0 && nd[2] > 0.0 && nd[3] > 0.0) { last_inside = i; take_next = true; }
    }
    if (last_inside == n_steps) {
#pragma unroll
      for (int q = 0; q < 4; q++) z_start[q] = zr[q];
      tau_out = tau_run;
    }
  }

  // :1409-1581
  GB_HD bool bisection(double *z, double &tau_inout, double dtau_in, int &iface, double *dzdtau)
  {
    double tau = tau_inout, dtau = dtau_in, z_save[4], nd[4];
#pragma unroll
    for (int i = 0; i < 4; i++) z_save[i] = z[i];
    bool converged = false;
    for (int l = 1; l <= 2 && !converged; l++) {
      int k = 0;
      bool stop_all = false;
      while (!converged) {
        distances(z, nd);
        const double mn = sel4(nd, minloc4(nd));
        if (mn < -dist_min) {
          dtau = -fabs(dtau / 2.0);
          if (any_gt(nd, dist_max)) {
            const double dtau_save = dtau;
            dtau = tau + dtau;
            tau = 0.0;
#pragma unroll
            for (int i = 0; i < 4; i++) z[i] = P.z_init[i];
            integration_step(z, dtau, dzdtau);
            tau = tau + dtau;
            dtau = dtau_save;
          } else {
            integration_step(z, dtau, dzdtau);
            tau = tau + dtau;
          }
        } else if (mn > dist_min) {
          dtau = +fabs(dtau / 2.0);
          integration_step(z, dtau, dzdtau);
          tau = tau + dtau;
        }
        distances(z, nd);
        const int im = minloc4(nd);
        if (fabs(sel4(nd, im)) < dist_min) {
          if (nvel(im, dzdtau, z) > 0.0) {
            dtau = +fabs(dtau / 2.0);
            integration_step(z, dtau, dzdtau);
            tau = tau + dtau;
          } else {
            int j = 0;
#pragma unroll
            for (int i = 0; i < 4; i++)
              if (nd[i] < 0.0) j++;
            if (j <= 1) {
              iface = im;
              converged = true;
            } else {
              dtau = -fabs(dtau / 2.0);
              integration_step(z, dtau, dzdtau);
              tau = tau + dtau;
            }
          }
        }
        k++;
        if (k > GB_RK_KITER) {
          if (l == 1) {
            int nfc = 0;
#pragma unroll
            for (int i = 0; i < 4; i++)
              if (nd[i] < 0.0 && fabs(nd[i]) < dist_min) nfc++;
            if (nfc > 1) {
              dist_min = 2.0 * dist_min;
              tau = tau_inout;
              dtau = dtau_in;
#pragma unroll
              for (int i = 0; i < 4; i++) z[i] = z_save[i];
            } else {
              bisection_search_start(tau_inout, 1000, z, dtau, tau);
            }
          } else {
            stop_all = true;
          }
          break;
        }
      }
      if (stop_all) break;
    }
    tau_inout = tau;
    return converged;
  }

  // :1696-2071 ; z, tau, iface, dzdtau are outputs
  GB_HD bool last_line_defense(double *z, double &tau, int &iface, double *dzdtau)
  {
    fallback |= 2;
    bool turned_tangential = false, converged = false;
    double dtau = 0.0, nd[4], z_save[4];
    int iface_new = iface_init, iface_init_outside = 0, k;
    tau = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++) z[i] = P.z_init[i];
    if (iface_init != 0)
      if (distance(z, iface_init) < 0.0) iface_init_outside = iface_init;
    if (quad_analytic_approx(z, 0xFu, iface_new, dtau)) {
      integration_step(z, dtau, dzdtau);
      tau = tau + dtau;
    } else {
      dtau = dtau_ref;
      integration_step(z, dtau, dzdtau);
      tau = tau + dtau;
      distances(z, nd);
      iface_new = minloc4(nd);
    }
    k = 0;
    bool distance_bisection = false;
    for (;;) {
      k++;
      distances(z, nd);
      if (any_gt(nd, dist_max)) {
        distance_bisection = true;
        dtau = tau - 0.5 * fabs(dtau);
        tau = 0.0;
#pragma unroll
        for (int i = 0; i < 4; i++) z[i] = P.z_init[i];
        rk4_step(z, dtau, dzdtau);   // integration_step(..., .false.): "Set accuracy to FALSE, always!" (:1798)
        tau = tau + dtau;
      } else {
        if (EXT == 2 && distance_bisection && acc) {   // :1802-1808
#pragma unroll
          for (int i = 0; i < 4; i++) z[i] = P.z_init[i];
          dtau = tau;
          integration_step(z, dtau, dzdtau);
          distance_bisection = false;
        } else {
          break;
        }
      }
      if (k > GB_RK_KITER) return false;
    }
    if (iface_init_outside != 0) {
      if (distance(z, iface_init_outside) < 0.0) {
        if (nvel(iface_init_outside, dzdtau, z) < 0.0) {
          for (int i = 1; i <= 3; i++) {
            const int j = ((iface_init_outside + i - 1) & 3) + 1;
            if (distance(z, j) < 0.0) turned_tangential = true;
          }
          iface_new = iface_init_outside;
          if (fabs(distance(z, iface_new)) < dist_min) converged = true;
        }
      }
    }
    if (!converged) {
      if (turned_tangential) {
        if (!bisection(z, tau, dtau, iface_new, dzdtau)) return false;
      } else {
        bool dtau_decreased = false;
        k = 0;
        for (;;) {
          k++;
          distances(z, nd);
          int n_out = 0, first_out = 0;
#pragma unroll
          for (int i = 3; i >= 0; i--)
            if (nd[i] < 0.0) { n_out++; first_out = i + 1; }
          if (n_out == 0) {
            dtau = dtau_decreased ? 0.5 * fabs(dtau) : 2.0 * fabs(dtau);
          } else if (n_out == 1) {
            iface_new = first_out;
            if (nvel(iface_new, dzdtau, z) >= 0.0) {
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
#pragma unroll
            for (int i = 0; i < 4; i++)
              if (nd[i] < 0.0 && fabs(nd[i]) < dist_min) l++;
            if (l == n_out) {
              int j = 0;
              for (int i = 0; i < 4; i++) {
                const double di = sel4(nd, i + 1);
                if (!(di < 0.0)) continue;
                if (fabs(di) >= dist_min) continue;
                if (nvel(i + 1, dzdtau, z) > 0.0) j++;
              }
              if (j > 0) {
                dtau = 2.0 * fabs(dtau);
              } else {
                int best = 0;
                double bv = 0.0;
#pragma unroll
                for (int i = 0; i < 4; i++)
                  if (nd[i] < 0.0 && (best == 0 || fabs(nd[i]) < bv)) { best = i + 1; bv = fabs(nd[i]); }
                iface_new = best;
                break;
              }
            } else {
              dtau = -0.5 * fabs(dtau);
              dtau_decreased = true;
            }
          }
          if (any_gt(nd, dist_max)) {
            dtau = tau + dtau;
            tau = 0.0;
#pragma unroll
            for (int i = 0; i < 4; i++) z[i] = P.z_init[i];
          }
          integration_step(z, dtau, dzdtau);
          tau = tau + dtau;
          if (k > GB_RK_KITER) return false;
        }
#pragma unroll
        for (int i = 0; i < 4; i++) z_save[i] = z[i];
        const double dtau_save = dtau, tau_save = tau;
        bool newton_ok = newton(z, tau, iface_new, dzdtau, false);
        for (int i = 1; i <= 3; i++) {
          const int j = ((iface_new + i - 1) & 3) + 1;
          if (distance(z, j) < 0.0) newton_ok = false;
        }
        if ((!newton_ok) || (nvel(iface_new, dzdtau, z) >= 0.0)) {
#pragma unroll
          for (int i = 0; i < 4; i++) z[i] = z_save[i];
          tau = tau_save;
          dtau = dtau_save;
          if (!bisection(z, tau, dtau, iface_new, dzdtau)) return false;
        }
      }
    }
    iface = iface_new;
    return true;
  }

  GB_HD void pass_through(const double *z, double tau, int iface_new, bool finished, PushOut &o) const
  {
#pragma unroll
    for (int i = 0; i < 3; i++) {
      o.x[i] = z[i] + P.r.x1s(i);
      o.z_save[i] = z[i];
    }
    o.z_save_set = 1;
    o.vpar = z[3];
    o.t_pass = tau * P.dt_dtau_const;
    o.finished = finished ? 1 : 0;
    P.handover(iface_new, o.x, o.ind_tetr, o.iface);
  }

  // 0 = decided and stored in o, 1 = particle removed, 2 = (FAST only) needs the complete path
  template <bool FAST>
  GB_HD int final_processing(double *z, double tau, int iface_new, PushOut &o)
  {
    double dzdtau[4], nd[4], nd_save[4], z_save[4];
    const double t_pass0 = tau * P.dt_dtau_const;
    // x and t_pass are assigned before anything can fail (:2100-2105); a removal further down keeps them
#pragma unroll
    for (int i = 0; i < 3; i++) o.x[i] = z[i] + P.r.x1s(i);
    o.t_pass = t_pass0;
    if (!(fabs(t_remain) < fabs(t_pass0))) {
      pass_through(z, tau, iface_new, false, o);
      return 0;
    }
    // the orbit stops inside the cell (:2120-2400)
#pragma unroll
    for (int i = 0; i < 4; i++) { z[i] = P.z_init[i]; z_save[i] = z[i]; }
    tau = 0.0;
    const double dtau = t_remain / P.dt_dtau_const;
    distances(z, nd_save);
    integration_step(z, dtau, dzdtau);
    if (any_gt(nd_save, dist_max)) {
      if (FAST) return 2;
#pragma unroll
      for (int i = 0; i < 4; i++) z[i] = z_save[i];
      if (!last_line_defense(z, tau, iface_new, dzdtau)) { /* reference ignores the flag here */ }
      for (int j = 1; j <= 3; j++) {
        const int k = ((iface_new + j - 1) & 3) + 1;
        if (distance(z, k) < 0.0) return 1;
      }
      if (nvel(iface_new, dzdtau, z) > 0.0) return 1;
#pragma unroll
      for (int i = 0; i < 3; i++) o.x[i] = z[i] + P.r.x1s(i);
      o.t_pass = tau * P.dt_dtau_const;
      if (fabs(tau * P.dt_dtau_const) <= fabs(t_remain)) {
        pass_through(z, tau, iface_new, false, o);
        return 0;
      }
      return 1;
    }
    tau = tau + dtau;
#pragma unroll
    for (int i = 0; i < 3; i++) o.x[i] = z[i] + P.r.x1s(i);
    o.t_pass = tau * P.dt_dtau_const;
    distances(z, nd);
    int n_out = 0, iface_outside = 0;
    bool any_conv = false;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (nd[i] < 0.0) { n_out++; iface_outside = i + 1; }
      if (fabs(nd[i]) < dist_min) any_conv = true;
    }
    if (any_conv) {
      if (n_out == 1) iface_new = iface_outside;
      else {
#pragma unroll
        for (int i = 0; i < 4; i++)
          if (fabs(nd[i]) < dist_min) iface_new = i + 1;
      }
      if (nvel(iface_new, dzdtau, z) < 0.0) {
        pass_through(z, tau, iface_new, true, o);
      } else {
        // converged on a face at t_remain but flying inwards: handed to the same tetrahedron again
#pragma unroll
        for (int i = 0; i < 3; i++) { o.x[i] = z[i] + P.r.x1s(i); o.z_save[i] = z[i]; }
        o.z_save_set = 1;
        o.vpar = z[3];
        o.t_pass = tau * P.dt_dtau_const;
        o.finished = 1;
        o.ind_tetr = P.ind_tetr;
        o.iface = iface_new;
      }
      return 0;
    }
    if (n_out != 0) {
      if (FAST) return 2;
      fallback |= 8;
      if (n_out == 1) {
        iface_new = iface_outside;
        const double tau_save = tau;
#pragma unroll
        for (int i = 0; i < 4; i++) z_save[i] = z[i];
        bool ok = newton(z, tau, iface_new, dzdtau, true);
        if (!ok) {
#pragma unroll
          for (int i = 0; i < 4; i++) z[i] = z_save[i];
          tau = tau_save;
          rk4_step(z, 0.0, dzdtau);
          ok = newton(z, tau, iface_new, dzdtau, false);
          if (!ok) {
#pragma unroll
            for (int i = 0; i < 4; i++) z[i] = z_save[i];
            tau = tau_save;
            if (!bisection(z, tau, tau, iface_new, dzdtau)) return 1;
          }
        }
        if (nvel(iface_new, dzdtau, z) > 0.0) {
#pragma unroll
          for (int i = 0; i < 4; i++) z[i] = z_save[i];
          tau = tau_save;
          if (!bisection(z, tau, tau, iface_new, dzdtau)) return 1;
        }
      } else {
        if (!bisection(z, tau, tau, iface_new, dzdtau)) return 1;
      }
      pass_through(z, tau, iface_new, false, o);
      return 0;
    }
    // orbit time is finished inside the tetrahedron
#pragma unroll
    for (int i = 0; i < 3; i++) { o.x[i] = z[i] + P.r.x1s(i); o.z_save[i] = z[i]; }
    o.z_save_set = 1;
    o.vpar = z[3];
    o.t_pass = tau * P.dt_dtau_const;
    o.finished = 1;
    o.ind_tetr = P.ind_tetr;
    o.iface = 0;
    return 0;
  }

  // ==== EXT = 2: orbit events for the RK pusher ==========================================================================
  // module par_adiab_inv_rk_mod (:2589-2798): v_par^2 is integrated along the orbit as a fifth equation of the RKF45
  // integration (relative error rel_err_ode45) from z_init over tau = t_pass / dt_dtau_const; at a bounce (v_par from < 0
  // to > 0) the turning point is approached by steps of alternating sign and halving length until |v_par| <= 10 cm/s.
  // Followed by the toroidal mappings of gorilla_plot_orbit_integration (SRC/gorilla_plot_mod.f90:601-636).
  GB_HD void par_adiab_tau(const OdeLin &L, double dtau, double *z, double &J_tau) const
  {
    VecN<5> y;
#pragma unroll
    for (int i = 0; i < 4; i++) y.v[i] = z[i];
    y.v[4] = 0.0;
    y = odeint_rkf45_n<5>(L, y, dtau, P.mp->rel_err_ode45);
#pragma unroll
    for (int i = 0; i < 4; i++) z[i] = y.v[i];
    J_tau = y.v[4];
  }
  GB_HD void events_after_push(double vpar_in, const PushOut &o, int iper_phi, EvState &es)
  {
    es.n = 0;
    if (es.flags & 6) {
      OdeLin L;
#pragma unroll
      for (int i = 0; i < 4; i++) L.b[i] = P.b[i];
      L.A = P.A;
      const double tau = o.t_pass / P.dt_dtau_const;
      double z[4], J_tau;
#pragma unroll
      for (int i = 0; i < 4; i++) z[i] = P.z_init[i];
      if ((o.vpar > 0.0) && (vpar_in < 0.0)) {
        // calc_par_adiab_until_root (:2703-2760)
        VecN<5> y;
#pragma unroll
        for (int i = 0; i < 4; i++) y.v[i] = z[i];
        y.v[4] = 0.0;
        double dtau = tau, tau_part1 = 0.0;
        int it = 0;
        while (fabs(y.v[3]) > 1.e1) {
          it++;
          const double vpar_save = y.v[3];
          y = odeint_rkf45_n<5>(L, y, dtau, P.mp->rel_err_ode45);
          tau_part1 = tau_part1 + dtau;
          const bool same_side = (vpar_save > 0.0) == (y.v[3] > 0.0);
          if (same_side) dtau = (dtau > 0.0) ? fabs(dtau / 2) : -fabs(dtau / 2);
          else dtau = (dtau > 0.0) ? -fabs(dtau / 2) : fabs(dtau / 2);
          if (it > 100) break;
        }
#pragma unroll
        for (int i = 0; i < 4; i++) z[i] = y.v[i];
        es.J = es.J + y.v[4] * P.dt_dtau_const;
        if (es.cnt_v > 1 && (es.cnt_v / es.nskip_v * es.nskip_v == es.cnt_v)) {
          EvRec &e = es.e[es.n++];
          e.kind = 2;
          e.counter = es.cnt_v;
#pragma unroll
          for (int i = 0; i < 3; i++) e.x[i] = z[i] + P.r.x1s(i);
          e.v[0] = es.J;
          e.v[1] = P.energy_tot(z);
        }
        es.cnt_v = es.cnt_v + 1;
        es.J = 0.0;
        par_adiab_tau(L, tau - tau_part1, z, J_tau);
        es.J = es.J + J_tau * P.dt_dtau_const;
      } else {
        par_adiab_tau(L, tau, z, J_tau);
        es.J = es.J + J_tau * P.dt_dtau_const;
      }
    }
    if (iper_phi != 0) {
      es.cnt_p = es.cnt_p + iper_phi;
      if ((es.flags & 1) && (es.cnt_p / es.nskip_p * es.nskip_p == es.cnt_p)) {
        const double zv[4] = {o.z_save[0], o.z_save[1], o.z_save[2], o.vpar};
        EvRec &e = es.e[es.n++];
        e.kind = 1;
        e.counter = es.cnt_p;
#pragma unroll
        for (int i = 0; i < 3; i++) e.x[i] = o.x[i];
        e.v[0] = P.p_phi(o.vpar, o.z_save);
        e.v[1] = P.energy_tot(zv);
      }
    }
  }

  // pusher_tetra_rk (:197-575).  Returns false only with FAST = true ("take the complete path").
  template <bool FAST>
  GB_HD bool push(PushOut &o)
  {
    unsigned allowed = 0xFu;
    double z[4], dzdtau[4], nd[4], tau = 0.0, dtau = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++) z[i] = P.z_init[i];
    int iface_new = iface_init;
    o.finished = 0;
    o.z_save_set = 0;
    o.t_pass = 0.0;
    bool removed = false, converged = false;
    const bool have_guess = quad_analytic_approx(z, allowed, iface_new, dtau);
    // GATHER kernels: the record behind the quadratic guess of the exit face -> shared memory (every lane calls it: the
    // cooperative form is a warp operation)
    if (FAST) P.r.prefetch_next(*P.mp, have_guess ? P.r.nb(iface_new - 1) : 0);
    if (have_guess) {
      if (FAST && P.mp->prefetch) prefetch_record<PHI>(*P.mp, P.r.nb(iface_new - 1), P.r.gmode != 0);   // quadratic guess of the exit face
      integration_step(z, dtau, dzdtau);
      tau = tau + dtau;
    } else {
      if (FAST) return false;
      dtau = dtau_ref;
      rk4_step(z, dtau, dzdtau);   // a plain rk4_step in the reference (:315)
      tau = tau + dtau;
      distances(z, nd);
      double ad[4] = {fabs(nd[0]), fabs(nd[1]), fabs(nd[2]), fabs(nd[3])};
      iface_new = minloc4(ad);
    }
    distances(z, nd);
    if (any_gt(nd, dist_max)) {
      if (FAST) return false;
      if (!last_line_defense(z, tau, iface_new, dzdtau)) removed = true;
    }
    if (!removed) {
      for (int it = 1; it <= 5; it++) {
        converged = true;
        bool llod = false, requad = false, reset_first = false;
        if (!newton(z, tau, iface_new, dzdtau, false)) {
          fallback |= 1;
          allowed &= ~(1u << (iface_new - 1));
          if (allowed == 0) llod = true;
          else requad = true;
        } else {
          bool cycled = false;
          for (int j = 1; j <= 3 && !cycled; j++) {
            const int k = ((iface_new + j - 1) & 3) + 1;
            if (distance(z, k) < 0.0) {
              fallback |= 4;
              allowed &= ~(1u << (iface_new - 1));
              if (allowed == 0) llod = true;
              else if (allowed & (1u << (k - 1))) iface_new = k;
              else llod = true;
              converged = false;
              cycled = true;
            }
          }
          if (!cycled) {
            if (nvel(iface_new, dzdtau, z) > 0.0) {
              fallback |= 8;
              allowed &= ~(1u << (iface_new - 1));
              if (allowed == 0) llod = true;
              else requad = true;
            } else if (tau <= 0.0) {
              allowed &= ~(1u << (iface_new - 1));
#pragma unroll
              for (int i = 0; i < 4; i++) z[i] = P.z_init[i];
              tau = 0.0;
              if (allowed == 0) llod = true;
              else { requad = true; reset_first = true; }
            } else {
              break;  // converged
            }
          } else if (!llod) {
            continue;
          }
        }
        if (llod) {
          if (FAST) return false;  // only the last-line-of-defence / bisection ladder is left to the complete path
          last_line_defense(z, tau, iface_new, dzdtau);  // its flag is not examined in the loop (:358-361)
          converged = false;
          continue;
        }
        if (requad) {
          if (!reset_first && tau > dtau_quad) {
#pragma unroll
            for (int i = 0; i < 4; i++) z[i] = P.z_init[i];
            tau = 0.0;
          }
          if (!quad_analytic_approx(z, allowed, iface_new, dtau)) {
            if (FAST) return false;
            last_line_defense(z, tau, iface_new, dzdtau);
            converged = false;
            continue;
          }
          if (!reset_first) {
            distances(z, nd);
            if (any_gt(nd, dist_max)) {
              dtau = tau + dtau;
              tau = 0.0;
#pragma unroll
              for (int i = 0; i < 4; i++) z[i] = P.z_init[i];
            }
          }
          integration_step(z, dtau, dzdtau);
          tau = tau + dtau;
          converged = false;
          continue;
        }
      }
      if (!converged) {
        if (FAST) return false;
        removed = true;
      }
    }
    if (!removed) {
      const int rc = final_processing<FAST>(z, tau, iface_new, o);
      if (rc == 2) return false;
      if (rc == 1) removed = true;
    }
    if (!removed) {
      if ((fabs(o.t_pass) >= fabs(t_remain)) && !o.finished) removed = true;
      else if ((o.t_pass * (double)sign_t_step) <= 0.0) removed = true;
    }
    if (removed) {
      // x and vpar keep their input values unless final_processing already wrote them (reference: intent(out))
      o.ind_tetr = -1;
      o.iface = -1;
      o.finished = 0;
      o.z_save_set = 0;
    }
    o.fallback = fallback;
    // EXT = 2: the toroidal period the hand-over crossed travels in bits 8/9 (read by rk_events_call only)
    if (EXT == 2) o.fallback |= (P.iper_phi == 1) ? 256 : (P.iper_phi == -1) ? 512 : 0;
    return true;
  }
};

template <int PHI, int EXT = 0>
GB_HD_NOINLINE PushOut push_rk_full_call(const MeshDev *mp, double perpinv, int ind_tetr, int iface, double x0, double x1,
                                         double x2, double vpar, double t_remain)
{
  RkPusher<PHI, EXT> R;
  double stash[6];
  R.P.r.set_stash(stash, 1);
  PushOut o;
  const double x[3] = {x0, x1, x2};
  o.x[0] = x0; o.x[1] = x1; o.x[2] = x2; o.vpar = vpar;
  o.z_save[0] = o.z_save[1] = o.z_save[2] = 0.0;
  R.init(mp, perpinv, ind_tetr, x, iface, vpar, t_remain);
  R.template push<false>(o);
  return o;
}

// events of one RK push, out of line: the pusher state is set up again from the push's inputs (init is deterministic), so
// the orbit loop carries nothing extra for a feature that is off in production runs
template <int PHI>
GB_HD_NOINLINE EvState rk_events_call(const MeshDev *mp, double perpinv, int ind_tetr, int iface, double x0, double x1, double x2,
                                      double vpar_in, double t_remain, PushOut o, EvState es)
{
  RkPusher<PHI, 2> R;
  double stash[6];
  R.P.r.set_stash(stash, 1);
  const double x[3] = {x0, x1, x2};
  R.init(mp, perpinv, ind_tetr, x, iface, vpar_in, t_remain);
  const int iper_phi = (o.fallback & 256) ? 1 : (o.fallback & 512) ? -1 : 0;
  R.events_after_push(vpar_in, o, iper_phi, es);
  return es;
}

} // namespace gb
