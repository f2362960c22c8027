// mesh_efit_flux.cpp -- grid_kind = 2: field-aligned grid of an axisymmetric (EFIT) equilibrium in symmetry flux
// coordinates (s, theta, phi).
//
// Reference: field_line_integration_for_SYNCH   SRC/field_line_integration_for_SYNCH.f90:19-397
//            preload_for_SYNCH / load_magdata_in_symfluxcoord / magdata_in_symfluxcoord_ext
//                                                 SRC/preload_for_SYNCH.f90, SRC/magdata_in_symfluxcoord.f90
//            create_points_2d (inp_label 1)       SRC/points_2d.f90:18-167 ; set_grid_size SRC/tetra_grid_settings_mod.f90:117-135
//            make_grid_aligned                    SRC/tetra_grid_mod.f90:215-276
//            vector_potential_sthetaphi           SRC/tetra_physics_mod.f90:1113-1163 ; metric_determinant :1256-1287
//
// The flux coordinates are CONSTRUCTED, not read: field lines are followed around the torus to find the magnetic
// axis, the last closed surface and the X-point; on nlabel start points along the axis -> X-point ray one poloidal turn
// gives the safety factor, the enclosed toroidal flux and the poloidal flux of each surface, and ntheta equal steps in
// toroidal angle along the line give R, Z, |B| at equidistant symmetry-flux poloidal angle.  Interpolation is a periodic
// cubic spline in theta and 4-point Lagrange in the surface label, as in the reference.
// Differences to the reference, all at the level of the integrator tolerance (1e-9): the ODE integrator is an adaptive
// Dormand-Prince 5(4) pair instead of RKF45, the last closed surface is found by bisection instead of a linear scan of
// nsurfmax start points, and nothing is written to / re-read from files (box_size_axis.dat, flux_functions.dat,
// twodim_functions.dat round the numbers through list-directed formatting in the reference).
// theta_geom_flux = 1: poloidal grid equidistant in the flux angle; 2: equidistant in the geometrical angle around the
// magnetic axis (theta_geom2theta_flux, SRC/points_2d.f90:254-373), see geom_to_flux_angles below.
#include "mesh_efit.hpp"
#include <algorithm>
#include <cmath>

namespace gbhost {

void make_field_aligned_topology(Mesh &m, int n1, int n2, int n3);  // mesh_vmec.cpp (circular_mesh.f90 calc_mesh)

namespace {

const double PI_T = 3.14159265358979;  // truncated pi of the reference's field-line / mesh modules

// ---- adaptive Dormand-Prince 5(4) from x0 to x1 (either direction), mixed error control rel*|y| + abs
template <class F>
void integrate(F &&rhs, double *y, int n, double x0, double x1, double rel)
{
  static const double c2 = 1.0 / 5, c3 = 3.0 / 10, c4 = 4.0 / 5, c5 = 8.0 / 9;
  static const double a21 = 1.0 / 5, a31 = 3.0 / 40, a32 = 9.0 / 40, a41 = 44.0 / 45, a42 = -56.0 / 15, a43 = 32.0 / 9,
                      a51 = 19372.0 / 6561, a52 = -25360.0 / 2187, a53 = 64448.0 / 6561, a54 = -212.0 / 729,
                      a61 = 9017.0 / 3168, a62 = -355.0 / 33, a63 = 46732.0 / 5247, a64 = 49.0 / 176, a65 = -5103.0 / 18656,
                      b1 = 35.0 / 384, b3 = 500.0 / 1113, b4 = 125.0 / 192, b5 = -2187.0 / 6784, b6 = 11.0 / 84,
                      e1 = 71.0 / 57600, e3 = -71.0 / 16695, e4 = 71.0 / 1920, e5 = -17253.0 / 339200, e6 = 22.0 / 525,
                      e7 = -1.0 / 40;
  if (x1 == x0) return;
  const double dir = x1 > x0 ? 1.0 : -1.0, abs_tol = 1e-31;
  double x = x0, h = (x1 - x0) * 0.1;
  double k1[8], k2[8], k3[8], k4[8], k5[8], k6[8], k7[8], yt[8], yn[8];
  rhs(x, y, k1);
  for (int guard = 0; guard < 1000000; guard++) {
    if ((x + h - x1) * dir > 0.0) h = x1 - x;
    for (int i = 0; i < n; i++) yt[i] = y[i] + h * a21 * k1[i];
    rhs(x + c2 * h, yt, k2);
    for (int i = 0; i < n; i++) yt[i] = y[i] + h * (a31 * k1[i] + a32 * k2[i]);
    rhs(x + c3 * h, yt, k3);
    for (int i = 0; i < n; i++) yt[i] = y[i] + h * (a41 * k1[i] + a42 * k2[i] + a43 * k3[i]);
    rhs(x + c4 * h, yt, k4);
    for (int i = 0; i < n; i++) yt[i] = y[i] + h * (a51 * k1[i] + a52 * k2[i] + a53 * k3[i] + a54 * k4[i]);
    rhs(x + c5 * h, yt, k5);
    for (int i = 0; i < n; i++) yt[i] = y[i] + h * (a61 * k1[i] + a62 * k2[i] + a63 * k3[i] + a64 * k4[i] + a65 * k5[i]);
    rhs(x + h, yt, k6);
    for (int i = 0; i < n; i++) yn[i] = y[i] + h * (b1 * k1[i] + b3 * k3[i] + b4 * k4[i] + b5 * k5[i] + b6 * k6[i]);
    rhs(x + h, yn, k7);
    double err = 0.0;
    for (int i = 0; i < n; i++) {
      const double ei = h * (e1 * k1[i] + e3 * k3[i] + e4 * k4[i] + e5 * k5[i] + e6 * k6[i] + e7 * k7[i]);
      const double sc = abs_tol + rel * std::max(std::fabs(y[i]), std::fabs(yn[i]));
      err = std::max(err, std::fabs(ei) / sc);
    }
    if (err <= 1.0 || std::fabs(h) < 1e-14 * std::max(1.0, std::fabs(x))) {
      x += h;
      for (int i = 0; i < n; i++) { y[i] = yn[i]; k1[i] = k7[i]; }
      if ((x - x1) * dir >= 0.0) return;
    }
    const double fac = (err > 0.0) ? 0.9 * std::pow(err, -0.2) : 5.0;
    h *= std::min(5.0, std::max(0.2, fac));
  }
}

double cross_sign(const double a[2], const double b[2]) { return std::copysign(1.0, a[0] * b[1] - a[1] * b[0]); }

// periodic cubic spline of n+1 points f[0..n] with f[0] == f[n], step h: coefficients c[k][i], k = 0..3, i = 0..n
void spline_periodic3(int n, double h, const double *f, std::vector<double> &coef /* [n+1][4] */)
{
  // second derivatives M_i from the cyclic tridiagonal system M_{i-1} + 4 M_i + M_{i+1} = 6 (f_{i+1} - 2 f_i + f_{i-1}) / h^2
  std::vector<double> rhs(n), M(n), cp(n), dp(n), u(n), z(n);
  for (int i = 0; i < n; i++) {
    const double fm = f[(i + n - 1) % n], fp = f[(i + 1) % n];
    rhs[i] = 6.0 * (fp - 2.0 * f[i] + fm) / (h * h);
  }
  // Sherman-Morrison for the cyclic system with diagonal 4, off-diagonals 1
  const double gamma = -4.0;
  std::vector<double> bb(n, 4.0);
  bb[0] = 4.0 - gamma;
  bb[n - 1] = 4.0 - 1.0 / gamma;
  auto tridiag = [&](const std::vector<double> &r, std::vector<double> &x) {
    cp[0] = 1.0 / bb[0];
    dp[0] = r[0] / bb[0];
    for (int i = 1; i < n; i++) {
      const double m = bb[i] - cp[i - 1];
      cp[i] = 1.0 / m;
      dp[i] = (r[i] - dp[i - 1]) / m;
    }
    x[n - 1] = dp[n - 1];
    for (int i = n - 2; i >= 0; i--) x[i] = dp[i] - cp[i] * x[i + 1];
  };
  std::vector<double> uu(n, 0.0);
  uu[0] = gamma;
  uu[n - 1] = 1.0;
  tridiag(rhs, M);
  tridiag(uu, z);
  const double fact = (M[0] + M[n - 1] / gamma) / (1.0 + z[0] + z[n - 1] / gamma);
  for (int i = 0; i < n; i++) M[i] -= fact * z[i];
  coef.assign((size_t)(n + 1) * 4, 0.0);
  for (int i = 0; i <= n; i++) {
    const int i0 = i % n, i1 = (i + 1) % n;
    double *c = &coef[(size_t)4 * i];
    c[0] = f[i0];
    c[1] = (f[i1] - f[i0]) / h - h * (2.0 * M[i0] + M[i1]) / 6.0;
    c[2] = M[i0] / 2.0;
    c[3] = (M[i1] - M[i0]) / (6.0 * h);
  }
}

void lagrange(int npoi, double x, const double *xp, double *c0, double *c1)
{
  for (int i = 0; i < npoi; i++) {
    c0[i] = 1.0;
    for (int k = 0; k < npoi; k++)
      if (k != i) c0[i] = c0[i] * (x - xp[k]) / (xp[i] - xp[k]);
  }
  if (!c1) return;
  double dummy[16];
  for (int i = 0; i < npoi; i++) {
    for (int j = 0; j < npoi; j++) dummy[j] = 1.0;
    dummy[i] = 0.0;
    for (int k = 0; k < npoi; k++) {
      if (k == i) continue;
      const double fac = (x - xp[k]) / (xp[i] - xp[k]);
      for (int j = 0; j < npoi; j++) {
        if (j == k) dummy[j] = dummy[j] / (xp[i] - xp[k]);
        else dummy[j] = dummy[j] * fac;
      }
    }
    double s = 0.0;
    for (int j = 0; j < npoi; j++) s += dummy[j];
    c1[i] = s;
  }
}

struct FluxCoords {
  int nlabel = 500, ntheta = 500;
  double raxis = 0, zaxis = 0, psipol_max = 0, psitor_max = 0, theta0 = 0, sigma = 1, h_theta = 0;
  double x_point[2] = {0, 0};
  std::vector<double> rbeg, rsmall, qsaf;        // [nlabel], index = label - 1
  std::vector<double> psisurf, phitor;           // [0..nlabel], normalised
  std::vector<double> Rs, Zs, Bs, Gs;            // [nlabel][ntheta+1][4] periodic cubic spline over theta

  // magdata_in_symfluxcoord_ext, inp_label = 1
  void eval(double s, double theta, double &psi, double &q, double &sqrtg, double &bmod, double &R, double &dR_ds,
            double &dR_dt, double &Z, double &dZ_ds, double &dZ_dt) const
  {
    // binsrc(phitor(0:nlabel), 0, nlabel, s, ibeg)
    int imin = 0, imax = nlabel, i = 0;
    for (int k = 1; k <= nlabel; k++) {
      i = (imax - imin) / 2 + imin;
      if (phitor[i] > s) imax = i;
      else imin = i;
      if (imax == imin + 1) break;
    }
    int ibeg = std::max(1, imax - 2), iend = ibeg + 3;
    if (iend > nlabel) { iend = nlabel; ibeg = iend - 3; }
    double c0[4], c1[4];
    lagrange(4, s, &phitor[ibeg], c0, c1);
    psi = 0.0; q = 0.0;
    for (int k = 0; k < 4; k++) { psi += c0[k] * psisurf[ibeg + k]; q += c0[k] * qsaf[ibeg + k - 1]; }
    psi *= psipol_max;
    const double twopi = std::atan(1.0) * 8.0;
    double dth = std::fmod(theta, twopi);
    if (dth < 0.0) dth += twopi;
    dth = dth / h_theta;
    const int it = std::max(0, std::min(ntheta - 1, (int)dth));
    dth = (dth - (double)it) * h_theta;
    sqrtg = bmod = R = Z = dR_ds = dZ_ds = dR_dt = dZ_dt = 0.0;
    for (int k = 0; k < 4; k++) {
      const size_t o = ((size_t)(ibeg + k - 1) * (ntheta + 1) + it) * 4;
      auto val = [&](const std::vector<double> &a) { return ((a[o + 3] * dth + a[o + 2]) * dth + a[o + 1]) * dth + a[o]; };
      auto der = [&](const std::vector<double> &a) { return (a[o + 3] * 3.0 * dth + a[o + 2] * 2.0) * dth + a[o + 1]; };
      const double Rk = val(Rs), Zk = val(Zs);
      sqrtg += c0[k] * val(Gs);
      bmod += c0[k] * val(Bs);
      R += c0[k] * Rk; Z += c0[k] * Zk;
      dR_ds += c1[k] * Rk; dZ_ds += c1[k] * Zk;
      dR_dt += c0[k] * der(Rs); dZ_dt += c0[k] * der(Zs);
    }
  }
};

// theta_geom2theta_flux (SRC/points_2d.f90:254-373): the symmetry-flux angles at which the surface `s` has the geometrical
// poloidal angles theta_geom[0..n) around the magnetic axis (measured from the axis -> X-point ray when theta0_at_xpoint).
// The geometrical angle is sampled at ntheta_interp = 500 equidistant flux angles (+ nplag/2 wrapped samples either side),
// made monotonic across the 2 pi cut and inverted with nplag = 10 point Lagrange interpolation.
// One deviation: the geometrical angle 0 maps to the flux angle 0 exactly (both are the axis -> X-point ray, resp. the
// outboard horizontal ray, by construction; the reference carries this shortcut commented out, :323-326), so that the first
// vertex of every ring keeps the "theta = 0 is also theta = 2 pi" vertex convention of make_tetra_physics
// (SRC/tetra_physics_mod.f90:534-540) whatever the integrator tolerance (1e-9) left in the last digits.
int geom_to_flux_angles(const FluxCoords &F, bool theta0_at_xpoint, double s, const double *theta_geom, int n,
                        double *theta_flux, std::string &err)
{
  constexpr int NI = 500, NP = 10, NT = NI + NP;
  const double twopi = 2.0 * PI_T;  // points_2d.f90:4
  auto modulo = [](double a, double p) { double r = std::fmod(a, p); if (r != 0.0 && ((r < 0.0) != (p < 0.0))) r += p; return r; };
  double tf[NT], tg[NT];
  for (int i = 0; i < NT; i++) tf[i] = (double)(i - NP / 2) / (double)NI * twopi;
  for (int i = 0; i < NT; i++) {
    double psi, q, sqrtg, bmod, R, dR_ds, dR_dt, Z, dZ_ds, dZ_dt;
    F.eval(s, modulo(tf[i], twopi), psi, q, sqrtg, bmod, R, dR_ds, dR_dt, Z, dZ_ds, dZ_dt);
    const double a = std::atan2(Z - F.zaxis, R - F.raxis);
    tg[i] = modulo(theta0_at_xpoint ? a - F.theta0 : a, twopi) - twopi;
    if (i > 0) {
      if (tg[i] < tg[i - 1]) tg[i] = tg[i] + twopi * std::ceil((tg[i - 1] - tg[i]) / twopi);
      if (tg[i] - tg[i - 1] > PI_T) {
        err = "theta_geom2theta_flux: big jump in the geometrical angle (the map flux angle -> geometrical angle is not monotonic)";
        return GORILLA_ERR_DOMAIN;
      }
    }
  }
  for (int k = 0; k < n; k++) {
    double th = theta_geom[k];
    if (th == 0.0) { theta_flux[k] = 0.0; continue; }
    if (th < tg[NP / 2]) th = th + twopi;
    else if (th > tg[NI + NP / 2 - 1]) th = th - twopi;
    // binsrc(theta_geom_interp(nplag/2+1 : ntheta_interp+nplag/2), 1, ntheta_interp, theta, i)
    const double *p = &tg[NP / 2] - 1;  // 1-based view
    int imin = 1, imax = NI, i = 0;
    for (int it = 1; it <= NI - 1; it++) {
      i = (imax - imin) / 2 + imin;
      if (p[i] > th) imax = i;
      else imin = i;
      if (imax == imin + 1) break;
    }
    const int ig = imax + NP / 2;            // 1-based index into the full sample arrays
    const int first = ig - NP / 2 - 1;       // 0-based start of the nplag nodes
    double c[NP];
    lagrange(NP, th, &tg[first], c, nullptr);
    double sum = 0.0;
    for (int j = 0; j < NP; j++) sum += c[j] * tf[first + j];
    theta_flux[k] = modulo(sum, twopi);
  }
  return GORILLA_OK;
}

int build_flux_coords(const EfitField &f, bool theta0_at_xpoint, FluxCoords &F, std::string &err)
{
  const double relerr = 1e-9;
  const int nsurfmax = 10000, niter_axis = 20, nmap = 10, niter = 50, nstep_min = 10;
  const int nlabel = F.nlabel, ntheta = F.ntheta;
  const double rmn = f.rad.front(), rmx = f.rad.back(), zmn = f.zet.front(), zmx = f.zet.back();
  double dr_dphi = 0, dz_dphi = 0;
  auto rhs_axis = [&](double, const double *y, double *dy) {
    double Br, Bp, Bz, psi;
    f.field_eq(y[0], y[1], Br, Bp, Bz, psi);
    dy[0] = Br * y[0] / Bp; dy[1] = Bz * y[0] / Bp; dy[2] = y[0]; dy[3] = y[1];
  };
  auto rhs_surf = [&](double, const double *y, double *dy) {
    double Br, Bp, Bz, psi;
    f.field_eq(y[0], y[1], Br, Bp, Bz, psi);
    dy[0] = Br * y[0] / Bp; dy[1] = Bz * y[0] / Bp; dy[2] = y[0] * dy[1]; dy[3] = y[0] * y[1] * Br;
    dr_dphi = dy[0]; dz_dphi = dy[1];
  };
  // ---- magnetic axis: the average position of a field line converges to it
  double y[4] = {0.5 * (rmn + rmx), 0.5 * (zmn + zmx), 0, 0};
  for (int iter = 0; iter < niter_axis; iter++) {
    y[2] = y[3] = 0.0;
    for (int i = 0; i < nmap; i++) integrate(rhs_axis, y, 4, 0.0, 2.0 * PI_T, relerr);
    y[0] = y[2] / (2.0 * PI_T) / (double)nmap;
    y[1] = y[3] / (2.0 * PI_T) / (double)nmap;
  }
  const double raxis = y[0], zaxis = y[1];
  F.raxis = raxis; F.zaxis = zaxis;
  double Br, Bp, Bz, psi_axis;
  f.field_eq(raxis, zaxis, Br, Bp, Bz, psi_axis);
  double hbr = (rmx - raxis) / nsurfmax;
  {
    double p;
    f.field_eq(raxis + hbr, zaxis, Br, Bp, Bz, p);
  }
  const double sigma = std::copysign(1.0, Bz * Bp);
  F.sigma = sigma;
  const double h = 2.0 * PI_T / nstep_min;
  // ---- last closed surface: largest start radius whose line completes a poloidal turn inside the box
  auto closes = [&](int isurf) {
    double ys[4] = {raxis + hbr * isurf, zaxis, 0, 0};
    integrate(rhs_surf, ys, 4, 0.0, h, relerr);
    for (int half = 0; half < 2; half++) {
      const double sig = ys[1] - zaxis;
      int guard = 0;
      while (sig * (ys[1] - zaxis) > 0.0) {
        integrate(rhs_surf, ys, 4, 0.0, h, relerr);
        if (ys[0] < rmn || ys[0] > rmx || ys[1] < zmn || ys[1] > zmx) return false;
        if (++guard > 100000) return false;
      }
    }
    return true;
  };
  int lo = 1, hi = nsurfmax;
  if (!closes(lo)) { err = "EFIT flux coordinates: no closed flux surface next to the magnetic axis"; return GORILLA_ERR_DOMAIN; }
  while (hi - lo > 1) {
    const int mid = (lo + hi) / 2;
    if (closes(mid)) lo = mid;
    else hi = mid;
  }
  // reference: nsurf = (first failing isurf) - 1, then "last point is bad, remove it"
  const int nsurf = lo - 1;
  const double r_sep = raxis + hbr * lo;
  if (nsurf < 10) { err = "EFIT flux coordinates: separatrix search failed"; return GORILLA_ERR_DOMAIN; }
  hbr = hbr * (double)nsurf / (double)nlabel;
  // ---- X-point: where the outermost line moves slowest poloidally
  const double axis[2] = {raxis, zaxis};
  double theta_axis[2];
  {
    double ys[4] = {raxis + hbr * (double)nlabel, zaxis, 0, 0};
    theta_axis[0] = ys[0] - raxis; theta_axis[1] = ys[1] - zaxis;
    double prev[2] = {ys[0], ys[1]}, min_d = HUGE_VAL, sig_start = 1.0, sig_end = 1.0;
    int guard = 0;
    while (sig_start >= sig_end) {
      integrate(rhs_surf, ys, 4, 0.0, h * sigma, relerr);
      const double nd = (prev[0] - ys[0]) * (prev[0] - ys[0]) + (prev[1] - ys[1]) * (prev[1] - ys[1]);
      if (nd < min_d) { min_d = nd; F.x_point[0] = ys[0]; F.x_point[1] = ys[1]; }
      prev[0] = ys[0]; prev[1] = ys[1];
      sig_start = sig_end;
      const double d[2] = {ys[0] - axis[0], ys[1] - axis[1]};
      sig_end = cross_sign(theta_axis, d);
      if (++guard > 1000000) { err = "EFIT flux coordinates: X-point search did not terminate"; return GORILLA_ERR_DOMAIN; }
    }
  }
  if (theta0_at_xpoint) { theta_axis[0] = F.x_point[0] - raxis; theta_axis[1] = F.x_point[1] - zaxis; }
  else { theta_axis[0] = 1.0 * (r_sep - raxis); theta_axis[1] = 0.0 * (r_sep - raxis); }
  const double theta0 = std::atan2(theta_axis[1], theta_axis[0]);
  F.theta0 = theta0;
  // ---- flux functions on nlabel surfaces
  F.rbeg.assign(nlabel, 0.0); F.rsmall.assign(nlabel, 0.0); F.qsaf.assign(nlabel, 0.0);
  std::vector<double> psisurf(nlabel, 0.0), phitor(nlabel, 0.0);
  bool failed = false;
#pragma omp parallel for schedule(dynamic, 4)
  for (int isurf = 2; isurf <= nlabel; isurf++) {
    double drp = 0, dzp = 0;
    auto rhs = [&](double, const double *yy, double *dy) {
      double br, bp, bz, p;
      f.field_eq(yy[0], yy[1], br, bp, bz, p);
      dy[0] = br * yy[0] / bp; dy[1] = bz * yy[0] / bp; dy[2] = yy[0] * dy[1]; dy[3] = yy[0] * yy[1] * br;
      drp = dy[0]; dzp = dy[1];
    };
    const double frac = ((isurf - 1) * 1.0) / (nlabel - 1);
    double ys[4] = {raxis + theta_axis[0] * frac, zaxis + theta_axis[1] * frac, 0, 0};
    double phi_sep = 0.0, sig_start = 1.0 * sigma, sig_end = 1.0 * sigma;
    int guard = 0;
    while (sig_start * sigma >= sig_end * sigma) {
      integrate(rhs, ys, 4, 0.0, h, relerr);
      phi_sep += h;
      sig_start = sig_end;
      const double d[2] = {ys[0] - raxis, ys[1] - zaxis};
      sig_end = cross_sign(theta_axis, d);
      if (++guard > 1000000) { failed = true; break; }
    }
    for (int iter = 0; iter < niter; iter++) {  // Newton: land on the theta = 0 ray
      const double ya[2] = {ys[0] - raxis, ys[1] - zaxis};
      const double alpha = std::atan2(dzp, drp) - theta0, beta = std::atan2(ya[1], ya[0]) - theta0;
      const double phiout = std::sqrt(ya[0] * ya[0] + ya[1] * ya[1]) * std::fabs(std::sin(beta) / std::sin(alpha)) /
                            std::sqrt(drp * drp + dzp * dzp) * cross_sign(ya, theta_axis) * sigma;
      if (!(std::fabs(phiout) > 0.0) || !std::isfinite(phiout)) break;
      integrate(rhs, ys, 4, 0.0, phiout, relerr);
      phi_sep += phiout;
    }
    const double aiota = 2.0 * PI_T / phi_sep;
    double br, bp, bz, p;
    f.field_eq(ys[0], ys[1], br, bp, bz, p);
    F.rbeg[isurf - 1] = std::hypot(ys[0] - raxis, ys[1] - zaxis);
    F.rsmall[isurf - 1] = std::sqrt(std::fabs(ys[2]) / PI_T);
    F.qsaf[isurf - 1] = sigma / aiota;
    psisurf[isurf - 1] = p - psi_axis;
    phitor[isurf - 1] = ys[3] / (2.0 * PI_T);
  }
  if (failed) { err = "EFIT flux coordinates: a field line did not return to the theta = 0 ray"; return GORILLA_ERR_DOMAIN; }
  {
    double c0[4];
    lagrange(4, 0.0, &F.rbeg[1], c0, nullptr);
    F.qsaf[0] = 0.0;
    for (int k = 0; k < 4; k++) F.qsaf[0] += F.qsaf[1 + k] * c0[k];
  }
  // ---- R, Z, |B|, sqrt(g)/norm at ntheta equidistant flux angles per surface
  std::vector<double> Rst((size_t)nlabel * ntheta), Zst((size_t)nlabel * ntheta), Bst((size_t)nlabel * ntheta),
      Gst((size_t)nlabel * ntheta);
#pragma omp parallel for schedule(dynamic, 4)
  for (int isurf = 2; isurf <= nlabel; isurf++) {
    auto rhs = [&](double, const double *yy, double *dy) {
      double br, bp, bz, p;
      f.field_eq(yy[0], yy[1], br, bp, bz, p);
      dy[0] = br * yy[0] / bp; dy[1] = bz * yy[0] / bp; dy[2] = yy[0] * dy[1]; dy[3] = yy[0] * yy[1] * br;
    };
    const double phiout = 2.0 * PI_T * 1.0 * F.qsaf[isurf - 1] / ntheta;
    const double frac = ((isurf - 1) * 1.0) / (nlabel - 1);
    double ys[4] = {raxis + theta_axis[0] * frac, zaxis + theta_axis[1] * frac, 0, 0};
    for (int j = 0; j < ntheta; j++) {
      integrate(rhs, ys, 4, 0.0, phiout, relerr);
      double br, bp, bz, p;
      f.field_eq(ys[0], ys[1], br, bp, bz, p);
      const size_t o = (size_t)(isurf - 1) * ntheta + j;
      Rst[o] = ys[0]; Zst[o] = ys[1];
      Bst[o] = std::sqrt(br * br + bp * bp + bz * bz);
      Gst[o] = ys[0] / std::fabs(bp);
    }
  }
  {
    f.field_eq(raxis, zaxis, Br, Bp, Bz, psi_axis);
    for (int j = 0; j < ntheta; j++) {
      Rst[j] = raxis; Zst[j] = zaxis;
      Bst[j] = std::sqrt(Br * Br + Bp * Bp + Bz * Bz);
      Gst[j] = raxis / std::fabs(Bp);
    }
  }
  // ---- load_magdata_in_symfluxcoord: periodic splines over theta, normalised flux labels
  F.h_theta = (std::atan(1.0) * 8.0) / (double)ntheta;
  auto spline_all = [&](const std::vector<double> &src, std::vector<double> &dst) {
    dst.assign((size_t)nlabel * (ntheta + 1) * 4, 0.0);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nlabel; i++) {
      std::vector<double> fpts(ntheta + 1), coef;
      for (int j = 1; j <= ntheta; j++) fpts[j] = src[(size_t)i * ntheta + j - 1];
      fpts[0] = fpts[ntheta];
      spline_periodic3(ntheta, F.h_theta, fpts.data(), coef);
      std::copy(coef.begin(), coef.end(), dst.begin() + (size_t)i * (ntheta + 1) * 4);
    }
  };
  spline_all(Rst, F.Rs); spline_all(Zst, F.Zs); spline_all(Bst, F.Bs); spline_all(Gst, F.Gs);
  F.psisurf.assign(nlabel + 1, 0.0); F.phitor.assign(nlabel + 1, 0.0);
  for (int i = 1; i <= nlabel; i++) { F.psisurf[i] = psisurf[i - 1]; F.phitor[i] = phitor[i - 1]; }
  F.psipol_max = F.psisurf[nlabel];
  F.psitor_max = F.phitor[nlabel];
  for (auto &v : F.psisurf) v = v / F.psipol_max;
  for (auto &v : F.phitor) v = v / F.psitor_max;
  return GORILLA_OK;
}

}  // namespace

int build_efit_flux(const gorilla_grid_settings &gs, const gorilla_settings &st, Mesh &m, std::string &err)
{
  if (st.coord_system != 2) { err = "grid_kind 2 with coord_system 1 is not built by this library (use coord_system = 2)"; return GORILLA_ERR_UNSUPPORTED; }
  if (gs.theta_geom_flux != 1 && gs.theta_geom_flux != 2) { err = "grid_kind 2: theta_geom_flux must be 1 (flux angle) or 2 (geometrical angle)"; return GORILLA_ERR_ARG; }
  if (gs.n1 < 1 || gs.n2 < 3 || gs.n3 < 3) { err = "field-aligned grid needs n1 >= 1, n2 >= 3, n3 >= 3"; return GORILLA_ERR_ARG; }
  if (!(gs.sfc_s_min > 0.0 && gs.sfc_s_min < 1.0)) { err = "sfc_s_min must be in (0, 1)"; return GORILLA_ERR_ARG; }
  EfitField f;
  f.nwindow_r = gs.nwindow_r; f.nwindow_z = gs.nwindow_z;
  int rc = f.load_efit(gs.g_file_filename, err);
  if (rc) return rc;
  if (gs.convex_wall_filename && gs.convex_wall_filename[0]) {
    rc = f.load_convex_wall(gs.convex_wall_filename, err);
    if (rc) return rc;
  }
  FluxCoords F;
  rc = build_flux_coords(f, gs.theta0_at_xpoint != 0.0, F, err);
  if (rc) return rc;

  // set_grid_size: extra, logarithmically spaced rings next to the axis when the first regular ring is far from s_min
  const double s_min = gs.sfc_s_min;
  const double s_ratio = (s_min + (1.0 - s_min) / (double)gs.n1) / s_min;
  const int n_extra = (int)std::fabs(std::log(s_ratio) / std::log(10.0));
  const int n1 = gs.n1 + n_extra, n2 = gs.n2, n3 = gs.n3;
  m.grid_kind = 2;
  m.coord_system = 2;
  m.grid_size[0] = n1; m.grid_size[1] = n2; m.grid_size[2] = n3;
  m.n_field_periods = gs.boole_n_field_periods ? 1 : gs.n_field_periods_manual;
  m.sfc_s_min = s_min;
  m.psitor_max = F.psitor_max;
  make_field_aligned_topology(m, n1, n2, n3);
  std::vector<double> r_frac(n1 + 1, 0.0);  // 1-based ring index
  if (gs.i_radial_spacing == 2) {
    const double q0 = std::sqrt(s_min);
    for (int i = 1; i <= n1 - n_extra; i++) {
      const double q = q0 + (double)i * (1.0 - q0) / (double)(n1 - n_extra);
      r_frac[n_extra + i] = q * q;
    }
  } else {
    for (int i = 1; i <= n1 - n_extra; i++) r_frac[n_extra + i] = s_min + ((double)i * (1.0 - s_min)) / (double)(n1 - n_extra);
  }
  for (int i = 1; i <= n_extra; i++)
    r_frac[i] = std::exp(std::log(s_min) + (double)i * (std::log(r_frac[n_extra + 1]) - std::log(s_min)) / (double)(n_extra + 1));

  // poloidal angles of the vertices of each ring (create_points_2d, SRC/points_2d.f90:135-149)
  std::vector<double> ring_theta((size_t)(n1 + 1) * n3);
  {
    std::vector<double> frac(n3);
    for (int j = 0; j < n3; j++) frac[j] = ((double)j / (double)n3) * 2.0 * PI_T;
    bool bad = false;
#pragma omp parallel for schedule(dynamic, 1)
    for (int ring = 0; ring <= n1; ring++) {
      double *out = &ring_theta[(size_t)ring * n3];
      if (gs.theta_geom_flux == 1) {
        for (int j = 0; j < n3; j++) out[j] = frac[j];
      } else {
        std::string e;
        if (geom_to_flux_angles(F, gs.theta0_at_xpoint != 0.0, ring == 0 ? s_min : r_frac[ring], frac.data(), n3, out, e)) {
#pragma omp critical
          { bad = true; err = e; }
        }
      }
    }
    if (bad) return GORILLA_ERR_DOMAIN;
  }
  const int64_t vps = (int64_t)(n1 + 1) * n3;
  m.verts_sthetaphi.assign((size_t)m.nvert * 3, 0.0);
  m.verts_rphiz.assign((size_t)m.nvert * 3, 0.0);
  VertexFields vf;
  vf.resize((size_t)m.nvert, true, false);
#pragma omp parallel for schedule(static)
  for (int64_t iv = 0; iv < m.nvert; iv++) {
    const int slice = (int)(iv / vps), ring = (int)((iv % vps) / n3), j = (int)(iv % n3);
    const double s = (ring == 0) ? s_min : r_frac[ring];
    const double theta = ring_theta[(size_t)ring * n3 + j];
    const double phi = slice == 0 ? 0.0 : (2.0 * PI_T / m.n_field_periods * slice) / n2;
    double psi, q, sqrtg, b1, R, dR_ds, dR_dt, Z, dZ_ds, dZ_dt;
    F.eval(s, theta, psi, q, sqrtg, b1, R, dR_ds, dR_dt, Z, dZ_ds, dZ_dt);
    double *vs = &m.verts_sthetaphi[3 * iv], *vr = &m.verts_rphiz[3 * iv];
    vs[0] = s; vs[1] = theta; vs[2] = phi;
    vr[0] = R; vr[1] = phi; vr[2] = Z;
    double Br, Bp, Bz, psif;
    f.field(R, Z, Br, Bp, Bz, psif);
    const double bmod = std::sqrt(Br * Br + Bp * Bp + Bz * Bz) * m.bmod_multiplier;
    vf.A_x1[iv] = 0.0;
    vf.A_x2[iv] = s * F.psitor_max;
    vf.A_x3[iv] = psi;
    if (st.boole_helical_pert)  // analytical perturbation (tetra_physics_mod.f90:1158-1161)
      vf.A_x3[iv] = psi + psi * st.helical_pert_eps_Aphi * std::cos(st.helical_pert_m_fourier * theta + st.helical_pert_n_fourier * phi);
    vf.bmod[iv] = bmod;
    vf.h_x1[iv] = (Br * dR_ds + Bz * dZ_ds) / bmod;
    vf.h_x2[iv] = (Br * dR_dt + Bz * dZ_dt) / bmod;
    vf.h_x3[iv] = (Bp * R) / bmod;
    vf.dR_ds[iv] = dR_ds; vf.dZ_ds[iv] = dZ_ds;
    vf.phi_elec[iv] = vf.A_x2[iv] * st.eps_Phi;
  }
  {
    double psi, q, sqrtg, b1, dR_ds, dR_dt, dZ_ds, dZ_dt;
    F.eval(0.0, 2.706, psi, q, sqrtg, b1, m.mag_axis_R0, dR_ds, dR_dt, m.mag_axis_Z0, dZ_ds, dZ_dt);
  }
  m.Rmin = m.Rmax = m.Zmin = m.Zmax = 0.0;
  apply_vertex_noise(m, st, vf);
  linearise_tetrahedra(m, vf);
  check_tetra_overlaps(m);
  return GORILLA_OK;
}

}  // namespace gbhost
