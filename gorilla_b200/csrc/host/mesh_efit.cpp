// mesh_efit.cpp -- axisymmetric tokamak equilibria (EFIT g-file, WEST table) and grid_kind = 1: rectangular
// (R, phi, Z) grid over the equilibrium box.  The field object is shared with the SOLEDGE3X builder.
//
// Reference: field / field_eq              SRC/field_divB0.f90:19-168, SRC/bdivfree.f90:1086-1274
//            read_eqfile2 / read_eqfile_west / set_eqcoords / spline_fpol / splint_fpol
//                                          SRC/utils_bdivfree.f90:713-840,935-959
//            spl_five_reg / s2dcut / spline   SRC/spline5_RZ.f90:7-322 (quintic spline on an equidistant mesh)
//            stretch_coords                SRC/bdivfree.f90:974-1054
//            vector_potential_rphiz        SRC/tetra_physics_mod.f90:1076-1109
//            make_tetra_grid case(1)       SRC/tetra_grid_mod.f90:73-95
// Restated, not transcribed: same interpolant (end conditions, sweep order) so that vertex fields agree with
// the reference to round-off; the hot path never sees this code.  The moving-average psi filter of field_divB0.inp
// (nwindow_r, nwindow_z) is applied to the raw table as in the reference; every input the reference ships sets 0 0.
#include "mesh_efit.hpp"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

namespace gbhost {

// ---- quintic spline on an equidistant mesh ---------------------------------------------------------------
// Input a[0..n-1] (values), output b..f: s(x) = a_i + b_i t + c_i t^2 + d_i t^3 + e_i t^4 + f_i t^5, t = x - x_i.
// End conditions: the odd/even derivatives at both ends are those of the quintic through the six outermost points
// (3x3 Cramer systems below); the interior follows from two passes of a factorised penta-diagonal solve whose
// characteristic roots are 13 +- sqrt(105).
namespace {
struct Edge3 {  // solves [[a11 a12 a13],[a21 a22 a23],[a31 a32 a33]] x = rhs by Cramer's rule, same term order
  double a11, a12, a13, a21, a22, a23, a31, a32, a33, det;
  void setdet() { det = a11 * a22 * a33 + a12 * a23 * a31 + a13 * a21 * a32 - a12 * a21 * a33 - a13 * a22 * a31 - a11 * a23 * a32; }
  void solve(double b1, double b2, double b3, double &x1, double &x2, double &x3) const
  {
    x1 = (b1 * a22 * a33 + a12 * a23 * b3 + a13 * b2 * a32 - a12 * b2 * a33 - a13 * a22 * b3 - b1 * a23 * a32) / det;
    x2 = (a11 * b2 * a33 + b1 * a23 * a31 + a13 * a21 * b3 - b1 * a21 * a33 - a13 * b2 * a31 - a11 * a23 * b3) / det;
    x3 = (a11 * a22 * b3 + a12 * b2 * a31 + b1 * a21 * a32 - a12 * a21 * b3 - b1 * a22 * a31 - a11 * b2 * a32) / det;
  }
};
}  // namespace

void spl_five_reg(int n, double h, const double *a, double *b, double *c, double *d, double *e, double *f)
{
  const double rhop = 13.0 + std::sqrt(105.0), rhom = 13.0 - std::sqrt(105.0);
  // odd part (b, d, f) from differences, even part (a, c, e) from sums of the mirrored point pairs
  Edge3 odd{1.0, 1.0 / 4.0, 1.0 / 16.0, 3.0, 27.0 / 4.0, 9.0 * 27.0 / 16.0, 5.0, 125.0 / 4.0, std::pow(5.0, 5) / 16.0, 0};
  odd.setdet();
  Edge3 even{2.0, 1.0 / 2.0, 1.0 / 8.0, 2.0, 9.0 / 2.0, 81.0 / 8.0, 2.0, 25.0 / 2.0, 625.0 / 8.0, 0};
  even.setdet();
  double bbeg, dbeg, fbeg, bend, dend, fend, abeg, cbeg, ebeg, aend, cend, eend;
  odd.solve(a[3] - a[2], a[4] - a[1], a[5] - a[0], bbeg, dbeg, fbeg);
  odd.solve(a[n - 3] - a[n - 4], a[n - 2] - a[n - 5], a[n - 1] - a[n - 6], bend, dend, fend);
  even.solve(a[3] + a[2], a[4] + a[1], a[5] + a[0], abeg, cbeg, ebeg);
  even.solve(a[n - 3] + a[n - 4], a[n - 2] + a[n - 5], a[n - 1] + a[n - 6], aend, cend, eend);
  (void)bbeg; (void)dbeg; (void)abeg; (void)cbeg; (void)bend; (void)aend; (void)cend;

  std::vector<double> alp(n + 1), bet(n + 1), gam(n + 1);
  // 1-based indexing as in the algorithm's description: A(i) = a[i-1]
  auto A = [&](int i) { return a[i - 1]; };
  double *E = e - 1, *F = f - 1, *D = d - 1, *C = c - 1, *B = b - 1;
  alp[1] = 0.0;
  bet[1] = ebeg * (2.0 + rhom) - 5.0 * fbeg * (3.0 + 1.5 * rhom);
  for (int i = 1; i <= n - 4; i++) {
    const int ip1 = i + 1;
    alp[ip1] = -1.0 / (rhop + alp[i]);
    bet[ip1] = alp[ip1] * (bet[i] - 5.0 * (A(i + 4) - 4.0 * A(i + 3) + 6.0 * A(i + 2) - 4.0 * A(ip1) + A(i)));
  }
  gam[n - 2] = eend * (2.0 + rhom) + 5.0 * fend * (3.0 + 1.5 * rhom);
  for (int i = n - 3; i >= 1; i--) gam[i] = gam[i + 1] * alp[i] + bet[i];
  alp[1] = 0.0;
  bet[1] = ebeg - 2.5 * 5.0 * fbeg;
  for (int i = 1; i <= n - 2; i++) {
    const int ip1 = i + 1;
    alp[ip1] = -1.0 / (rhom + alp[i]);
    bet[ip1] = alp[ip1] * (bet[i] - gam[i]);
  }
  E[n] = eend + 2.5 * 5.0 * fend;
  E[n - 1] = E[n] * alp[n - 1] + bet[n - 1];
  F[n - 1] = (E[n] - E[n - 1]) / 5.0;
  E[n - 2] = E[n - 1] * alp[n - 2] + bet[n - 2];
  F[n - 2] = (E[n - 1] - E[n - 2]) / 5.0;
  D[n - 2] = dend + 1.5 * 4.0 * eend + 1.5 * 1.5 * 10.0 * fend;
  for (int i = n - 3; i >= 1; i--) {
    E[i] = E[i + 1] * alp[i] + bet[i];
    F[i] = (E[i + 1] - E[i]) / 5.0;
    D[i] = (A(i + 3) - 3.0 * A(i + 2) + 3.0 * A(i + 1) - A(i)) / 6.0 -
           (E[i + 3] + 27.0 * E[i + 2] + 93.0 * E[i + 1] + 59.0 * E[i]) / 30.0;
    C[i] = 0.5 * (A(i + 2) + A(i)) - A(i + 1) - 0.5 * D[i + 1] - 2.5 * D[i] -
           0.1 * (E[i + 2] + 18.0 * E[i + 1] + 31.0 * E[i]);
    B[i] = A(i + 1) - A(i) - C[i] - D[i] - 0.2 * (4.0 * E[i] + E[i + 1]);
  }
  for (int i = n - 3; i <= n; i++) {
    B[i] = B[i - 1] + 2.0 * C[i - 1] + 3.0 * D[i - 1] + 4.0 * E[i - 1] + 5.0 * F[i - 1];
    C[i] = C[i - 1] + 3.0 * D[i - 1] + 6.0 * E[i - 1] + 10.0 * F[i - 1];
    D[i] = D[i - 1] + 4.0 * E[i - 1] + 10.0 * F[i - 1];
    if (i != n) F[i] = A(i + 1) - A(i) - B[i] - C[i] - D[i] - E[i];
  }
  F[n] = F[n - 1];
  double fac = 1.0 / h;
  for (int i = 0; i < n; i++) b[i] *= fac;
  fac = fac / h;
  for (int i = 0; i < n; i++) c[i] *= fac;
  fac = fac / h;
  for (int i = 0; i < n; i++) d[i] *= fac;
  fac = fac / h;
  for (int i = 0; i < n; i++) e[i] *= fac;
  fac = fac / h;
  for (int i = 0; i < n; i++) f[i] *= fac;
}

// ---- 2-D tensor-product quintic spline on the full rectangle (s2dcut with imi = jmi = 1) ---------------------
void Spline2D::build(int nx_, int ny_, double hx_, double hy_, const std::vector<double> &fxy /* f(i,j) at [i + nx*j] */)
{
  nx = nx_; ny = ny_; hx = hx_; hy = hy_;
  spl.assign((size_t)36 * nx * ny, 0.0);
  const int nmax = std::max(nx, ny);
  std::vector<double> ai(nmax), bi(nmax), ci(nmax), di(nmax), ei(nmax), fi(nmax);
  // along y for every x: coefficient (1, l)
  for (int i = 0; i < nx; i++) {
    for (int j = 0; j < ny; j++) ai[j] = fxy[i + (size_t)nx * j];
    spl_five_reg(ny, hy, ai.data(), bi.data(), ci.data(), di.data(), ei.data(), fi.data());
    for (int j = 0; j < ny; j++) {
      double *s = at(i, j);
      s[0 + 6 * 0] = ai[j]; s[0 + 6 * 1] = bi[j]; s[0 + 6 * 2] = ci[j];
      s[0 + 6 * 3] = di[j]; s[0 + 6 * 4] = ei[j]; s[0 + 6 * 5] = fi[j];
    }
  }
  // along x for every y and every y-power l: coefficients (2..6, l)
  for (int j = 0; j < ny; j++)
    for (int l = 0; l < 6; l++) {
      for (int i = 0; i < nx; i++) ai[i] = at(i, j)[0 + 6 * l];
      spl_five_reg(nx, hx, ai.data(), bi.data(), ci.data(), di.data(), ei.data(), fi.data());
      for (int i = 0; i < nx; i++) {
        double *s = at(i, j);
        s[1 + 6 * l] = bi[i]; s[2 + 6 * l] = ci[i]; s[3 + 6 * l] = di[i]; s[4 + 6 * l] = ei[i]; s[5 + 6 * l] = fi[i];
      }
    }
}

void Spline2D::eval(double x0, double y0, double xb, double yb, double &u, double &ux, double &uy, double &uxx, double &uxy,
                    double &uyy) const
{
  int kx = (int)((xb - x0) / hx) + 1;
  kx = std::min(nx, std::max(1, kx));
  int ky = (int)((yb - y0) / hy) + 1;
  ky = std::min(ny, std::max(1, ky));
  const double dx = xb - (x0 + (kx - 1) * hx), dy = yb - (y0 + (ky - 1) * hy);
  const double *s = at(kx - 1, ky - 1);
  double a[6], ax[6], axx[6];
  for (int l = 0; l < 6; l++) {
    const double *q = s + 6 * l;
    a[l] = q[0] + dx * (q[1] + dx * (q[2] + dx * (q[3] + dx * (q[4] + dx * q[5]))));
    ax[l] = q[1] + dx * (2.0 * q[2] + dx * (3.0 * q[3] + dx * (4.0 * q[4] + dx * 5.0 * q[5])));
    axx[l] = 2.0 * q[2] + dx * (6.0 * q[3] + dx * (12.0 * q[4] + dx * (20.0 * q[5])));
  }
  u = a[0] + dy * (a[1] + dy * (a[2] + dy * (a[3] + dy * (a[4] + dy * a[5]))));
  ux = ax[0] + dy * (ax[1] + dy * (ax[2] + dy * (ax[3] + dy * (ax[4] + dy * ax[5]))));
  uy = a[1] + dy * (2.0 * a[2] + dy * (3.0 * a[3] + dy * (4.0 * a[4] + dy * 5.0 * a[5])));
  uxx = axx[0] + dy * (axx[1] + dy * (axx[2] + dy * (axx[3] + dy * (axx[4] + dy * axx[5]))));
  uxy = ax[1] + dy * (2.0 * ax[2] + dy * (3.0 * ax[3] + dy * (4.0 * ax[4] + dy * 5.0 * ax[5])));
  uyy = 2.0 * a[2] + dy * (6.0 * a[3] + dy * (12.0 * a[4] + dy * 20.0 * a[5]));
}

// ---- equilibrium files ----------------------------------------------------------------------------------------
namespace {
bool read_all(const char *path, std::string &out)
{
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  std::ostringstream ss;
  ss << f.rdbuf();
  out = ss.str();
  return true;
}

// g-file numbers are written with format 5(e16.9): fixed 16-character fields that may touch each other
struct FixedReader {
  std::vector<std::string> lines;
  size_t line = 0, col = 0;
  bool next(double &v)
  {
    while (line < lines.size()) {
      const std::string &L = lines[line];
      if (col + 16 <= L.size() || (col < L.size() && L.find_first_not_of(" \r", col) != std::string::npos)) {
        std::string field = L.substr(col, 16);
        col += 16;
        if (field.find_first_not_of(" \r") == std::string::npos) continue;
        v = std::strtod(field.c_str(), nullptr);
        return true;
      }
      line++;
      col = 0;
    }
    return false;
  }
  void next_line() { if (col != 0) { line++; col = 0; } }
};
}  // namespace

// window_filter (utils_bdivfree.f90:859-872): centred moving average, the window shrinks towards the ends of the row
void EfitField::filter_psi(std::vector<double> &psi) const
{
  if (nwindow_r <= 0 && nwindow_z <= 0) return;  // windows of 0: sum of one element / 1, the identity
  std::vector<double> psi0(psi.size());
  const int nwr = std::max(nwindow_r, 0), nwz = std::max(nwindow_z, 0);
  for (int iz = 0; iz < nzet; iz++)
    for (int i = 1; i <= nrad; i++) {
      const int nwa = std::min(nwr, std::min(i - 1, nrad - i));
      double s = 0.0;
      for (int k = i - nwa; k <= i + nwa; k++) s += psi[(k - 1) + (size_t)nrad * iz];
      psi0[(i - 1) + (size_t)nrad * iz] = s / (double)(2 * nwa + 1);
    }
  for (int ir = 0; ir < nrad; ir++)
    for (int i = 1; i <= nzet; i++) {
      const int nwa = std::min(nwz, std::min(i - 1, nzet - i));
      double s = 0.0;
      for (int k = i - nwa; k <= i + nwa; k++) s += psi0[ir + (size_t)nrad * (k - 1)];
      psi[ir + (size_t)nrad * (i - 1)] = s / (double)(2 * nwa + 1);
    }
}

int EfitField::load_efit(const char *path, std::string &err)
{
  std::string txt;
  if (!path || !read_all(path, txt)) { err = std::string("cannot read g-file ") + (path ? path : "(null)"); return GORILLA_ERR_IO; }
  FixedReader R;
  {
    std::istringstream ss(txt);
    std::string l;
    while (std::getline(ss, l)) R.lines.push_back(l);
  }
  if (R.lines.empty() || R.lines[0].size() < 60) { err = "g-file: header line too short"; return GORILLA_ERR_IO; }
  // format(6a8,3i4)
  const std::string &h = R.lines[0];
  nrad = std::atoi(h.substr(52, 4).c_str());
  nzet = std::atoi(h.substr(56, 4).c_str());
  if (nrad < 6 || nzet < 6) { err = "g-file: bad grid dimensions"; return GORILLA_ERR_IO; }
  R.line = 1;
  double v[20];
  for (int i = 0; i < 20; i++)
    if (!R.next(v[i])) { err = "g-file: early EOF in the header block"; return GORILLA_ERR_IO; }
  const double xdim = v[0], zdim = v[1], rzero = v[2], r1 = v[3], zmid = v[4];
  double rmaxis = v[5], zmaxis = v[6], psi_axis_in = v[7], psi_sep_in = v[8];
  const double bt0 = v[9];
  // third/fourth lines repeat the axis values (read order as in read_eqfile2)
  psi_axis_in = v[11]; rmaxis = v[13]; zmaxis = v[15]; psi_sep_in = v[17];
  std::vector<double> fpol(nrad), dummy(nrad);
  auto rd = [&](std::vector<double> &a) {
    R.next_line();
    for (auto &x : a)
      if (!R.next(x)) return false;
    return true;
  };
  std::vector<double> psi_in((size_t)nrad * nzet);
  if (!rd(fpol) || !rd(dummy) || !rd(dummy) || !rd(dummy) || !rd(psi_in)) { err = "g-file: early EOF in the profile block"; return GORILLA_ERR_IO; }
  rad.resize(nrad); zet.resize(nzet);
  for (int j = 1; j <= nrad; j++) rad[j - 1] = r1 + (j - 1) * (xdim / (nrad - 1));
  const double z1 = zmid - zdim / 2.0;
  for (int k = 1; k <= nzet; k++) zet[k - 1] = z1 + (k - 1) * (zdim / (nzet - 1));
  // field_eq, use_fpol branch (:1133-1139, :1184-1191)
  use_fpol = true;
  btf = bt0; rtf = rzero;
  double psib = -psi_axis_in;
  psi_sep = (psi_sep_in - psi_axis_in) * 1.0e8;
  splfpol.assign((size_t)6 * nrad, 0.0);
  {
    std::vector<double> a(nrad), b(nrad), c(nrad), d(nrad), e(nrad), f(nrad);
    for (int i = 0; i < nrad; i++) a[i] = fpol[i] * 1.0e6;
    hfpol = 1.0 / (double)(nrad - 1);
    spl_five_reg(nrad, hfpol, a.data(), b.data(), c.data(), d.data(), e.data(), f.data());
    for (int i = 0; i < nrad; i++) {
      double *s = &splfpol[(size_t)6 * i];
      s[0] = a[i]; s[1] = b[i]; s[2] = c[i]; s[3] = d[i]; s[4] = e[i]; s[5] = f[i];
    }
  }
  filter_psi(psi_in);
  for (auto &x : rad) x = x * 1.0e2;
  for (auto &x : zet) x = x * 1.0e2;
  rtf = rtf * 1.0e2;
  for (auto &x : psi_in) x = x * 1.0e8;
  psib = psib * 1.0e8;
  btf = btf * 1.0e4;
  for (auto &x : psi_in) x = x + psib;
  axis_R = rmaxis * 1.0e2; axis_Z = zmaxis * 1.0e2;
  hrad = rad[1] - rad[0];
  hzet = zet[1] - zet[0];
  psi_spl.build(nrad, nzet, hrad, hzet, psi_in);
  return GORILLA_OK;
}

int EfitField::load_west(const char *path, std::string &err)
{
  std::string txt;
  if (!path || !read_all(path, txt)) { err = std::string("cannot read equilibrium file ") + (path ? path : "(null)"); return GORILLA_ERR_IO; }
  std::istringstream ss(txt);
  double b;
  if (!(ss >> nrad >> nzet >> b) || nrad < 6 || nzet < 6) { err = "WEST equilibrium: bad header"; return GORILLA_ERR_IO; }
  rad.resize(nrad); zet.resize(nzet);
  for (auto &x : rad) if (!(ss >> x)) { err = "WEST equilibrium: early EOF"; return GORILLA_ERR_IO; }
  for (auto &x : zet) if (!(ss >> x)) { err = "WEST equilibrium: early EOF"; return GORILLA_ERR_IO; }
  std::vector<double> psi_in((size_t)nrad * nzet);
  for (int ir = 0; ir < nrad; ir++)
    for (int iz = 0; iz < nzet; iz++)
      if (!(ss >> psi_in[ir + (size_t)nrad * iz])) { err = "WEST equilibrium: early EOF in psi"; return GORILLA_ERR_IO; }
  use_fpol = false;
  rtf = 0.5 * (rad[0] + rad[nrad - 1]);
  btf = b / rtf;
  filter_psi(psi_in);
  for (auto &x : rad) x = x * 1.0e2;
  for (auto &x : zet) x = x * 1.0e2;
  rtf = rtf * 1.0e2;
  for (auto &x : psi_in) x = x * 1.0e8;
  btf = btf * 1.0e4;
  for (auto &x : psi_in) x = x + 0.0;  // psib = 0
  axis_R = 240.0; axis_Z = 0.0;        // hard coded for grid_kind 4 (tetra_physics_mod.f90:296-298)
  hrad = rad[1] - rad[0];
  hzet = zet[1] - zet[0];
  psi_spl.build(nrad, nzet, hrad, hzet, psi_in);
  return GORILLA_OK;
}

int EfitField::load_convex_wall(const char *path, std::string &err)
{
  std::string txt;
  if (!path || !read_all(path, txt)) { err = std::string("cannot read convex wall file ") + (path ? path : "(null)"); return GORILLA_ERR_IO; }
  for (auto &ch : txt)
    if (ch == 'd' || ch == 'D') ch = 'e';  // Fortran exponent letter
  std::istringstream ss(txt);
  std::vector<double> rw, zw;
  std::string line;
  while ((int)rw.size() < 100 && std::getline(ss, line)) {  // list-directed read: two numbers per record, rest ignored
    std::istringstream ls(line);
    double r, z;
    if (!(ls >> r >> z)) break;
    rw.push_back(r); zw.push_back(z);
  }
  if (rw.size() < 3) { err = "convex wall: fewer than 3 points"; return GORILLA_ERR_IO; }
  rw.push_back(rw[0]); zw.push_back(zw[0]);
  const int nrz = (int)rw.size();
  const double pi = 3.14159265358979;
  wall_R0 = (*std::max_element(rw.begin(), rw.end()) + *std::min_element(rw.begin(), rw.end())) * 0.5;
  std::vector<double> tht_w(nrz);
  for (int i = 0; i < nrz; i++) {
    tht_w[i] = std::atan2(zw[i], rw[i] - wall_R0);
    if (tht_w[i] < 0.0) tht_w[i] = tht_w[i] + 2.0 * pi;
  }
  const int nt = 360;
  wall_htht = 2.0 * pi / (nt - 1);
  rho_wall.assign(nt, 0.0);
  tht_wall.assign(nt, 0.0);
  double Rw = 0.0, Zw = 0.0;
  for (int i = 2; i <= nt; i++) {
    const double t = wall_htht * (i - 1);
    tht_wall[i - 1] = t;
    for (int j = 0; j < nrz - 1; j++) {
      if (t >= tht_w[j] && t <= tht_w[j + 1]) {
        if (std::fabs((rw[j + 1] - rw[j]) / rw[j]) > 1.e-3f) {
          const double a = (zw[j + 1] - zw[j]) / (rw[j + 1] - rw[j]);
          const double b = zw[j] - a * (rw[j] - wall_R0);
          Rw = b / (std::tan(t) - a) + wall_R0;
          Zw = a * (Rw - wall_R0) + b;
        } else {
          const double a = (rw[j + 1] - rw[j]) / (zw[j + 1] - zw[j]);
          const double b = rw[j] - wall_R0 - a * zw[j];
          Zw = b / (1.0 / std::tan(t) - a);
          Rw = a * Zw + b + wall_R0;
        }
      }
    }
    rho_wall[i - 1] = std::sqrt((Rw - wall_R0) * (Rw - wall_R0) + Zw * Zw);
  }
  tht_wall[0] = 0.0;
  rho_wall[0] = rho_wall[nt - 1];
  have_wall = true;
  return GORILLA_OK;
}

void EfitField::stretch_coords(double r, double z, double &rm, double &zm) const
{
  rm = r; zm = z;
  if (!have_wall) return;
  const double pi = 3.14159265358979, delta = 1.0;
  const int nt = (int)rho_wall.size();
  double rho = std::sqrt((r - wall_R0) * (r - wall_R0) + z * z);
  double tht = std::atan2(z, r - wall_R0);
  if (tht < 0.0) tht = tht + 2.0 * pi;
  int i = ((int)(tht / wall_htht)) % (nt - 1);
  if (i < 0) i += nt - 1;
  const double rho_c = (rho_wall[i + 1] - rho_wall[i]) / (tht_wall[i + 1] - tht_wall[i]) * (tht - tht_wall[i]) + rho_wall[i];
  if (rho >= rho_c) {
    rho = rho_c + delta * std::atan2(rho - rho_c, delta);
    rm = rho * std::cos(tht) + wall_R0;
    zm = rho * std::sin(tht);
  }
}

void EfitField::field(double r, double z, double &Br, double &Bp, double &Bz, double &psif) const
{
  double rm, zm;
  stretch_coords(r, z, rm, zm);
  field_eq(rm, zm, Br, Bp, Bz, psif);
}

void EfitField::field_eq(double rm, double zm, double &Br, double &Bp, double &Bz, double &psif) const
{
  const double rrr = std::max(rad[0], std::min(rad[nrad - 1], rm));
  const double zzz = std::max(zet[0], std::min(zet[nzet - 1], zm));
  double dpdr, dpdz, d2r, d2rz, d2z;
  psi_spl.eval(rad[0], zet[0], rrr, zzz, psif, dpdr, dpdz, d2r, d2rz, d2z);
  Br = -dpdz / rrr;
  Bz = dpdr / rrr;
  if (use_fpol) {
    const double psihat = psif / psi_sep;
    double fpol;
    if (psihat > 1.0) {
      fpol = splfpol[(size_t)6 * (nrad - 1)];
    } else {
      int k = std::max(0, (int)(psihat / hfpol));
      const double dx = psihat - k * hfpol;
      if (k > nrad - 1) k = nrad - 1;
      const double *s = &splfpol[(size_t)6 * k];
      fpol = s[5];
      for (int j = 4; j >= 0; j--) fpol = fpol * dx + s[j];
    }
    Bp = fpol / rrr;
  } else {
    Bp = btf * rtf / rrr;
  }
}

// vertex fields of a cylindrical grid from this equilibrium (make_tetra_physics :340-343, :418-446)
void EfitField::vertex_fields(const Mesh &m, const gorilla_settings &st, int n2, VertexFields &vf) const
{
  vf.resize((size_t)m.nvert, false, false);
#pragma omp parallel for schedule(static)
  for (int64_t iv = 0; iv < m.nvert; iv++) {
    const double r = m.verts_rphiz[3 * iv], z = m.verts_rphiz[3 * iv + 2];
    double Br, Bp, Bz, psif;
    field(r, z, Br, Bp, Bz, psif);
    const double bmod = std::sqrt(Br * Br + Bp * Bp + Bz * Bz) * m.bmod_multiplier;
    vf.A_x1[iv] = 0.0;
    vf.A_x2[iv] = psif;
    vf.A_x3[iv] = -rtf * btf * std::log(r);
    vf.bmod[iv] = bmod;
    vf.h_x1[iv] = Br / bmod;
    vf.h_x2[iv] = (Bp * r) / bmod;
    vf.h_x3[iv] = Bz / bmod;
    vf.phi_elec[iv] = vf.A_x2[iv] * st.eps_Phi;
  }
  if (st.boole_strong_electric_field) {
    auto psif_at = [&](double r, double z) {
      double Br, Bp, Bz, psif;
      field(r, z, Br, Bp, Bz, psif);
      return psif;
    };
    strong_electric_vertex_fields(m, n2, st.eps_Phi, psif_at, vf);
  }
}

// ---- grid_kind = 1 -----------------------------------------------------------------------------------------------
int build_efit_rect(const gorilla_grid_settings &gs, const gorilla_settings &st, Mesh &m, std::string &err)
{
  if (st.coord_system != 1) {
    err = "grid_kind 1 requires coord_system = 1 (tetra_physics_mod.f90:243-247)";
    return GORILLA_ERR_ARG;
  }
  if (gs.n1 < 1 || gs.n2 < 1 || gs.n3 < 1) { err = "n1,n2,n3 must be positive"; return GORILLA_ERR_ARG; }
  EfitField f;
  f.nwindow_r = gs.nwindow_r; f.nwindow_z = gs.nwindow_z;
  int rc = f.load_efit(gs.g_file_filename, err);
  if (rc) return rc;
  if (gs.convex_wall_filename && gs.convex_wall_filename[0]) {
    rc = f.load_convex_wall(gs.convex_wall_filename, err);
    if (rc) return rc;
  }
  m.grid_kind = 1;
  m.coord_system = 1;
  m.grid_size[0] = gs.n1; m.grid_size[1] = gs.n2; m.grid_size[2] = gs.n3;
  m.n_field_periods = gs.boole_n_field_periods ? 1 : gs.n_field_periods_manual;
  m.sfc_s_min = gs.sfc_s_min;
  m.Rmin = f.rad.front(); m.Rmax = f.rad.back(); m.Zmin = f.zet.front(); m.Zmax = f.zet.back();
  // the reference leaves the axis position unset for grid_kind 1 (only used by the Er_mod estimate); the g-file's
  // magnetic axis is used here
  m.mag_axis_R0 = f.axis_R; m.mag_axis_Z0 = f.axis_Z;
  make_grid_rect(m);
  VertexFields vf;
  f.vertex_fields(m, st, gs.n2, vf);
  apply_vertex_noise(m, st, vf);
  linearise_tetrahedra(m, vf);
  check_tetra_overlaps(m);
  return GORILLA_OK;
}

}  // namespace gbhost
