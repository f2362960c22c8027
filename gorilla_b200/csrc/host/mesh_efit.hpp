// mesh_efit.hpp -- axisymmetric equilibrium field (EFIT g-file / WEST table) shared by the grid_kind 1 and 4 builders
#pragma once
#include "mesh_common.hpp"

namespace gbhost {

void spl_five_reg(int n, double h, const double *a, double *b, double *c, double *d, double *e, double *f);

struct Spline2D {
  int nx = 0, ny = 0;
  double hx = 0, hy = 0;
  std::vector<double> spl;  // [ny][nx][6(y power l)][6(x power)] : at(i,j)[xpow + 6*l]
  double *at(int i, int j) { return &spl[(size_t)36 * (i + (size_t)nx * j)]; }
  const double *at(int i, int j) const { return &spl[(size_t)36 * (i + (size_t)nx * j)]; }
  void build(int nx, int ny, double hx, double hy, const std::vector<double> &fxy);
  void eval(double x0, double y0, double xb, double yb, double &u, double &ux, double &uy, double &uxx, double &uxy,
            double &uyy) const;
};

struct EfitField {
  int nrad = 0, nzet = 0;
  std::vector<double> rad, zet;  // cm
  double hrad = 0, hzet = 0, btf = 0, rtf = 0, psi_sep = 0, hfpol = 0;
  double axis_R = 0, axis_Z = 0;
  bool use_fpol = false;
  std::vector<double> splfpol;   // [nrad][6]
  Spline2D psi_spl;
  // convex wall for stretch_coords
  bool have_wall = false;
  double wall_R0 = 0, wall_htht = 0;
  std::vector<double> rho_wall, tht_wall;

  int nwindow_r = 0, nwindow_z = 0;  // field_divB0.inp: psi filter windows, set before load_*
  void filter_psi(std::vector<double> &psi) const;  // window_filter over R, then over Z (bdivfree.f90:1144-1164)
  int load_efit(const char *path, std::string &err);
  int load_west(const char *path, std::string &err);
  int load_convex_wall(const char *path, std::string &err);
  void stretch_coords(double r, double z, double &rm, double &zm) const;
  void field_eq(double r, double z, double &Br, double &Bp, double &Bz, double &psif) const;  // no stretching
  void field(double r, double z, double &Br, double &Bp, double &Bz, double &psif) const;     // stretch_coords + field_eq
  void vertex_fields(const Mesh &m, const gorilla_settings &st, int n2, VertexFields &vf) const;
};

}  // namespace gbhost
