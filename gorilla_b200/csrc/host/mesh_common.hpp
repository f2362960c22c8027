// mesh_common.hpp -- host-side mesh container and the per-tetrahedron linearisation shared by all
// grid kinds.  Runs once (north_star: "mesh build, field splines and equilibrium reading stay on the
// host"); it produces the reference's own AoS layout so that its output is interchangeable with
// arrays a Fortran caller passes from tetra_physics_mod / tetra_grid_mod.
//
// Reference: make_tetra_physics  SRC/tetra_physics_mod.f90:127-1034 (per-tetra block :463-1015),
//            differentiate       SRC/differentiate.f90:7-63 (LAPACK dgesv on a 3x3 system),
//            check_tetra_overlaps SRC/tetra_physics_mod.f90:1291-1336,
//            species constants   SRC/orbit_timestep_gorilla.f90:204-249.
#pragma once
#include <cstdint>
#include <functional>
#include <string>
#include <vector>
#include "../../../include/gorilla_b200.h"

namespace gbhost {

// offsets into type tetrahedron_physics (doubles)
enum {
  TP_X1 = 0, TP_DIST_REF = 3, TP_DIST_REF_VEC = 4, TP_TETRA_DIST_REF = 8, TP_ANORM = 9, TP_CURLA = 21,
  TP_BMOD1 = 24, TP_ATHETA1 = 25, TP_APHI1 = 26, TP_H1_1 = 27, TP_H2_1 = 28, TP_H3_1 = 29, TP_PHI1 = 30,
  TP_R1 = 31, TP_Z1 = 32, TP_VE1_1 = 33, TP_VE2_1 = 34, TP_VE3_1 = 35, TP_V2EMOD_1 = 36, TP_ER_MOD = 37, TP_VE_MOD_AVG = 38, TP_SQG1 = 39, TP_DT_DTAU_CONST = 40, TP_GBXCURLA = 41,
  TP_GPHIXCURLA = 42, TP_SPALPMAT = 47, TP_SPBETMAT = 48, TP_GBXH1 = 50, TP_GPHIXH1 = 53, TP_GB = 59,
  TP_GPHI = 62, TP_GR = 65, TP_GZ = 68, TP_GSQG = 71, TP_GATHETA = 74, TP_GAPHI = 77, TP_GH1 = 80, TP_GH2 = 83,
  TP_GH3 = 86, TP_CURLH = 89, TP_ALPMAT = 107, TP_BETMAT = 116, TP_ACOEF_PRE = 134, TP_N = 142,
  // strong electric field (boole_strong_electric_field)
  TP_GV2EMODXCURLA = 43, TP_GBXCURLVE = 44, TP_GPHIXCURLVE = 45, TP_GV2EMODXCURLVE = 46, TP_SPGAMMAT = 49,
  TP_GV2EMODXH1 = 56, TP_GVE1 = 92, TP_GVE2 = 95, TP_GVE3 = 98, TP_CURLVE = 101, TP_GV2EMOD = 104, TP_GAMMAT = 125,
  TP_ACOEF_PRE_SE = 138
};
enum { TG_KNOT = 0, TG_NEIGH = 4, TG_NFACE = 8, TG_PERPHI = 12, TG_PERTHETA = 16, TG_N = 20 };

struct Mesh {
  int64_t ntetr = 0, nvert = 0;
  std::vector<double> tetra_physics;  // [ntetr][142]
  std::vector<int32_t> tetra_grid;    // [ntetr][20]
  std::vector<double> verts_rphiz;    // [nvert][3]
  std::vector<double> verts_sthetaphi;  // [nvert][3] (flux-coordinate grids)
  std::vector<double> verts_theta_vmec; // [nvert]
  // handover_processing_kind = 2: type tetrahedron_skew_coord (tetra_physics_mod.f90:89-99), [ntetr][168]
  int32_t handover_processing_kind = 1;
  std::vector<double> tetra_skew_coord;
  double cm_over_e = 0, particle_mass = 0, particle_charge = 0;
  int32_t sign_sqg = 1, coord_system = 1, n_field_periods = 1, grid_kind = 0;
  int32_t grid_size[3] = {0, 0, 0};
  double Rmin = 0, Rmax = 0, Zmin = 0, Zmax = 0, sfc_s_min = 0;
  double mag_axis_R0 = 0, mag_axis_Z0 = 0;
  double psitor_max = 0;  // EFIT flux coordinates (grid_kind 2): toroidal flux at the last surface, A_theta = s*psitor_max
  int64_t n_overlaps = 0;
  double bmod_multiplier = 1.0;  // make_tetra_physics' optional argument (tetra_physics_mod.f90:281-286), set by gorilla_mesh_build
};

// values of the field at the vertices, as gathered in make_tetra_physics :330-446
struct VertexFields {
  std::vector<double> A_x1, A_x2, A_x3, h_x1, h_x2, h_x3, bmod, phi_elec;
  std::vector<double> sqg, dR_ds, dZ_ds;  // flux coordinates only (sqg: VMEC only)
  // strong electric field mode (cylindrical grids): covariant ExB drift v_E and its square, per vertex
  bool strong = false;
  std::vector<double> vE_x1, vE_x2, vE_x3, v2E;
  void resize(size_t n, bool flux, bool vmec)
  {
    A_x1.resize(n); A_x2.resize(n); A_x3.resize(n); h_x1.resize(n); h_x2.resize(n); h_x3.resize(n);
    bmod.resize(n); phi_elec.resize(n);
    if (flux) { dR_ds.resize(n); dZ_ds.resize(n); }
    if (vmec) sqg.resize(n);
  }
  void resize_strong(size_t n)
  {
    strong = true;
    vE_x1.resize(n); vE_x2.resize(n); vE_x3.resize(n); v2E.resize(n);
  }
};

extern const double PI;          // constants_mod.f90:5 (full precision)
extern const double CLIGHT, ECHARGE, AMP, AME;

int set_species(Mesh &m, int ispecies, std::string &err);
// per-tetra block of make_tetra_physics; m.tetra_grid, m.verts_* and m.grid_size must be filled
void linearise_tetrahedra(Mesh &m, const VertexFields &vf);
// optional random noise on the vertex potentials (make_tetra_physics :256-261, :400-415, :441-444); call after the vertex
// fields are complete and before linearise_tetrahedra
void apply_vertex_noise(const Mesh &m, const gorilla_settings &st, VertexFields &vf);
void check_tetra_overlaps(Mesh &m);

// make_grid_rect (SRC/tetra_grid_mod.f90:344-680): rectangular (R, phi, Z) grid over [Rmin,Rmax] x [Zmin,Zmax], grid_size set
void make_grid_rect(Mesh &m);
// strong_electric_field_mod.f90: potential = psif*eps_Phi (option 2), E = -grad(Phi) by central differences with step
// 1e-6 * (coordinate extent / points per direction) (:49-85, :215-252), v_E = c ExB/B^2 covariant (:145-165).
// psif_at(R, Z) is the poloidal flux of the equilibrium (axisymmetric); vf.h_*, vf.bmod must be filled.
void strong_electric_vertex_fields(const Mesh &m, int n2, double eps_Phi, const std::function<double(double, double)> &psif_at,
                                   VertexFields &vf);

// grid builders (one translation unit each)
int build_analytic_circ(const gorilla_grid_settings &g, const gorilla_settings &s, Mesh &m, std::string &err);
int build_vmec(const gorilla_grid_settings &g, const gorilla_settings &s, Mesh &m, std::string &err);
int build_efit_rect(const gorilla_grid_settings &g, const gorilla_settings &s, Mesh &m, std::string &err);
int build_efit_flux(const gorilla_grid_settings &g, const gorilla_settings &s, Mesh &m, std::string &err);
int build_soledge3x(const gorilla_grid_settings &g, const gorilla_settings &s, Mesh &m, std::string &err);

} // namespace gbhost
