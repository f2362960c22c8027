// mesh_analytic.cpp -- grid_kind = 5: analytic large-aspect-ratio circular tokamak on a rectangular
// (R, phi, Z) grid.  No input files.
//
// Reference: field_analytic_circ      SRC/field_analytic_circ_mod.f90:41-91
//            make_tetra_grid case(5)  SRC/tetra_grid_mod.f90:138-150
//            make_grid_rect           SRC/tetra_grid_mod.f90:344-680 (+ check_neighbour :684-735)
//            vector_potential_rphiz   SRC/tetra_physics_mod.f90:1076-1109
//            per-vertex block         SRC/tetra_physics_mod.f90:330-446
#include "mesh_common.hpp"
#include <cmath>

namespace gbhost {

namespace {

// Each hexahedron (ir, iz, iphi) is cut into two prisms of three tetrahedra; corner offsets
// (dr, dz, dphi) of the three tetrahedra of the first prism -- the second prism mirrors them.
const int PRISM[3][4][3] = {
    {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}},
    {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}},
    {{0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {0, 1, 1}},
};

// shared face of two tetrahedra given by their vertex lists (check_neighbour)
void check_neighbour(const int32_t k1[4], const int32_t k2[4], int &iface1, int &iface2)
{
  iface1 = iface2 = -1;
  int matches = 0, match1[3], match2[3];
  for (int i = 0; i < 4; i++) {
    bool three = false;
    for (int j = 0; j < 4; j++) {
      if (k1[i] == k2[j]) {
        match1[matches] = i;
        match2[matches] = j;
        matches++;
        if (matches == 3) { three = true; break; }
      }
    }
    if (three) break;
    if ((i + 1) - matches > 1) return;
  }
  if (matches < 3) return;
  bool u1[4] = {true, true, true, true}, u2[4] = {true, true, true, true};
  for (int i = 0; i < 3; i++) { u1[match1[i]] = false; u2[match2[i]] = false; }
  for (int i = 0; i < 4; i++) if (u1[i]) { iface1 = i + 1; break; }
  for (int i = 0; i < 4; i++) if (u2[i]) { iface2 = i + 1; break; }
}

struct RectGrid {
  int nr, nphi, nz;
  Mesh &m;
  int64_t node(int ir, int iz, int iphi) const  // inodes(ir,iz,iphi), 1-based vertex number
  {
    return ((int64_t)iphi * (nr + 1) + ir) * (nz + 1) + iz + 1;
  }
  int64_t tbeg(int ir, int iz, int iphi) const  // itetrbeg, 1-based cell indices
  {
    return (((int64_t)(iphi - 1) * nr + (ir - 1)) * nz + (iz - 1)) * 6;
  }
  int32_t *knots(int64_t t) { return &m.tetra_grid[(size_t)(t - 1) * TG_N + TG_KNOT]; }
  void link(int64_t t1, int64_t t2)
  {
    int f1, f2;
    check_neighbour(knots(t1), knots(t2), f1, f2);
    if (f1 != -1) {
      m.tetra_grid[(size_t)(t1 - 1) * TG_N + TG_NEIGH + f1 - 1] = (int32_t)t2;
      m.tetra_grid[(size_t)(t1 - 1) * TG_N + TG_NFACE + f1 - 1] = f2;
    }
  }
};

}  // namespace

void make_grid_rect(Mesh &m)
{
  const int nr = m.grid_size[0], nphi = m.grid_size[1], nz = m.grid_size[2];
  RectGrid g{nr, nphi, nz, m};
  m.ntetr = (int64_t)nr * nphi * nz * 6;
  m.nvert = (int64_t)(nr + 1) * (nphi + 1) * (nz + 1);
  m.tetra_grid.assign((size_t)m.ntetr * TG_N, 0);
  m.verts_rphiz.assign((size_t)m.nvert * 3, 0.0);
  const double hr = (m.Rmax - m.Rmin) / nr, hphi = 2.0 * PI / nphi, hz = (m.Zmax - m.Zmin) / nz;
  const double R_c = 0.5 * (m.Rmax + m.Rmin), Z_c = 0.5 * (m.Zmax + m.Zmin);
  for (int iphi = 0; iphi <= nphi; iphi++) {
    const double phi = hphi * iphi;
    for (int ir = 0; ir <= nr; ir++)
      for (int iz = 0; iz <= nz; iz++) {
        double r = m.Rmin + hr * ir, z = m.Zmin + hz * iz;
        const double x = r - R_c, y = z - Z_c;
        // nper = 0 in the reference: rotation by cos(0)=1, sin(0)=0 kept for value identity
        r = R_c + x * std::cos(0 * phi) + y * std::sin(0 * phi);
        z = Z_c - x * std::sin(0 * phi) + y * std::cos(0 * phi);
        double *v = &m.verts_rphiz[3 * (g.node(ir, iz, iphi) - 1)];
        v[0] = r; v[1] = phi; v[2] = z;
      }
  }
  for (int iphi = 1; iphi <= nphi; iphi++)
    for (int ir = 1; ir <= nr; ir++)
      for (int iz = 1; iz <= nz; iz++) {
        const int64_t t0 = g.tbeg(ir, iz, iphi);
        for (int it = 0; it < 6; it++) {
          int32_t *row = &m.tetra_grid[(size_t)(t0 + it) * TG_N];
          for (int i = 0; i < 4; i++) {
            const int *o = PRISM[it % 3][i];
            row[TG_KNOT + i] = (it < 3) ? (int32_t)g.node(ir - 1 + o[0], iz - 1 + o[1], iphi - 1 + o[2])
                                        : (int32_t)g.node(ir - o[0], iz - o[1], iphi - o[2]);
            row[TG_NEIGH + i] = -1;
            row[TG_NFACE + i] = -1;
            row[TG_PERPHI + i] = 0;
            row[TG_PERTHETA + i] = 0;
          }
        }
        for (int i = 1; i <= 6; i++)
          for (int j = 1; j <= 6; j++)
            if (j != i) g.link(t0 + i, t0 + j);
      }
  for (int iphi = 1; iphi <= nphi; iphi++)
    for (int ir = 1; ir <= nr; ir++)
      for (int iz = 1; iz <= nz; iz++) {
        const int64_t t0 = g.tbeg(ir, iz, iphi);
        auto cross = [&](int64_t tother, int a0, int a1, int b0, int b1) {
          for (int i = a0; i <= a1; i++)
            for (int j = b0; j <= b1; j++) g.link(t0 + i, tother + j);
        };
        if (ir > 1) cross(g.tbeg(ir - 1, iz, iphi), 1, 3, 4, 6);
        if (ir < nr) cross(g.tbeg(ir + 1, iz, iphi), 4, 6, 1, 3);
        if (iz > 1) cross(g.tbeg(ir, iz - 1, iphi), 1, 3, 4, 6);
        if (iz < nz) cross(g.tbeg(ir, iz + 1, iphi), 4, 6, 1, 3);
        if (iphi > 1) {
          cross(g.tbeg(ir, iz, iphi - 1), 1, 3, 1, 3);
          cross(g.tbeg(ir, iz, iphi - 1), 4, 6, 4, 6);
        }
        if (iphi < nphi) {
          cross(g.tbeg(ir, iz, iphi + 1), 1, 3, 1, 3);
          cross(g.tbeg(ir, iz, iphi + 1), 4, 6, 4, 6);
        }
      }
  // neighbours through the periodic boundary phi = 0 <-> 2 pi
  const int64_t nvertinner = (int64_t)(nr + 1) * (nz + 1) * nphi;
  for (int ir = 1; ir <= nr; ir++)
    for (int iz = 1; iz <= nz; iz++)
      for (int half = 0; half < 2; half++)
        for (int i = 1 + 3 * half; i <= 3 + 3 * half; i++) {
          const int64_t t1 = g.tbeg(ir, iz, 1) + i;
          for (int j = 1 + 3 * half; j <= 3 + 3 * half; j++) {
            const int64_t t2 = g.tbeg(ir, iz, nphi) + j;
            int32_t k2[4];
            for (int q = 0; q < 4; q++) k2[q] = (int32_t)(g.knots(t2)[q] % nvertinner);
            int f1, f2;
            check_neighbour(g.knots(t1), k2, f1, f2);
            if (f1 != -1) {
              int32_t *r1 = &m.tetra_grid[(size_t)(t1 - 1) * TG_N], *r2 = &m.tetra_grid[(size_t)(t2 - 1) * TG_N];
              r1[TG_NEIGH + f1 - 1] = (int32_t)t2; r1[TG_NFACE + f1 - 1] = f2; r1[TG_PERPHI + f1 - 1] = -1;
              r2[TG_NEIGH + f2 - 1] = (int32_t)t1; r2[TG_NFACE + f2 - 1] = f1; r2[TG_PERPHI + f2 - 1] = 1;
            }
          }
        }
}

namespace {

struct AnalyticCirc {
  double R0, a, B0, q0, q1;
  // returns Br, Bp, Bz and psif
  void field(double r, double z, double &Br, double &Bp, double &Bz, double &psif) const
  {
    const double Rshift = r - R0;
    const double rho = std::sqrt(Rshift * Rshift + z * z);
    const double t = rho / a;
    const double q = q0 + q1 * (t * t);
    if (q1 > 0.0) psif = B0 * (a * a / (2.0 * q1)) * std::log(q / q0);
    else psif = B0 * (rho * rho) / (2.0 * q0);
    Bp = B0 * R0 / r;
    Br = -B0 * z / (r * q);
    Bz = B0 * Rshift / (r * q);
  }
};

} // namespace

int build_analytic_circ(const gorilla_grid_settings &gs, const gorilla_settings &st, Mesh &m, std::string &err)
{
  if (st.coord_system != 1) {
    err = "grid_kind 5 requires coord_system = 1 (tetra_physics_mod.f90:243-247)";
    return GORILLA_ERR_ARG;
  }
  if (gs.n1 < 1 || gs.n2 < 1 || gs.n3 < 1) { err = "n1,n2,n3 must be positive"; return GORILLA_ERR_ARG; }
  m.grid_kind = 5;
  m.coord_system = 1;
  m.grid_size[0] = gs.n1; m.grid_size[1] = gs.n2; m.grid_size[2] = gs.n3;
  m.n_field_periods = gs.boole_n_field_periods ? 1 : gs.n_field_periods_manual;
  m.sfc_s_min = gs.sfc_s_min;
  AnalyticCirc f{gs.R0_analytic_circ, gs.a_analytic_circ, gs.B0_analytic_circ, gs.q0_analytic_circ, gs.q1_analytic_circ};
  m.Rmin = f.R0 - f.a; m.Rmax = f.R0 + f.a; m.Zmin = -f.a; m.Zmax = f.a;
  m.mag_axis_R0 = f.R0; m.mag_axis_Z0 = 0.0;  // hard coded in the reference (:297-299)
  make_grid_rect(m);
  VertexFields vf;
  vf.resize((size_t)m.nvert, false, false);
  const double rtf = f.R0, btf = f.B0;  // set_field_analytic_circ
#pragma omp parallel for schedule(static)
  for (int64_t iv = 0; iv < m.nvert; iv++) {
    const double r = m.verts_rphiz[3 * iv], z = m.verts_rphiz[3 * iv + 2];
    double Br, Bp, Bz, psif;
    f.field(r, z, Br, Bp, Bz, psif);
    const double bmod = std::sqrt(Br * Br + Bp * Bp + Bz * Bz) * m.bmod_multiplier;
    vf.A_x1[iv] = 0.0;
    vf.A_x2[iv] = psif;
    vf.A_x3[iv] = -rtf * btf * std::log(r);
    vf.bmod[iv] = bmod;
    vf.h_x1[iv] = Br / bmod;
    vf.h_x2[iv] = (Bp * r) / bmod;  // covariant phi component
    vf.h_x3[iv] = Bz / bmod;
    vf.phi_elec[iv] = vf.A_x2[iv] * st.eps_Phi;
  }
  if (st.boole_strong_electric_field) {
    auto psif_at = [&](double r, double z) {
      double Br, Bp, Bz, psif;
      f.field(r, z, Br, Bp, Bz, psif);
      return psif;
    };
    strong_electric_vertex_fields(m, gs.n2, st.eps_Phi, psif_at, vf);
  }
  apply_vertex_noise(m, st, vf);
  linearise_tetrahedra(m, vf);
  check_tetra_overlaps(m);
  return GORILLA_OK;
}

} // namespace gbhost
