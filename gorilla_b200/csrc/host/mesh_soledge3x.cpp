// mesh_soledge3x.cpp -- grid_kind = 4: a 2-D triangle mesh of the poloidal plane (SOLEDGE3X-EIRENE knots/triangles
// files) extruded toroidally into prisms of three tetrahedra, with the field of an axisymmetric equilibrium in the
// WEST table format.
//
// Reference: create_points / calc_mesh / calc_triangle_type / connect_* / repair
//                                    SRC/circular_mesh_SOLEDGE3X_EIRENE.f90:11-772
//            make_grid_SOLEDGE3X_EIRENE   SRC/tetra_grid_mod.f90:279-309
//            make_tetra_grid case(4)      SRC/tetra_grid_mod.f90:124-136 ; iaxieq_in = 1 (tetra_grid_settings_mod.f90:90-93)
//
// How a prism is cut (restated from the reference's mask arithmetic): every triangle has a BASE edge -- the one whose
// end points differ least in poloidal flux -- and a FREE vertex opposite to it; `top` says whether the base lies at
// higher flux than the free vertex.  With N = vertices per slice and (f, a, b) = (free, base a, base b):
//   top    : (f, a, b, b+N)   (f, a, a+N, b+N)   (f, f+N, a+N, b+N)
//   bottom : (b, a, f, f+N)   (b, a, a+N, f+N)   (b, b+N, a+N, f+N)
// so the quadrilateral over the base edge is cut along a -- (b+N) in both cases and the two quadrilaterals at the free
// vertex along f -- (x+N) (top) or x -- (f+N) (bottom).  Neighbouring prisms only connect when they cut their shared
// quadrilateral along the same diagonal; the flux ordering makes that true almost everywhere, and `repair` re-types
// the few triangles where it is not.
//
// Connectivity is found through an edge -> triangles map instead of the reference's all-pairs search (same result:
// tetrahedra of different prisms can only share a face across a shared triangle edge).
#include "mesh_efit.hpp"
#include <algorithm>
#include <array>
#include <cmath>
#include <fstream>
#include <map>
#include <sstream>

namespace gbhost {

namespace {

struct PlaneMesh {
  int n_tri = 0, N = 0;                       // triangles, vertices per slice
  std::vector<std::array<int, 3>> tri;        // 1-based vertex numbers
  std::vector<int> top, free_pos;             // triangle_type(:,1) (0 = top facing, 1 = bottom), (:,2) in 1..3
  std::vector<std::array<int, 4>> verts;      // [3*n_tri] first-slice tetrahedra (1-based vertex numbers, +N = next slice)
  std::vector<std::array<int, 4>> neigh, nface;
  std::vector<std::vector<int>> adj;          // edge neighbours of every triangle, ascending

  void set_prism_verts(int t)
  {
    const int fp = free_pos[t];  // 1..3
    const int f = tri[t][fp - 1], a = tri[t][fp % 3], b = tri[t][(fp + 1) % 3];
    std::array<int, 4> *v = &verts[(size_t)3 * t];
    if (top[t] == 0) {
      v[0] = {f, a, b, b + N};
      v[1] = {f, a, a + N, b + N};
      v[2] = {f, f + N, a + N, b + N};
    } else {
      v[0] = {b, a, f, f + N};
      v[1] = {b, a, a + N, f + N};
      v[2] = {b, b + N, a + N, f + N};
    }
  }

  // connect every pair of tetrahedra of prisms p1, p2 (0-based) that share exactly three vertices
  bool connect_prisms(int p1, int p2)
  {
    bool match = false;
    for (int o1 = 0; o1 < 3; o1++) {
      const int t1 = 3 * p1 + o1;
      for (int o2 = 0; o2 < 3; o2++) {
        const int t2 = 3 * p2 + o2;
        if (t1 == t2) continue;
        bool s1[4], s2[4];
        int cnt = 0;
        for (int i = 0; i < 4; i++) {
          s1[i] = s2[i] = false;
          for (int k = 0; k < 4; k++) {
            if (verts[t2][k] == verts[t1][i]) s1[i] = true;
            if (verts[t2][i] == verts[t1][k]) s2[i] = true;
          }
          if (s1[i]) cnt++;
        }
        if (cnt != 3) continue;
        int f1 = 0, f2 = 0;
        while (s1[f1]) f1++;
        while (s2[f2]) f2++;
        neigh[t1][f1] = t2 + 1; neigh[t2][f2] = t1 + 1;
        nface[t1][f1] = f2 + 1; nface[t2][f2] = f1 + 1;
        match = true;
      }
    }
    return match;
  }
};

bool read_table(const char *path, std::vector<double> &vals, int &rows, int &cols)
{
  std::ifstream f(path);
  if (!f) return false;
  if (!(f >> rows >> cols) || rows < 1 || cols < 1) return false;
  vals.resize((size_t)rows * cols);
  for (auto &v : vals)
    if (!(f >> v)) return false;
  return true;
}

}  // namespace

int build_soledge3x(const gorilla_grid_settings &gs, const gorilla_settings &st, Mesh &m, std::string &err)
{
  if (st.coord_system != 1) { err = "grid_kind 4 requires coord_system = 1"; return GORILLA_ERR_ARG; }
  const int n_slices = gs.n2;
  if (n_slices < 3) { err = "grid_kind 4: n2 (toroidal slices) must be >= 3"; return GORILLA_ERR_ARG; }
  EfitField fld;
  fld.nwindow_r = gs.nwindow_r; fld.nwindow_z = gs.nwindow_z;
  int rc = fld.load_west(gs.g_file_filename, err);
  if (rc) return rc;
  if (gs.convex_wall_filename && gs.convex_wall_filename[0]) {
    rc = fld.load_convex_wall(gs.convex_wall_filename, err);
    if (rc) return rc;
  }
  std::vector<double> knots, tris;
  int nk, kc, nt, tc;
  if (!gs.knots_SOLEDGE3X_EIRENE_filename || !read_table(gs.knots_SOLEDGE3X_EIRENE_filename, knots, nk, kc) || kc < 2) {
    err = "cannot read SOLEDGE3X-EIRENE knots file";
    return GORILLA_ERR_IO;
  }
  if (!gs.triangles_SOLEDGE3X_EIRENE_filename || !read_table(gs.triangles_SOLEDGE3X_EIRENE_filename, tris, nt, tc) || tc < 3) {
    err = "cannot read SOLEDGE3X-EIRENE triangles file";
    return GORILLA_ERR_IO;
  }
  PlaneMesh P;
  P.n_tri = nt;
  P.N = nk;
  P.tri.resize(nt);
  for (int t = 0; t < nt; t++)
    for (int k = 0; k < 3; k++) {
      const int v = (int)tris[(size_t)t * tc + k];
      if (v < 1 || v > nk) { err = "triangles file: vertex index out of range"; return GORILLA_ERR_IO; }
      P.tri[t][k] = v;
    }

  m.grid_kind = 4;
  m.coord_system = 1;
  m.grid_size[0] = gs.n1; m.grid_size[1] = gs.n2; m.grid_size[2] = gs.n3;
  m.n_field_periods = gs.boole_n_field_periods ? 1 : gs.n_field_periods_manual;
  m.sfc_s_min = gs.sfc_s_min;
  m.mag_axis_R0 = 240.0; m.mag_axis_Z0 = 0.0;  // hard coded in the reference (tetra_physics_mod.f90:296-298)

  // ---- vertices: the plane copied to n_slices toroidal angles (create_points, extrude_points)
  const double pi_trunc = 3.14159265358979;  // this module's own pi
  m.nvert = (int64_t)nk * n_slices;
  m.verts_rphiz.assign((size_t)m.nvert * 3, 0.0);
  for (int s = 0; s < n_slices; s++) {
    const double phi = (s == 0) ? 0.0 : (2.0 * pi_trunc / m.n_field_periods * s) / n_slices;
    for (int i = 0; i < nk; i++) {
      double *v = &m.verts_rphiz[3 * ((size_t)s * nk + i)];
      v[0] = knots[(size_t)i * kc]; v[1] = phi; v[2] = knots[(size_t)i * kc + 1];
    }
  }
  m.Rmin = m.Zmin = INFINITY; m.Rmax = m.Zmax = -INFINITY;
  for (int i = 0; i < nk; i++) {
    m.Rmin = std::min(m.Rmin, knots[(size_t)i * kc]); m.Rmax = std::max(m.Rmax, knots[(size_t)i * kc]);
    m.Zmin = std::min(m.Zmin, knots[(size_t)i * kc + 1]); m.Zmax = std::max(m.Zmax, knots[(size_t)i * kc + 1]);
  }

  // ---- triangle types from the poloidal flux at the knots (calc_triangle_type)
  std::vector<double> psi(nk);
  for (int i = 0; i < nk; i++) {
    double Br, Bp, Bz;
    fld.field(knots[(size_t)i * kc], knots[(size_t)i * kc + 1], Br, Bp, Bz, psi[i]);
  }
  P.top.resize(nt); P.free_pos.resize(nt);
  for (int t = 0; t < nt; t++) {
    const double a1 = psi[P.tri[t][0] - 1], a2 = psi[P.tri[t][1] - 1], a3 = psi[P.tri[t][2] - 1];
    const double d[3] = {std::fabs(a1 - a2), std::fabs(a2 - a3), std::fabs(a3 - a1)};
    int imin = 0;
    if (d[1] < d[imin]) imin = 1;
    if (d[2] < d[imin]) imin = 2;
    if (imin == 0) { P.top[t] = (a1 > a3) ? 0 : 1; P.free_pos[t] = 3; }
    else if (imin == 1) { P.top[t] = (a2 > a1) ? 0 : 1; P.free_pos[t] = 1; }
    else { P.top[t] = (a3 > a2) ? 0 : 1; P.free_pos[t] = 2; }
  }

  // ---- first slice: tetrahedra, edge adjacency, connections (connect_plane)
  const int tps = 3 * nt;  // tetrahedra per slice
  P.verts.resize(tps);
  P.neigh.assign(tps, {0, 0, 0, 0});
  P.nface.assign(tps, {-1, -1, -1, -1});
  for (int t = 0; t < nt; t++) P.set_prism_verts(t);
  {
    std::map<std::pair<int, int>, std::vector<int>> edges;
    for (int t = 0; t < nt; t++)
      for (int k = 0; k < 3; k++) {
        int u = P.tri[t][k], v = P.tri[t][(k + 1) % 3];
        if (u > v) std::swap(u, v);
        edges[{u, v}].push_back(t);
      }
    P.adj.assign(nt, {});
    for (auto &e : edges)
      for (size_t i = 0; i < e.second.size(); i++)
        for (size_t j = 0; j < e.second.size(); j++)
          if (i != j) P.adj[e.second[i]].push_back(e.second[j]);
    for (auto &a : P.adj) {
      std::sort(a.begin(), a.end());
      a.erase(std::unique(a.begin(), a.end()), a.end());
    }
  }
  std::vector<int> count_connected(nt, 0);
  for (int i = 0; i < nt; i++) {
    const int t0 = 3 * i;
    P.neigh[t0][3] = t0 + 1 + 2 - tps;  P.nface[t0][3] = 1;         // face 4 of tetra 1 <-> previous slice
    P.neigh[t0 + 2][0] = t0 + 1 + tps;  P.nface[t0 + 2][0] = 4;     // face 1 of tetra 3 <-> next slice
    P.connect_prisms(i, i);
    for (int j : P.adj[i]) {
      if (j < i) continue;
      if (P.connect_prisms(i, j)) { count_connected[i]++; count_connected[j]++; }
    }
  }

  // ---- repair: re-type triangles whose prisms did not connect to all their edge neighbours (:661-772)
  int n_error = 0, n_repair = 0;
  for (int k = 0; k < nt; k++) {
    const int n_nb = std::min<int>(3, (int)P.adj[k].size());
    if (!(count_connected[k] < 3 && n_nb != count_connected[k])) continue;
    n_error++;
    int nb[3] = {-1, -1, -1};
    for (int q = 0; q < n_nb; q++) nb[q] = P.adj[k][q];
    // state of the prism and its neighbours, restored if no re-typing fits
    std::array<int, 4> old_n[12], old_f[12];
    for (int o = 0; o < 3; o++) { old_n[o] = P.neigh[3 * k + o]; old_f[o] = P.nface[3 * k + o]; }
    for (int q = 0; q < 3; q++)
      if (nb[q] != -1)
        for (int o = 0; o < 3; o++) { old_n[3 * (q + 1) + o] = P.neigh[3 * nb[q] + o]; old_f[3 * (q + 1) + o] = P.nface[3 * nb[q] + o]; }
    bool fixed = false;
    for (int l = 0; l < 2 && !fixed; l++) {
      P.top[k] = (P.top[k] + 1) % 2;
      for (int r = 0; r < 3 && !fixed; r++) {
        P.free_pos[k] = P.free_pos[k] % 3 + 1;
        P.set_prism_verts(k);
        for (int o = 0; o < 3; o++) P.nface[3 * k + o] = {-1, -1, -1, -1};
        P.nface[3 * k][3] = 1;
        P.nface[3 * k + 2][0] = 4;
        P.connect_prisms(k, k);
        bool all = true;
        for (int q = 0; q < 3; q++) {
          if (nb[q] == -1) continue;
          if (!P.connect_prisms(k, nb[q])) { all = false; break; }
        }
        if (all) fixed = true;
      }
    }
    if (fixed) { n_repair++; continue; }
    for (int o = 0; o < 3; o++) { P.neigh[3 * k + o] = old_n[o]; P.nface[3 * k + o] = old_f[o]; }
    for (int q = 0; q < 3; q++)
      if (nb[q] != -1)
        for (int o = 0; o < 3; o++) { P.neigh[3 * nb[q] + o] = old_n[3 * (q + 1) + o]; P.nface[3 * nb[q] + o] = old_f[3 * (q + 1) + o]; }
  }
  (void)n_error; (void)n_repair;

  // ---- all slices: shift vertex and neighbour numbers, wrap around the torus, periodic-boundary flags
  m.ntetr = (int64_t)tps * n_slices;
  m.tetra_grid.assign((size_t)m.ntetr * TG_N, 0);
  const int64_t n_verts = m.nvert, n_tetras = m.ntetr;
  auto wrap = [](int64_t idx, int64_t period) { return ((idx - 1) % period + period) % period + 1; };
  for (int s = 0; s < n_slices; s++)
    for (int t = 0; t < tps; t++) {
      int32_t *G = &m.tetra_grid[((size_t)s * tps + t) * TG_N];
      for (int i = 0; i < 4; i++) {
        int64_t v = (int64_t)P.verts[t][i] + (int64_t)s * nk;
        if (s == n_slices - 1) v = wrap(v, n_verts);
        G[TG_KNOT + i] = (int32_t)v;
        const int nf = P.nface[t][i];
        G[TG_NFACE + i] = nf;
        G[TG_NEIGH + i] = (nf == -1) ? -1 : (int32_t)wrap((int64_t)P.neigh[t][i] + (int64_t)s * tps, n_tetras);
        G[TG_PERPHI + i] = 0;
      }
    }
  for (int t = 0; t < tps; t += 3) m.tetra_grid[(size_t)t * TG_N + TG_PERPHI + 3] = -1;
  for (int64_t t = n_tetras - tps + 2; t < n_tetras; t += 3) m.tetra_grid[(size_t)t * TG_N + TG_PERPHI + 0] = 1;
  // consistency: a connection must be mutual, over the same face, with opposite periodic-boundary flags
  for (int64_t i = 0; i < n_tetras; i++)
    for (int f = 0; f < 4; f++) {
      const int32_t *G = &m.tetra_grid[(size_t)i * TG_N];
      const int nb = G[TG_NEIGH + f], nf = G[TG_NFACE + f];
      if (nb == -1 && nf == -1) continue;
      const int32_t *H = &m.tetra_grid[(size_t)(nb - 1) * TG_N];
      if (H[TG_NEIGH + nf - 1] != i + 1 || H[TG_PERPHI + nf - 1] != -G[TG_PERPHI + f]) {
        err = "SOLEDGE3X mesh: neighbour consistency check failed (mesh is broken)";
        return GORILLA_ERR_DOMAIN;
      }
    }

  VertexFields vf;
  fld.vertex_fields(m, st, gs.n2, vf);
  apply_vertex_noise(m, st, vf);
  linearise_tetrahedra(m, vf);
  check_tetra_overlaps(m);
  return GORILLA_OK;
}

}  // namespace gbhost
