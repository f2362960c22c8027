// gb_repack.hpp -- host: reference AoS (tetrahedron_physics / tetrahedron_grid) -> device sub-record SoA
// (layout documented in gb_mesh.cuh).  Pure host code, used by gorilla_b200_init.
#pragma once
#include <cstring>
#include <vector>
#include "../../include/gorilla_b200.h"
#include "gb_mesh.cuh"

namespace gb {

inline bool repack_mesh(const gorilla_mesh_desc *md, std::vector<double> &geom, std::vector<double> &bpart,
                        std::vector<double> &phi, std::vector<double> &cold, bool &has_phi, std::vector<double> *se = nullptr)
{
  const int64_t nt = md->ntetr;
  geom.assign((size_t)nt * GEOM_ND, 0.0);
  bpart.assign((size_t)nt * BPART_ND, 0.0);
  phi.assign((size_t)nt * PHI_ND, 0.0);
  cold.assign((size_t)nt * COLD_ND, 0.0);
  has_phi = false;
  if (se) se->assign((size_t)nt * SE_ND, 0.0);
  // offsets into type tetrahedron_physics (doubles), tetra_physics_mod.f90:9-83
  enum { TP_X1 = 0, TP_DIST_REF = 3, TP_TETRA_DIST_REF = 8, TP_ANORM = 9, TP_CURLA = 21, TP_BMOD1 = 24, TP_APHI1 = 26,
         TP_H2_1 = 28, TP_H3_1 = 29, TP_PHI1 = 30, TP_R1 = 31, TP_ER_MOD = 37, TP_DT_DTAU_CONST = 40, TP_GBXCURLA = 41,
         TP_GPHIXCURLA = 42, TP_SPALPMAT = 47, TP_SPBETMAT = 48, TP_GBXH1 = 50, TP_GPHIXH1 = 53, TP_GB = 59,
         TP_GPHI = 62, TP_GAPHI = 77, TP_GH2 = 83, TP_GH3 = 86, TP_CURLH = 89, TP_ALPMAT = 107, TP_BETMAT = 116,
         TP_VE2_1 = 34, TP_V2EMOD_1 = 36, TP_VE_MOD_AVG = 38, TP_GV2EMODXCURLA = 43, TP_GBXCURLVE = 44, TP_GPHIXCURLVE = 45,
         TP_GV2EMODXCURLVE = 46, TP_SPGAMMAT = 49, TP_GV2EMODXH1 = 56, TP_GVE2 = 95, TP_CURLVE = 101, TP_GV2EMOD = 104,
         TP_GAMMAT = 125 };
  for (int64_t t = 0; t < nt; t++) {
    const double *r = md->tetra_physics + t * GORILLA_TETRA_PHYSICS_NDOUBLES;
    const int32_t *g = md->tetra_grid + t * GORILLA_TETRA_GRID_NINTS;
    double *G = &geom[(size_t)t * GEOM_ND], *B = &bpart[(size_t)t * BPART_ND], *P = &phi[(size_t)t * PHI_ND],
           *C = &cold[(size_t)t * COLD_ND];
    for (int i = 0; i < 3; i++) G[i] = r[TP_X1 + i];
    G[3] = r[TP_DIST_REF];
    for (int i = 0; i < 12; i++) G[4 + i] = r[TP_ANORM + i];
    B[B_BMOD1] = r[TP_BMOD1];
    for (int i = 0; i < 3; i++) {
      B[B_GB + i] = r[TP_GB + i];
      B[B_CURLA + i] = r[TP_CURLA + i];
      B[B_CURLH + i] = r[TP_CURLH + i];
      B[B_GBXH1 + i] = r[TP_GBXH1 + i];
    }
    B[B_GBXCURLA] = r[TP_GBXCURLA];
    for (int i = 0; i < 9; i++) B[B_ALP + i] = r[TP_ALPMAT + i];
    B[B_SPALP] = r[TP_SPALPMAT];
    B[B_DTDTAU] = r[TP_DT_DTAU_CONST];
    int32_t topo[6] = {g[4], g[5], g[6], g[7], 0, 0};
    uint32_t flags = 0;
    for (int f = 0; f < 4; f++) {
      int nf = g[8 + f], pp = g[12 + f], pt = (md->coord_system == 2) ? g[16 + f] : 0;
      if (nf < -1 || nf > 4 || pp < -1 || pp > 1 || pt < -1 || pt > 1) return false;
      flags |= ((uint32_t)(nf + 1) | ((uint32_t)(pp + 1) << 3) | ((uint32_t)(pt + 1) << 5)) << (7 * f);
    }
    topo[4] = (int32_t)flags;
    memcpy(&B[B_TOPO], topo, sizeof(topo));
    P[P_PHI1] = r[TP_PHI1];
    for (int i = 0; i < 3; i++) {
      P[P_GPHI + i] = r[TP_GPHI + i];
      P[P_GPHIXH1 + i] = r[TP_GPHIXH1 + i];
    }
    P[P_GPHIXCURLA] = r[TP_GPHIXCURLA];
    for (int i = 0; i < 9; i++) P[P_BET + i] = r[TP_BETMAT + i];
    P[P_SPBET] = r[TP_SPBETMAT];
    for (int i = 0; i < 18; i++)
      if (P[i] != 0.0) has_phi = true; // NaN counts as "present"
    C[C_TETRA_DIST_REF] = r[TP_TETRA_DIST_REF];
    C[C_R1] = r[TP_R1];
    C[C_ER_MOD] = r[TP_ER_MOD];
    if (md->coord_system == 1) {
      C[C_HPHI1] = r[TP_H2_1];
      for (int i = 0; i < 3; i++) C[C_GHPHI + i] = r[TP_GH2 + i];
    } else {
      C[C_HPHI1] = r[TP_H3_1];
      for (int i = 0; i < 3; i++) C[C_GHPHI + i] = r[TP_GH3 + i];
    }
    C[C_APHI1] = r[TP_APHI1];
    for (int i = 0; i < 3; i++) C[C_GAPHI + i] = r[TP_GAPHI + i];
    if (se) {
      double *S = &(*se)[(size_t)t * SE_ND];
      S[S_V2EMOD1] = r[TP_V2EMOD_1];
      for (int i = 0; i < 3; i++) {
        S[S_GV2EMOD + i] = r[TP_GV2EMOD + i];
        S[S_GV2EMODXH1 + i] = r[TP_GV2EMODXH1 + i];
        S[S_CURLVE + i] = r[TP_CURLVE + i];
        S[S_GVE2 + i] = r[TP_GVE2 + i];
      }
      S[S_GBXCURLVE] = r[TP_GBXCURLVE];
      S[S_GPHIXCURLVE] = r[TP_GPHIXCURLVE];
      S[S_GV2EMODXCURLVE] = r[TP_GV2EMODXCURLVE];
      S[S_GV2EMODXCURLA] = r[TP_GV2EMODXCURLA];
      for (int i = 0; i < 9; i++) S[S_GAMMAT + i] = r[TP_GAMMAT + i];
      S[S_SPGAMMAT] = r[TP_SPGAMMAT];
      S[S_VE_MOD_AVG] = r[TP_VE_MOD_AVG];
      S[S_VE2_1] = r[TP_VE2_1];
    }
  }
  return true;
}

// type hamiltonian_time_type (tetra_physics_mod.f90:105-114), formed from the record exactly as make_tetra_physics does
// (:926-944): h1_in_curlA = sum(h_1 * curlA), h1_in_curlh = sum(h_1 * curlh), vec_mismatch_der = matmul(mat_gh, curlA),
// vec_parcurr_der = matmul(mat_gh, curlh) with mat_gh(:,j) = gh_j.
inline void repack_hamiltonian_time(const gorilla_mesh_desc *md, std::vector<double> &ham)
{
  enum { TP_CURLA = 21, TP_H1_1 = 27, TP_GH1 = 80, TP_GH2 = 83, TP_GH3 = 86, TP_CURLH = 89 };
  const int64_t nt = md->ntetr;
  ham.assign((size_t)nt * HAM_ND, 0.0);
  for (int64_t t = 0; t < nt; t++) {
    const double *r = md->tetra_physics + t * GORILLA_TETRA_PHYSICS_NDOUBLES;
    double *H = &ham[(size_t)t * HAM_ND];
    const double *h1 = r + TP_H1_1, *cA = r + TP_CURLA, *ch = r + TP_CURLH;
    H[0] = (h1[0] * cA[0] + h1[1] * cA[1]) + h1[2] * cA[2];
    H[1] = (h1[0] * ch[0] + h1[1] * ch[1]) + h1[2] * ch[2];
    for (int i = 0; i < 3; i++) {
      H[2 + i] = ((0.0 + r[TP_GH1 + i] * cA[0]) + r[TP_GH2 + i] * cA[1]) + r[TP_GH3 + i] * cA[2];
      H[5 + i] = ((0.0 + r[TP_GH1 + i] * ch[0]) + r[TP_GH2 + i] * ch[1]) + r[TP_GH3 + i] * ch[2];
    }
  }
}

// make_precomp_poly4 (SRC/tetra_physics_poly_precomp_mod.f90:160-476): the tetra_physics_poly4 records, [ntetr][544] in the
// reference's own member order (offsets P4_* of gb_mesh.cuh), formed from the tetra_physics records with the reference's
// operations: 4x4 matmul / matmul(n_vec, M) / sum() accumulate in ascending index order starting from 0, the sums of products
// of alpha and beta chains are added left to right as written.
inline void make_precomp_poly4(const gorilla_mesh_desc *md, std::vector<double> &poly4)
{
  enum { TP_ANORM = 9, TP_CURLA = 21, TP_GBXCURLA = 41, TP_GPHIXCURLA = 42, TP_SPALPMAT = 47, TP_SPBETMAT = 48, TP_GBXH1 = 50,
         TP_GPHIXH1 = 53, TP_CURLH = 89, TP_ALPMAT = 107, TP_BETMAT = 116 };
  const double CLIGHT = 2.9979e10;
  const int64_t nt = md->ntetr;
  const double cm = md->cm_over_e;
  poly4.assign((size_t)nt * P4_ND, 0.0);
  struct M4 { double a[4][4]; };
  auto mul = [](const M4 &A, const M4 &B) {
    M4 C;
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) {
        double s = 0.0;
        for (int k = 0; k < 4; k++) s = s + A.a[i][k] * B.a[k][j];
        C.a[i][j] = s;
      }
    return C;
  };
  auto add = [](const M4 &A, const M4 &B) {
    M4 C;
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) C.a[i][j] = A.a[i][j] + B.a[i][j];
    return C;
  };
#pragma omp parallel for schedule(static)
  for (int64_t t = 0; t < nt; t++) {
    const double *r = md->tetra_physics + t * GORILLA_TETRA_PHYSICS_NDOUBLES;
    double *o = &poly4[(size_t)t * P4_ND];
    M4 alp{}, bet{};
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        alp.a[i][j] = cm * r[TP_ALPMAT + i + 3 * j];
        bet.a[i][j] = -(CLIGHT * r[TP_BETMAT + i + 3 * j]);
      }
    alp.a[3][3] = cm * r[TP_SPALPMAT];
    for (int i = 0; i < 3; i++) bet.a[i][3] = r[TP_CURLA + i];
    bet.a[3][3] = -(CLIGHT * r[TP_SPBETMAT]);
    const M4 aa = mul(alp, alp), ab = mul(alp, bet), bb = mul(bet, bet), ba = mul(bet, alp);
    const M4 aaa = mul(alp, aa), aab = mul(alp, ab), aba = mul(alp, ba), abb = mul(alp, bb);
    const M4 baa = mul(bet, aa), bab = mul(bet, ab), bba = mul(bet, ba), bbb = mul(bet, bb);
    const M4 aaaa = mul(alp, aaa), aaab = mul(alp, aab), aaba = mul(alp, aba), aabb = mul(alp, abb);
    const M4 abaa = mul(alp, baa), abab = mul(alp, bab), abba = mul(alp, bba), abbb = mul(alp, bbb);
    const M4 baaa = mul(bet, aaa), baab = mul(bet, aab), baba = mul(bet, aba), babb = mul(bet, abb);
    const M4 bbaa = mul(bet, baa), bbab = mul(bet, bab), bbba = mul(bet, bba), bbbb = mul(bet, bbb);
    const M4 mats[14] = {
        bet, alp,                                                                  // amat1_0, amat1_1
        bb, add(ba, ab), aa,                                                       // amat2_0..2
        bbb, add(add(bba, bab), abb), add(add(baa, aba), aab), aaa,                // amat3_0..3
        bbbb, add(add(add(bbba, bbab), babb), abbb),                               // amat4_0, amat4_1
        add(add(add(add(add(bbaa, baba), baab), abba), abab), aabb),               // amat4_2
        add(add(add(abaa, aaba), aaab), baaa), aaaa};                              // amat4_3, amat4_4
    for (int k = 0; k < 14; k++) {
      for (int j = 0; j < 4; j++)
        for (int i = 0; i < 4; i++) o[P4_AMAT + 16 * k + i + 4 * j] = mats[k].a[i][j];
      for (int n = 0; n < 4; n++) {   // anorm_in_amat(:,n) = matmul(n_vec, amat), n_vec = (anorm(:,n), 0)
        const double nv[4] = {r[TP_ANORM + 3 * n], r[TP_ANORM + 3 * n + 1], r[TP_ANORM + 3 * n + 2], 0.0};
        for (int j = 0; j < 4; j++) {
          double s = 0.0;
          for (int i = 0; i < 4; i++) s = s + nv[i] * mats[k].a[i][j];
          o[P4_AN_AMAT + 16 * k + j + 4 * n] = s;
        }
      }
    }
    for (int i = 0; i < 3; i++) {
      o[P4_B0 + i] = -CLIGHT * r[TP_GPHIXH1 + i];
      o[P4_B1 + i] = cm * r[TP_CURLH + i];
      o[P4_B2 + i] = cm * r[TP_GBXH1 + i];
      o[P4_B3 + i] = -2.0 * CLIGHT * r[TP_CURLH + i];
    }
    o[P4_B0 + 3] = -CLIGHT / cm * r[TP_GPHIXCURLA];
    o[P4_B1 + 3] = 0.0;
    o[P4_B2 + 3] = r[TP_GBXCURLA];
    o[P4_B3 + 3] = 0.0;
    for (int q = 0; q < 2; q++)      // amat1_q in b_k
      for (int k = 0; k < 4; k++)
        for (int i = 0; i < 4; i++) {
          double s = 0.0;
          for (int j = 0; j < 4; j++) s = s + mats[q].a[i][j] * o[P4_B0 + 4 * k + j];
          o[(q == 0 ? P4_A10_B0 : P4_A11_B0) + 4 * k + i] = s;
        }
    for (int k = 0; k < 4; k++)
      for (int n = 0; n < 4; n++) {
        const double *an = r + TP_ANORM + 3 * n, *bk = o + P4_B0 + 4 * k;
        o[P4_AN_B0 + 4 * k + n] = ((0.0 + an[0] * bk[0]) + an[1] * bk[1]) + an[2] * bk[2];
        for (int q = 0; q < 2; q++) {
          const double *col = o + P4_AN_AMAT + 16 * q + 4 * n;
          double s = 0.0;
          for (int i = 0; i < 4; i++) s = s + col[i] * bk[i];
          o[(q == 0 ? P4_AN_A10_B0 : P4_AN_A11_B0) + 4 * k + n] = s;
        }
      }
  }
}

// type tetrahedron_skew_coord (168 doubles) -> per (tetrahedron, face) a "leave" and an "enter" block (gb_mesh.cuh)
inline void repack_skew(const gorilla_mesh_desc *md, std::vector<double> &skew)
{
  const int64_t nt = md->ntetr;
  skew.assign((size_t)nt * SKEW_ND, 0.0);
  for (int64_t t = 0; t < nt; t++) {
    const double *S = md->tetra_skew_coord + t * GORILLA_TETRA_SKEW_NDOUBLES;
    for (int k = 0; k < 4; k++) {
      double *L = &skew[(size_t)t * SKEW_ND + 48 * k], *E = L + 24;
      for (int i = 0; i < 3; i++) {
        L[i] = S[144 + 3 * k + i];
        L[21 + i] = S[156 + 3 * k + i];
        E[i] = S[156 + 3 * k + i];
        E[21 + i] = S[144 + 3 * k + i];
      }
      for (int q = 0; q < 9; q++) {
        L[3 + q] = S[72 + 9 * k + q];
        L[12 + q] = S[36 + 9 * k + q];
        E[3 + q] = S[108 + 9 * k + q];
        E[12 + q] = S[9 * k + q];
      }
    }
  }
}

// ---- find_tetra bins --------------------------------------------------------------------------------------------
// The slice-wise grids repeat the same 2-D cell pattern in every phi slice.  The tetrahedra of slice 0 are binned by
// the bounding box of their four vertices in the two non-toroidal coordinates; the vertices are reconstructed from the
// face planes of the record (vertex 1 = x1; vertex j lies on face 1, n1.z = -dist_ref, and on the two faces through x1
// other than face j), so the bins can be built from the arrays a Fortran caller passes.  Boxes are inflated by 1e-6 of
// their size: isinside() accepts points up to 1e-10 (relative) outside a face.  Items of a bin are ascending, so walking
// them visits the tetrahedra in the order of the reference's full scan.
struct FindBins {
  std::vector<int32_t> start, items;
  int32_t nu = 0, nv = 0, c0 = 0, c1 = 2;
  double u0 = 0, v0 = 0, du_inv = 0, dv_inv = 0;
};

inline bool build_find_bins(const gorilla_mesh_desc *md, FindBins &fb)
{
  if (!(md->grid_kind == 2 || md->grid_kind == 3 || md->grid_kind == 4) || md->grid_size[1] < 1) return false;
  const int64_t tps = md->ntetr / md->grid_size[1];
  if (tps < 1 || tps * md->grid_size[1] != md->ntetr) return false;
  fb.c0 = 0;
  fb.c1 = (md->coord_system == 2) ? 1 : 2;
  enum { TP_X1 = 0, TP_DIST_REF = 3, TP_ANORM = 9 };
  std::vector<double> lo((size_t)tps * 2), hi((size_t)tps * 2);
  double glo[2] = {1e300, 1e300}, ghi[2] = {-1e300, -1e300};
  auto solve3 = [](const double *a, const double *b, const double *c, double ra, double rb, double rc, double *z) {
    const double det = a[0] * (b[1] * c[2] - b[2] * c[1]) - a[1] * (b[0] * c[2] - b[2] * c[0]) + a[2] * (b[0] * c[1] - b[1] * c[0]);
    if (!(det != 0.0) || !(det == det)) return false;
    z[0] = (ra * (b[1] * c[2] - b[2] * c[1]) - a[1] * (rb * c[2] - b[2] * rc) + a[2] * (rb * c[1] - b[1] * rc)) / det;
    z[1] = (a[0] * (rb * c[2] - b[2] * rc) - ra * (b[0] * c[2] - b[2] * c[0]) + a[2] * (b[0] * rc - rb * c[0])) / det;
    z[2] = (a[0] * (b[1] * rc - rb * c[1]) - a[1] * (b[0] * rc - rb * c[0]) + ra * (b[0] * c[1] - b[1] * c[0])) / det;
    return true;
  };
  for (int64_t t = 0; t < tps; t++) {
    const double *r = md->tetra_physics + t * GORILLA_TETRA_PHYSICS_NDOUBLES;
    const double *x1 = r + TP_X1, *n = r + TP_ANORM;
    double l[2] = {x1[fb.c0], x1[fb.c1]}, h[2] = {x1[fb.c0], x1[fb.c1]};
    for (int j = 1; j < 4; j++) {  // vertex j+1: on face 1 and on the two faces through x1 that are not face j+1
      int f[2], k = 0;
      for (int q = 1; q < 4; q++)
        if (q != j) f[k++] = q;
      double z[3];
      if (!solve3(n, n + 3 * f[0], n + 3 * f[1], -r[TP_DIST_REF], 0.0, 0.0, z)) return false;
      const double p[2] = {x1[fb.c0] + z[fb.c0], x1[fb.c1] + z[fb.c1]};
      for (int d = 0; d < 2; d++) { l[d] = p[d] < l[d] ? p[d] : l[d]; h[d] = p[d] > h[d] ? p[d] : h[d]; }
    }
    for (int d = 0; d < 2; d++) {
      const double m = 1e-6 * (h[d] - l[d]) + 1e-300;
      lo[2 * t + d] = l[d] - m; hi[2 * t + d] = h[d] + m;
      if (!(lo[2 * t + d] == lo[2 * t + d]) || !(hi[2 * t + d] == hi[2 * t + d])) return false;
      glo[d] = lo[2 * t + d] < glo[d] ? lo[2 * t + d] : glo[d];
      ghi[d] = hi[2 * t + d] > ghi[d] ? hi[2 * t + d] : ghi[d];
    }
  }
  int nb = 1;
  while ((int64_t)nb * nb < tps / 3) nb++;
  fb.nu = fb.nv = nb < 1 ? 1 : nb;
  fb.u0 = glo[0]; fb.v0 = glo[1];
  fb.du_inv = fb.nu / (ghi[0] - glo[0]);
  fb.dv_inv = fb.nv / (ghi[1] - glo[1]);
  if (!(fb.du_inv > 0.0) || !(fb.dv_inv > 0.0)) return false;
  auto cell = [&](double v, double v0, double inv, int nn) {
    int c = (int)((v - v0) * inv);
    return c < 0 ? 0 : (c >= nn ? nn - 1 : c);
  };
  fb.start.assign((size_t)fb.nu * fb.nv + 1, 0);
  for (int pass = 0; pass < 2; pass++) {
    std::vector<int32_t> fill;
    if (pass == 1) {
      for (size_t i = 1; i < fb.start.size(); i++) fb.start[i] += fb.start[i - 1];
      fb.items.assign((size_t)fb.start.back(), 0);
      fill.assign(fb.start.begin(), fb.start.end() - 1);
    }
    for (int64_t t = 0; t < tps; t++) {  // ascending t: items of every bin come out ascending
      const int iu0 = cell(lo[2 * t], fb.u0, fb.du_inv, fb.nu), iu1 = cell(hi[2 * t], fb.u0, fb.du_inv, fb.nu);
      const int iv0 = cell(lo[2 * t + 1], fb.v0, fb.dv_inv, fb.nv), iv1 = cell(hi[2 * t + 1], fb.v0, fb.dv_inv, fb.nv);
      for (int iv = iv0; iv <= iv1; iv++)
        for (int iu = iu0; iu <= iu1; iu++) {
          const size_t b = (size_t)iv * fb.nu + iu;
          if (pass == 0) fb.start[b + 1]++;
          else fb.items[(size_t)fill[b]++] = (int32_t)t;
        }
    }
  }
  return true;
}

} // namespace gb
