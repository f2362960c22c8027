// mesh_common.cpp -- see mesh_common.hpp
#include "mesh_common.hpp"
#include <cmath>
#include <cstring>

namespace gbhost {

const double PI = 3.141592653589793238462643383;
const double CLIGHT = 2.9979e10, ECHARGE = 4.8032e-10, AMP = 1.6726e-24, AME = 9.1094e-28;

// orbit_timestep_gorilla.f90:204-249
int set_species(Mesh &m, int ispecies, std::string &err)
{
  switch (ispecies) {
    case 1:
      m.particle_charge = -ECHARGE;
      m.particle_mass = AME;
      m.cm_over_e = -CLIGHT * AME / ECHARGE;
      break;
    case 2:
      m.particle_charge = ECHARGE;
      m.particle_mass = 2.0 * AMP;
      m.cm_over_e = 2.0 * CLIGHT * AMP / ECHARGE;
      break;
    case 3:
      m.particle_charge = 2.0 * ECHARGE;
      m.particle_mass = 4.0 * AMP;
      m.cm_over_e = 2.0 * CLIGHT * AMP / ECHARGE;
      break;
    case 4:
      m.particle_charge = 74.0 * ECHARGE;
      m.particle_mass = 184.0 * AMP;
      m.cm_over_e = 184.0 * CLIGHT * AMP / (74.0 * ECHARGE);
      break;
    default:
      err = "invalid ispecies";
      return GORILLA_ERR_ARG;
  }
  return GORILLA_OK;
}

// LAPACK dgesv(3,3,a,3,ipiv,b=I,3) restated: LU with partial pivoting (dgetf2: pivot = first max |a|,
// column scaled by the reciprocal pivot), then forward/back substitution column by column.
// a, b are column-major 3x3; on return b = inverse(a).
static void invert3_lu(double a[9], double b[9])
{
  int ipiv[3];
  for (int j = 0; j < 3; j++) {
    int p = j;
    double amax = std::fabs(a[j + 3 * j]);
    for (int i = j + 1; i < 3; i++)
      if (std::fabs(a[i + 3 * j]) > amax) {
        amax = std::fabs(a[i + 3 * j]);
        p = i;
      }
    ipiv[j] = p;
    if (p != j)
      for (int k = 0; k < 3; k++) std::swap(a[j + 3 * k], a[p + 3 * k]);
    const double rp = 1.0 / a[j + 3 * j];
    for (int i = j + 1; i < 3; i++) a[i + 3 * j] *= rp;
    for (int k = j + 1; k < 3; k++)
      for (int i = j + 1; i < 3; i++) a[i + 3 * k] -= a[i + 3 * j] * a[j + 3 * k];
  }
  for (int i = 0; i < 9; i++) b[i] = 0.0;
  b[0] = b[4] = b[8] = 1.0;
  for (int j = 0; j < 3; j++)  // dlaswp
    if (ipiv[j] != j)
      for (int k = 0; k < 3; k++) std::swap(b[j + 3 * k], b[ipiv[j] + 3 * k]);
  for (int c = 0; c < 3; c++) {
    double *x = b + 3 * c;
    for (int k = 0; k < 3; k++)  // L (unit) forward
      for (int i = k + 1; i < 3; i++) x[i] -= x[k] * a[i + 3 * k];
    for (int k = 2; k >= 0; k--) {  // U backward
      x[k] = x[k] / a[k + 3 * k];
      for (int i = 0; i < k; i++) x[i] -= x[k] * a[i + 3 * k];
    }
  }
}

// differentiate.f90: derivatives of vertex-linear functions
static void differentiate(const double x[4], const double y[4], const double z[4], int n, const double *f /*[n][4]*/,
                          double *fx, double *fy, double *fz)
{
  double a[9], b[9];
  a[0 + 3 * 0] = x[1] - x[0]; a[0 + 3 * 1] = x[2] - x[0]; a[0 + 3 * 2] = x[3] - x[0];
  a[1 + 3 * 0] = y[1] - y[0]; a[1 + 3 * 1] = y[2] - y[0]; a[1 + 3 * 2] = y[3] - y[0];
  a[2 + 3 * 0] = z[1] - z[0]; a[2 + 3 * 1] = z[2] - z[0]; a[2 + 3 * 2] = z[3] - z[0];
  invert3_lu(a, b);
  for (int q = 0; q < n; q++) {
    const double *fq = f + 4 * q;
    const double df[3] = {fq[1] - fq[0], fq[2] - fq[0], fq[3] - fq[0]};
    // df = matmul(transpose(b), df): out(i) = sum_k b(k,i)*df(k)
    fx[q] = ((0.0 + b[0 + 3 * 0] * df[0]) + b[1 + 3 * 0] * df[1]) + b[2 + 3 * 0] * df[2];
    fy[q] = ((0.0 + b[0 + 3 * 1] * df[0]) + b[1 + 3 * 1] * df[1]) + b[2 + 3 * 1] * df[2];
    fz[q] = ((0.0 + b[0 + 3 * 2] * df[0]) + b[1 + 3 * 2] * df[1]) + b[2 + 3 * 2] * df[2];
  }
}

namespace {
struct Xoshiro256ss {  // xoshiro256** (Blackman, Vigna), state from splitmix64
  uint64_t s[4];
  explicit Xoshiro256ss(uint64_t seed)
  {
    for (auto &w : s) {
      seed += 0x9e3779b97f4a7c15ull;
      uint64_t z = seed;
      z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
      z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
      w = z ^ (z >> 31);
    }
  }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  double next()  // uniform in [0, 1), 53 bits
  {
    const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return (double)(r >> 11) * 0x1.0p-53;
  }
};
}  // namespace

void apply_vertex_noise(const Mesh &m, const gorilla_settings &st, VertexFields &vf)
{
  const bool axiA = st.boole_axi_noise_vector_pot != 0, axiP = st.boole_axi_noise_elec_pot != 0,
             nonA = st.boole_non_axi_noise_vector_pot != 0;
  if (!axiA && !axiP && !nonA) return;
  Xoshiro256ss rng(st.noise_seed ? (uint64_t)(uint32_t)st.noise_seed : 0x60121aull);
  const int64_t per_plane = m.nvert / m.grid_size[1];  // nvert / grid_size(2)
  std::vector<double> rnd_axi;
  if (axiA || axiP) {  // :256-261
    rnd_axi.resize((size_t)per_plane);
    for (auto &r : rnd_axi) r = rng.next();
  }
  for (int64_t iv = 0; iv < m.nvert; iv++) {  // in vertex order, as the reference's serial loop draws
    if (axiA) {  // :400-405
      const double r = rnd_axi[(size_t)(iv % per_plane)];
      vf.A_x1[iv] = vf.A_x1[iv] + vf.A_x1[iv] * st.axi_noise_eps_A * r;
      vf.A_x2[iv] = vf.A_x2[iv] + vf.A_x2[iv] * st.axi_noise_eps_A * r;
      vf.A_x3[iv] = vf.A_x3[iv] + vf.A_x3[iv] * st.axi_noise_eps_A * r;
    }
    if (nonA) {  // :407-415
      const double r1 = rng.next(), r2 = rng.next(), r3 = rng.next();
      vf.A_x1[iv] = vf.A_x1[iv] + vf.A_x1[iv] * st.non_axi_noise_eps_A * r1;
      vf.A_x2[iv] = vf.A_x2[iv] + vf.A_x2[iv] * st.non_axi_noise_eps_A * r2;
      vf.A_x3[iv] = vf.A_x3[iv] + vf.A_x3[iv] * st.non_axi_noise_eps_A * r3;
    }
    // :425-439 the electrostatic potential follows the (noisy) vector potential unless the strong-field mode set it
    if (!vf.strong && (axiA || nonA)) vf.phi_elec[iv] = vf.A_x2[iv] * st.eps_Phi;
    if (axiP) vf.phi_elec[iv] = vf.phi_elec[iv] + vf.phi_elec[iv] * st.axi_noise_eps_Phi * rnd_axi[(size_t)(iv % per_plane)];  // :441-444
  }
}

void linearise_tetrahedra(Mesh &m, const VertexFields &vf)
{
  const int64_t ntetr = m.ntetr;
  const int gk = m.grid_kind, cs = m.coord_system;
  const bool strong = vf.strong;  // mutually exclusive with grid_kind 3 (:217-227)
  const int navec = ((gk == 3) ? 11 : 10) + (strong ? 4 : 0);
  m.tetra_physics.assign((size_t)ntetr * TP_N, 0.0);
  const bool skew = m.handover_processing_kind == 2;
  if (skew) m.tetra_skew_coord.assign((size_t)ntetr * 168, 0.0);
  const int64_t last_slice_start = ntetr - ntetr / m.grid_size[1] + 1;
  const double two_pi_nfp = 2.0 * PI / m.n_field_periods;

#pragma omp parallel for schedule(static)
  for (int64_t it = 1; it <= ntetr; it++) {
    double *T = &m.tetra_physics[(size_t)(it - 1) * TP_N];
    const int32_t *G = &m.tetra_grid[(size_t)(it - 1) * TG_N];
    double p1[4], p2[4], p3[4], avec[14][4];
    for (int i = 0; i < 4; i++) {
      const int64_t iv = G[TG_KNOT + i] - 1;
      const double *vr = &m.verts_rphiz[3 * iv];
      if (cs == 1) {
        p1[i] = vr[0]; p2[i] = vr[1]; p3[i] = vr[2];
      } else {
        const double *vs = &m.verts_sthetaphi[3 * iv];
        p1[i] = vs[0]; p2[i] = vs[1]; p3[i] = vs[2];
      }
      // periodic boundary: vertices of the last phi slice stored with phi = 0  (:489-511)
      if (gk == 2 || gk == 4) {
        if (cs == 1) {
          if (it >= last_slice_start && vr[1] == 0.0) p2[i] = 2.0 * PI;
        } else {
          if (it >= last_slice_start && m.verts_sthetaphi[3 * iv + 2] == 0.0) p3[i] = two_pi_nfp;
        }
      } else if (gk == 3) {
        if (it >= last_slice_start && m.verts_sthetaphi[3 * iv + 2] == 0.0) p3[i] = two_pi_nfp;
      }
      avec[0][i] = vf.A_x1[iv]; avec[1][i] = vf.A_x2[iv]; avec[2][i] = vf.A_x3[iv];
      avec[3][i] = vf.h_x1[iv]; avec[4][i] = vf.h_x2[iv]; avec[5][i] = vf.h_x3[iv];
      avec[6][i] = vf.bmod[iv]; avec[7][i] = vf.phi_elec[iv];
      avec[8][i] = vr[0]; avec[9][i] = vr[2];
      if (gk == 3) avec[10][i] = vf.sqg[iv];
      if (strong) {  // (:531-536)
        avec[10][i] = vf.vE_x1[iv]; avec[11][i] = vf.vE_x2[iv]; avec[12][i] = vf.vE_x3[iv]; avec[13][i] = vf.v2E[iv];
      }
    }
    if (cs == 2) {  // theta = 2pi vertices stored as 0  (:534-540)
      for (int j = 0; j < 4; j++) {
        bool any_ge_pi = false;
        for (int k = 0; k < 4; k++)
          if (p2[k] >= PI) any_ge_pi = true;
        if (p2[j] == 0.0 && any_ge_pi) p2[j] = 2.0 * PI;
      }
    }
    T[TP_X1] = p1[0]; T[TP_X1 + 1] = p2[0]; T[TP_X1 + 2] = p3[0];
    {
      const int64_t iv1 = G[TG_KNOT] - 1;
      T[TP_R1] = m.verts_rphiz[3 * iv1];
      T[TP_Z1] = m.verts_rphiz[3 * iv1 + 2];
    }
    for (int i = 0; i < 4; i++) {  // face normals (:568-596)
      double c1[3], c2[3], c3[3];
      int k = 0;
      for (int j = 0; j < 4; j++) {
        if (j == i) continue;
        c1[k] = p1[j]; c2[k] = p2[j]; c3[k] = p3[j];
        k++;
      }
      c1[0] -= c1[2]; c1[1] -= c1[2];
      c2[0] -= c2[2]; c2[1] -= c2[2];
      c3[0] -= c3[2]; c3[1] -= c3[2];
      double *an = &T[TP_ANORM + 3 * i];
      an[0] = c2[0] * c3[1] - c2[1] * c3[0];
      an[1] = c3[0] * c1[1] - c3[1] * c1[0];
      an[2] = c1[0] * c2[1] - c1[1] * c2[0];
      double d = an[0] * (p1[i] - c1[2]) + an[1] * (p2[i] - c2[2]) + an[2] * (p3[i] - c3[2]);
      if (d < 0.0) {
        d = -d;
        an[0] = -an[0]; an[1] = -an[1]; an[2] = -an[2];
      }
      T[TP_DIST_REF_VEC + i] = d;
    }
    T[TP_DIST_REF] = T[TP_DIST_REF_VEC];
    T[TP_BMOD1] = avec[6][0];
    if (cs == 1) {
      T[TP_APHI1] = avec[1][0];
    } else {
      T[TP_ATHETA1] = avec[1][0];
      T[TP_APHI1] = avec[2][0];
    }
    T[TP_H1_1] = avec[3][0]; T[TP_H2_1] = avec[4][0]; T[TP_H3_1] = avec[5][0];
    T[TP_PHI1] = avec[7][0];
    if (gk == 3) T[TP_SQG1] = avec[10][0];
    if (strong) {
      T[TP_VE1_1] = avec[10][0]; T[TP_VE2_1] = avec[11][0]; T[TP_VE3_1] = avec[12][0]; T[TP_V2EMOD_1] = avec[13][0];
    }

    double d1[14], d2[14], d3[14];
    differentiate(p1, p2, p3, navec, &avec[0][0], d1, d2, d3);
    // 0-based quantity index q = Fortran index - 1: A1..A3 = 0..2, h1..h3 = 3..5, B = 6, Phi = 7, R = 8, Z = 9, sqg = 10
    double *curlA = &T[TP_CURLA], *curlh = &T[TP_CURLH];
    curlA[0] = d2[2] - d3[1]; curlA[1] = d3[0] - d1[2]; curlA[2] = d1[1] - d2[0];
    T[TP_GBXCURLA] = d1[6] * curlA[0] + d2[6] * curlA[1] + d3[6] * curlA[2];
    T[TP_GPHIXCURLA] = d1[7] * curlA[0] + d2[7] * curlA[1] + d3[7] * curlA[2];
    T[TP_GB] = d1[6]; T[TP_GB + 1] = d2[6]; T[TP_GB + 2] = d3[6];
    T[TP_GPHI] = d1[7]; T[TP_GPHI + 1] = d2[7]; T[TP_GPHI + 2] = d3[7];
    T[TP_GR] = d1[8]; T[TP_GR + 1] = d2[8]; T[TP_GR + 2] = d3[8];
    T[TP_GZ] = d1[9]; T[TP_GZ + 1] = d2[9]; T[TP_GZ + 2] = d3[9];
    if (gk == 3) { T[TP_GSQG] = d1[10]; T[TP_GSQG + 1] = d2[10]; T[TP_GSQG + 2] = d3[10]; }
    if (cs == 1) {
      T[TP_GAPHI] = d1[1]; T[TP_GAPHI + 1] = d2[1]; T[TP_GAPHI + 2] = d3[1];
    } else {
      T[TP_GATHETA] = d1[1]; T[TP_GATHETA + 1] = d2[1]; T[TP_GATHETA + 2] = d3[1];
      T[TP_GAPHI] = d1[2]; T[TP_GAPHI + 1] = d2[2]; T[TP_GAPHI + 2] = d3[2];
    }
    T[TP_GH1] = d1[3]; T[TP_GH1 + 1] = d2[3]; T[TP_GH1 + 2] = d3[3];
    T[TP_GH2] = d1[4]; T[TP_GH2 + 1] = d2[4]; T[TP_GH2 + 2] = d3[4];
    T[TP_GH3] = d1[5]; T[TP_GH3 + 1] = d2[5]; T[TP_GH3 + 2] = d3[5];
    curlh[0] = d2[5] - d3[4]; curlh[1] = d3[3] - d1[5]; curlh[2] = d1[4] - d2[3];
    const double h1 = avec[3][0], h2 = avec[4][0], h3 = avec[5][0];
    T[TP_GBXH1] = d2[6] * h3 - d3[6] * h2;
    T[TP_GBXH1 + 1] = d3[6] * h1 - d1[6] * h3;
    T[TP_GBXH1 + 2] = d1[6] * h2 - d2[6] * h1;
    T[TP_GPHIXH1] = d2[7] * h3 - d3[7] * h2;
    T[TP_GPHIXH1 + 1] = d3[7] * h1 - d1[7] * h3;
    T[TP_GPHIXH1 + 2] = d1[7] * h2 - d2[7] * h1;
    // alpmat / betmat (:726-795); mat(i,j) stored column-major at [i + 3*j]
    const double *dd[3] = {d1, d2, d3};
    for (int pass = 0; pass < 2; pass++) {
      const int q = (pass == 0) ? 6 : 7;  // B or Phi
      const double *grad = (pass == 0) ? &T[TP_GB] : &T[TP_GPHI];
      double *mat = (pass == 0) ? &T[TP_ALPMAT] : &T[TP_BETMAT];
      for (int j = 0; j < 3; j++) {
        const double *dj = dd[j];
        mat[0 + 3 * j] = 2.0 * curlh[0] * grad[j] + d2[q] * dj[5] - d3[q] * dj[4];
        mat[1 + 3 * j] = 2.0 * curlh[1] * grad[j] + d3[q] * dj[3] - d1[q] * dj[5];
        mat[2 + 3 * j] = 2.0 * curlh[2] * grad[j] + d1[q] * dj[4] - d2[q] * dj[3];
      }
    }
    if (strong) {  // (:708-749, :820-847) quantity indices: vE1..3 = 10..12, v2Emod = 13
      for (int q = 0; q < 3; q++) {
        double *g = &T[(q == 0) ? TP_GVE1 : (q == 1) ? TP_GVE2 : TP_GVE3];
        g[0] = d1[10 + q]; g[1] = d2[10 + q]; g[2] = d3[10 + q];
      }
      T[TP_GV2EMOD] = d1[13]; T[TP_GV2EMOD + 1] = d2[13]; T[TP_GV2EMOD + 2] = d3[13];
      double *cv = &T[TP_CURLVE];
      cv[0] = d2[12] - d3[11]; cv[1] = d3[10] - d1[12]; cv[2] = d1[11] - d2[10];
      T[TP_GV2EMODXH1] = d2[13] * h3 - d3[13] * h2;
      T[TP_GV2EMODXH1 + 1] = d3[13] * h1 - d1[13] * h3;
      T[TP_GV2EMODXH1 + 2] = d1[13] * h2 - d2[13] * h1;
      T[TP_GV2EMODXCURLA] = d1[13] * curlA[0] + d2[13] * curlA[1] + d3[13] * curlA[2];
      T[TP_GBXCURLVE] = d1[6] * cv[0] + d2[6] * cv[1] + d3[6] * cv[2];
      T[TP_GPHIXCURLVE] = d1[7] * cv[0] + d2[7] * cv[1] + d3[7] * cv[2];
      T[TP_GV2EMODXCURLVE] = d1[13] * cv[0] + d2[13] * cv[1] + d3[13] * cv[2];
      const double *grad = &T[TP_GV2EMOD];
      double *mat = &T[TP_GAMMAT];
      for (int j = 0; j < 3; j++) {
        const double *dj = dd[j];
        mat[0 + 3 * j] = 2.0 * curlh[0] * grad[j] + d2[13] * dj[5] - d3[13] * dj[4];
        mat[1 + 3 * j] = 2.0 * curlh[1] * grad[j] + d3[13] * dj[3] - d1[13] * dj[5];
        mat[2 + 3 * j] = 2.0 * curlh[2] * grad[j] + d1[13] * dj[4] - d2[13] * dj[3];
      }
      T[TP_SPGAMMAT] = T[TP_GAMMAT] + T[TP_GAMMAT + 4] + T[TP_GAMMAT + 8];
      for (int f = 0; f < 4; f++) {  // acoef_pre_strong_electric = matmul(curlvE, anorm) (:852-855)
        const double *an = &T[TP_ANORM + 3 * f];
        T[TP_ACOEF_PRE_SE + f] = ((0.0 + cv[0] * an[0]) + cv[1] * an[1]) + cv[2] * an[2];
      }
    }
    T[TP_SPALPMAT] = T[TP_ALPMAT] + T[TP_ALPMAT + 4] + T[TP_ALPMAT + 8];
    T[TP_SPBETMAT] = T[TP_BETMAT] + T[TP_BETMAT + 4] + T[TP_BETMAT + 8];
    for (int f = 0; f < 4; f++) {  // acoef_pre = matmul(curlA, anorm)
      const double *an = &T[TP_ANORM + 3 * f];
      T[TP_ACOEF_PRE + f] = ((0.0 + curlA[0] * an[0]) + curlA[1] * an[1]) + curlA[2] * an[2];
    }
    // dt_dtau_const = 1/4 sum_j sqrt(g)_j |B|_j  (:857-881)
    double dtd = 0.0;
    for (int j = 0; j < 4; j++) {
      double met_det;
      if (cs == 1) met_det = p1[j];
      else if (gk == 3) met_det = avec[10][j];
      else {  // EFIT flux coordinates: metric_determinant() of the LINEARISED quantities (:1276-1279)
        const double dx[3] = {p1[j] - p1[0], p2[j] - p2[0], p3[j] - p3[0]};
        auto lin = [&](double v1, const double *g) { return v1 + (((0.0 + g[0] * dx[0]) + g[1] * dx[1]) + g[2] * dx[2]); };
        const double Rl = lin(T[TP_R1], &T[TP_GR]);
        met_det = (Rl * Rl * m.psitor_max) / (lin(T[TP_H3_1], &T[TP_GH3]) * lin(T[TP_BMOD1], &T[TP_GB]));
      }
      dtd = dtd + met_det * avec[6][j];
    }
    T[TP_DT_DTAU_CONST] = dtd / 4.0;
    // ExB drift scale for the passing-time estimate (:884-920): the mean |v_E| in strong-field mode, else Er_mod
    double er = 0.0;
    if (strong) {
      double acc = 0.0;
      for (int j = 0; j < 4; j++) acc = acc + std::sqrt(avec[13][j]);
      T[TP_VE_MOD_AVG] = acc / 4.0;
    } else if (cs == 1) {
      for (int j = 0; j < 4; j++) {
        const double dr = avec[8][j] - m.mag_axis_R0, dz = avec[9][j] - m.mag_axis_Z0;
        const double r_minor = std::sqrt(dr * dr + dz * dz);
        if (r_minor > 0.0) er = er + (T[TP_GPHI] * dr / r_minor + T[TP_GPHI + 2] * dz / r_minor);
      }
    } else {
      for (int j = 0; j < 4; j++) {
        const int64_t iv = G[TG_KNOT + j] - 1;
        const double dr = avec[8][j] - m.mag_axis_R0, dz = avec[9][j] - m.mag_axis_Z0;
        er = er + T[TP_GPHI] * std::sqrt(dr * dr + dz * dz) / (dr * vf.dR_ds[iv] + dz * vf.dZ_ds[iv]);
      }
    }
    T[TP_ER_MOD] = std::fabs(er / 4.0);
    T[TP_TETRA_DIST_REF] = 2.0 * PI / m.n_field_periods * T[TP_R1] / m.grid_size[1];

    if (skew) {  // matrices for the position exchange via Cartesian coordinates (tetra_physics_mod.f90:946-1011)
      // type tetrahedron_skew_coord: skew_coord_x1x2x3(3,3,4) skew_coord_xyz(3,3,4) inv_skew_coord_x1x2x3(3,3,4)
      // inv_skew_coord_xyz(3,3,4) skew_ref_x1x2x3(3,4) skew_ref_xyz(3,4), column-major
      double *S = &m.tetra_skew_coord[(size_t)(it - 1) * 168];
      double xyz[4][3];
      for (int i = 0; i < 4; i++) {  // verts_xyz (tetra_grid_mod.f90:155-160)
        const double *vr = &m.verts_rphiz[3 * (size_t)(G[TG_KNOT + i] - 1)];
        xyz[i][0] = vr[0] * cos(vr[1]); xyz[i][1] = vr[0] * sin(vr[1]); xyz[i][2] = vr[2];
      }
      auto inv3 = [](const double *A, double *B) {  // dmatinv3 (various_functions_mod.f90:6-40), A(i,j) = A[i + 3*j]
        auto a = [&](int i, int j) { return A[(i - 1) + 3 * (j - 1)]; };
        double detinv = (a(1, 1) * a(2, 2) * a(3, 3) - a(1, 1) * a(2, 3) * a(3, 2) - a(1, 2) * a(2, 1) * a(3, 3) +
                         a(1, 2) * a(2, 3) * a(3, 1) + a(1, 3) * a(2, 1) * a(3, 2) - a(1, 3) * a(2, 2) * a(3, 1));
        if (detinv == 0.0) {
          for (int q = 0; q < 9; q++) B[q] = 0.0;
          return;
        }
        detinv = 1 / detinv;
        auto b = [&](int i, int j) -> double & { return B[(i - 1) + 3 * (j - 1)]; };
        b(1, 1) = +detinv * (a(2, 2) * a(3, 3) - a(2, 3) * a(3, 2));
        b(2, 1) = -detinv * (a(2, 1) * a(3, 3) - a(2, 3) * a(3, 1));
        b(3, 1) = +detinv * (a(2, 1) * a(3, 2) - a(2, 2) * a(3, 1));
        b(1, 2) = -detinv * (a(1, 2) * a(3, 3) - a(1, 3) * a(3, 2));
        b(2, 2) = +detinv * (a(1, 1) * a(3, 3) - a(1, 3) * a(3, 1));
        b(3, 2) = -detinv * (a(1, 1) * a(3, 2) - a(1, 2) * a(3, 1));
        b(1, 3) = +detinv * (a(1, 2) * a(2, 3) - a(1, 3) * a(2, 2));
        b(2, 3) = -detinv * (a(1, 1) * a(2, 3) - a(1, 3) * a(2, 1));
        b(3, 3) = +detinv * (a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1));
      };
      for (int k = 1; k <= 4; k++) {
        int vi[4];
        for (int l = 1; l <= 4; l++) vi[l - 1] = ((k + l - 1) % 4);  // modulo(k+l-1,4)+1, 0-based
        double *cx = S + 9 * (k - 1), *cc = S + 36 + 9 * (k - 1), *icx = S + 72 + 9 * (k - 1), *icc = S + 108 + 9 * (k - 1);
        double *rx = S + 144 + 3 * (k - 1), *rc = S + 156 + 3 * (k - 1);
        const double *pp[3] = {p1, p2, p3};
        for (int i = 0; i < 3; i++) {
          rx[i] = pp[i][vi[0]];
          rc[i] = xyz[vi[0]][i];
          for (int j = 0; j < 3; j++) {
            cx[i + 3 * j] = pp[i][vi[j + 1]] - pp[i][vi[0]];
            cc[i + 3 * j] = xyz[vi[j + 1]][i] - xyz[vi[0]][i];
          }
        }
        inv3(cx, icx);
        inv3(cc, icc);
      }
    }
  }
  // sign_sqg = sign(metric_determinant(1, x1(tetra 1)))  (:1019)
  {
    const double *T = &m.tetra_physics[0];
    double md = (cs == 1) ? T[TP_X1] : (gk == 3) ? T[TP_SQG1] : (T[TP_R1] * T[TP_R1] * m.psitor_max) / (T[TP_H3_1] * T[TP_BMOD1]);
    m.sign_sqg = std::signbit(md) ? -1 : 1;
  }
}

void strong_electric_vertex_fields(const Mesh &m, int n2, double eps_Phi, const std::function<double(double, double)> &psif_at,
                                   VertexFields &vf)
{
  vf.resize_strong((size_t)m.nvert);
  double lim[3][2];
  for (int k = 0; k < 3; k++) { lim[k][0] = INFINITY; lim[k][1] = -INFINITY; }
  for (int64_t iv = 0; iv < m.nvert; iv++)
    for (int k = 0; k < 3; k++) {
      const double v = m.verts_rphiz[3 * iv + k];
      if (v < lim[k][0]) lim[k][0] = v;
      if (v > lim[k][1]) lim[k][1] = v;
    }
  const double average_2D_n = std::sqrt((double)(m.nvert / n2));
  const double npts[3] = {average_2D_n, (double)n2, average_2D_n};
  double dx[3];
  for (int k = 0; k < 3; k++) dx[k] = std::fabs(lim[k][1] - lim[k][0]) / npts[k] * 1.0e-6;
  auto potential = [&](double r, double z) { return psif_at(r, z) * eps_Phi; };
#pragma omp parallel for schedule(static)
  for (int64_t iv = 0; iv < m.nvert; iv++) {
    const double R = m.verts_rphiz[3 * iv], Z = m.verts_rphiz[3 * iv + 2];
    const double E1 = -(potential(R + dx[0], Z) - potential(R + -dx[0], Z)) / (2 * dx[0]);
    const double E2 = -(potential(R, Z) - potential(R, Z)) / (2 * dx[1]);  // axisymmetric potential
    const double E3 = -(potential(R, Z + dx[2]) - potential(R, Z + -dx[2])) / (2 * dx[2]);
    vf.phi_elec[iv] = potential(R, Z);
    const double h1 = vf.h_x1[iv], h2 = vf.h_x2[iv], h3 = vf.h_x3[iv], B = vf.bmod[iv];
    const double v1 = (E2 * h3 - E3 * h2) / (R * B) * CLIGHT;
    const double v2 = (E3 * h1 - E1 * h3) / (B)*R * CLIGHT;
    const double v3 = (E1 * h2 - E2 * h1) / (R * B) * CLIGHT;
    vf.vE_x1[iv] = v1; vf.vE_x2[iv] = v2; vf.vE_x3[iv] = v3;
    vf.v2E[iv] = v1 * v1 + v2 * 1 / (R * R) * v2 + v3 * v3;
  }
}

void check_tetra_overlaps(Mesh &m)
{
  int64_t wrong = 0;
  for (int64_t i = 0; i < m.ntetr; i++) {
    for (int j = 0; j < 4; j++) {
      const int32_t nt = m.tetra_grid[(size_t)i * TG_N + TG_NEIGH + j];
      if (nt == -1) continue;
      const int32_t nf = m.tetra_grid[(size_t)i * TG_N + TG_NFACE + j];
      if (nf < 1) continue;
      const double *a = &m.tetra_physics[(size_t)i * TP_N + TP_ANORM + 3 * j];
      const double *b = &m.tetra_physics[(size_t)(nt - 1) * TP_N + TP_ANORM + 3 * (nf - 1)];
      const double s = ((0.0 + a[0] * b[0]) + a[1] * b[1]) + a[2] * b[2];
      if (s >= 0.0) {
        m.tetra_grid[(size_t)i * TG_N + TG_NFACE + j] = -1;
        wrong++;
      }
    }
  }
  m.n_overlaps = wrong;
}

} // namespace gbhost
