// mesh_vmec.cpp -- grid_kind = 3: field-aligned tetrahedral grid in symmetry-flux coordinates
// (s, vartheta, varphi) for a 3-D VMEC equilibrium (NetCDF "wout" file).  Host, runs once.
//
// What the reference does (file:line) and what is done here:
//   read wout, unit conversion, lambda half->full mesh      SRC/vmecinm_m.f90:36-86, new_vmec_allocation_stuff.f90:8-29
//       -> own NetCDF classic (CDF-1/CDF-2) reader below; same conversions.
//   field evaluation                                         SRC/spline_vmec_data.f90:7-276, splint_vmec_data.f90:7-207
//       The reference synthesises R, Z, lambda on a (rho, theta, phi) lattice and builds a 3-D tensor-product
//       quintic spline.  Here the Fourier series is summed EXACTLY in both angles at every requested point and
//       only the radial profiles of the harmonics are interpolated (6-point Lagrange on f_mn(s)/rho^m, the
//       same rho^m regularisation as s_to_rho_healaxis :424-497).  The two evaluations agree to interpolation
//       error; nothing downstream depends on the mesh provenance (the hot path takes the records as input).
//   A_phi(s) = -torflux * int iota ds                         spline_vmec_data.f90:56-76
//   metric / B components                                     splint_vmec_data.f90:161-207 (vmec_field)
//   mesh points, Newton for theta_vmec                        SRC/points_2d.f90:171-252, circular_mesh.f90:14-210
//   prism/tetrahedron topology                                SRC/circular_mesh.f90:212-512 (calc_mesh, all rings with
//       n3 vertices => prisms alternate up/down); neighbours are found by matching shared faces.
//   vertex fields                                             SRC/tetra_physics_mod.f90:1167-1210
#include "mesh_common.hpp"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <unordered_map>

namespace gbhost {

namespace {

const double PI_MESH = 3.14159265358979;  // truncated pi of the reference's mesh/VMEC modules

// ------------------------------------------------------------------------------------------------ NetCDF-3
struct NcVar {
  std::vector<int64_t> shape;
  int type = 0;
  int64_t begin = 0, vsize = 0;
};
struct NcFile {
  std::vector<unsigned char> buf;
  std::map<std::string, NcVar> vars;
  size_t pos = 0;
  bool is64 = false;
  uint32_t u32() { uint32_t v = ((uint32_t)buf[pos] << 24) | (buf[pos + 1] << 16) | (buf[pos + 2] << 8) | buf[pos + 3]; pos += 4; return v; }
  uint64_t u64() { uint64_t hi = u32(); uint64_t lo = u32(); return (hi << 32) | lo; }
  std::string name() { uint32_t n = u32(); std::string s((const char *)&buf[pos], n); pos += (n + 3) & ~3u; return s; }
  void skip_atts()
  {
    uint32_t tag = u32(), n = u32();
    if (tag == 0 && n == 0) return;
    for (uint32_t i = 0; i < n; i++) {
      name();
      uint32_t type = u32(), cnt = u32();
      static const int sz[7] = {0, 1, 1, 2, 4, 4, 8};
      pos += ((size_t)cnt * sz[type] + 3) & ~(size_t)3;
    }
  }
  bool open(const std::string &path, std::string &err)
  {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) { err = "cannot open " + path; return false; }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize((size_t)n);
    if (fread(buf.data(), 1, (size_t)n, f) != (size_t)n) { fclose(f); err = "short read " + path; return false; }
    fclose(f);
    if (n < 32 || memcmp(buf.data(), "CDF", 3) != 0 || (buf[3] != 1 && buf[3] != 2)) {
      err = path + ": not a NetCDF classic (CDF-1/CDF-2) file";
      return false;
    }
    is64 = buf[3] == 2;
    pos = 4;
    u32();  // numrecs
    std::vector<int64_t> dims;
    uint32_t tag = u32(), nd = u32();
    if (tag != 0 || nd != 0)
      for (uint32_t i = 0; i < nd; i++) { name(); dims.push_back(u32()); }
    skip_atts();
    tag = u32();
    uint32_t nv = u32();
    if (tag != 0 || nv != 0)
      for (uint32_t i = 0; i < nv; i++) {
        std::string nm = name();
        NcVar v;
        uint32_t rank = u32();
        for (uint32_t k = 0; k < rank; k++) v.shape.push_back(dims[u32()]);
        skip_atts();
        v.type = (int)u32();
        v.vsize = u32();
        v.begin = is64 ? (int64_t)u64() : (int64_t)u32();
        vars[nm] = v;
      }
    return true;
  }
  bool get(const std::string &nm, std::vector<double> &out, std::vector<int64_t> *shape, std::string &err)
  {
    auto it = vars.find(nm);
    if (it == vars.end()) { err = "variable '" + nm + "' missing in NetCDF file"; return false; }
    const NcVar &v = it->second;
    int64_t n = 1;
    for (int64_t s : v.shape) n *= s;
    out.resize((size_t)n);
    const unsigned char *p = &buf[(size_t)v.begin];
    for (int64_t i = 0; i < n; i++) {
      if (v.type == 6) {  // NC_DOUBLE, big endian
        uint64_t b = 0;
        for (int k = 0; k < 8; k++) b = (b << 8) | p[8 * i + k];
        double d;
        memcpy(&d, &b, 8);
        out[(size_t)i] = d;
      } else if (v.type == 4) {  // NC_INT
        uint32_t b = ((uint32_t)p[4 * i] << 24) | (p[4 * i + 1] << 16) | (p[4 * i + 2] << 8) | p[4 * i + 3];
        out[(size_t)i] = (double)(int32_t)b;
      } else if (v.type == 5) {  // NC_FLOAT
        uint32_t b = ((uint32_t)p[4 * i] << 24) | (p[4 * i + 1] << 16) | (p[4 * i + 2] << 8) | p[4 * i + 3];
        float fl;
        memcpy(&fl, &b, 4);
        out[(size_t)i] = fl;
      } else {
        err = "variable '" + nm + "': unsupported NetCDF type";
        return false;
      }
    }
    if (shape) *shape = v.shape;
    return true;
  }
};

// ------------------------------------------------------------------------------------------------ VMEC data
struct Vmec {
  int ns = 0, nm = 0, nfp = 1;
  double torflux = 0, hs = 0;
  std::vector<int> m, n;                 // mode numbers (n includes nfp)
  std::vector<double> rmn, zmn, lmn;      // [mode][ns], cm / rad, divided by rho^m at nodes with s>0
  std::vector<double> iota, aphi_node;   // [ns]

  // 6-point Lagrange weights (value and derivative) for an equidistant grid, nodes i0..i0+5
  void weights(double s, int mm, int &i0, double w[6], double dw[6]) const
  {
    i0 = (int)std::floor(s / hs) - 2;
    const int lo = (mm > 0) ? 1 : 0;  // node 0 (rho = 0) carries no information for m > 0
    i0 = std::max(lo, std::min(ns - 6, i0));
    const double t = s / hs - i0;  // position in units of hs relative to node i0
    for (int j = 0; j < 6; j++) {
      double num = 1.0, den = 1.0, dsum = 0.0;
      for (int k = 0; k < 6; k++)
        if (k != j) { num *= (t - k); den *= (double)(j - k); }
      for (int l = 0; l < 6; l++) {
        if (l == j) continue;
        double pr = 1.0;
        for (int k = 0; k < 6; k++)
          if (k != j && k != l) pr *= (t - k);
        dsum += pr;
      }
      w[j] = num / den;
      dw[j] = dsum / den / hs;
    }
  }

  struct Point {
    double R, Z, lam, dR_ds, dR_dt, dR_dp, dZ_ds, dZ_dt, dZ_dp, dl_ds, dl_dt, dl_dp;
    double A_phi, A_theta, dA_phi_ds, dA_theta_ds, aiota;
  };

  // splint_vmec_data equivalent: everything at (s, theta_vmec, varphi)
  void eval(double s, double theta, double phi, Point &P, bool lambda_only = false) const
  {
    memset(&P, 0, sizeof(P));
    const double rho = std::sqrt(s);
    int i0[2];
    double w[2][6], dw[2][6];
    weights(s, 0, i0[0], w[0], dw[0]);
    weights(s, 1, i0[1], w[1], dw[1]);
    double rp[66];  // rho^(k-2), k = 0..65
    rp[2] = 1.0; rp[1] = 1.0 / rho; rp[0] = rp[1] / rho;
    for (int k = 3; k < 66; k++) rp[k] = rp[k - 1] * rho;
    for (int k = 0; k < nm; k++) {
      const int mm = m[k], sel = mm > 0 ? 1 : 0;
      const double ang = mm * theta - n[k] * phi;
      const double c = std::cos(ang), sn = std::sin(ang);
      const double rm = rp[mm + 2];                                // rho^m
      const double drm = mm > 0 ? 0.5 * mm * rp[mm] : 0.0;         // d(rho^m)/ds = m/2 rho^(m-2)
      double gl = 0, dgl = 0, gr = 0, dgr = 0, gz = 0, dgz = 0;
      const double *pl = &lmn[(size_t)k * ns + i0[sel]], *pr = &rmn[(size_t)k * ns + i0[sel]],
                   *pz = &zmn[(size_t)k * ns + i0[sel]];
      for (int j = 0; j < 6; j++) {
        gl += w[sel][j] * pl[j];
        dgl += dw[sel][j] * pl[j];
        if (!lambda_only) {
          gr += w[sel][j] * pr[j];
          dgr += dw[sel][j] * pr[j];
          gz += w[sel][j] * pz[j];
          dgz += dw[sel][j] * pz[j];
        }
      }
      const double fl = gl * rm, dfl = dgl * rm + gl * drm;
      P.lam += fl * sn;
      P.dl_ds += dfl * sn;
      P.dl_dt += fl * mm * c;
      P.dl_dp += -fl * n[k] * c;
      if (!lambda_only) {
        const double fr = gr * rm, dfr = dgr * rm + gr * drm, fz = gz * rm, dfz = dgz * rm + gz * drm;
        P.R += fr * c;
        P.dR_ds += dfr * c;
        P.dR_dt += -fr * mm * sn;
        P.dR_dp += fr * n[k] * sn;
        P.Z += fz * sn;
        P.dZ_ds += dfz * sn;
        P.dZ_dt += fz * mm * c;
        P.dZ_dp += -fz * n[k] * c;
      }
    }
    // vector potential: A_theta = torflux*s, dA_phi/ds = -torflux*iota(s)
    double wi[6], dwi[6];
    int ii;
    weights(s, 0, ii, wi, dwi);
    double io = 0.0;
    for (int j = 0; j < 6; j++) io += wi[j] * iota[ii + j];
    P.aiota = io;
    P.A_theta = torflux * s;
    P.dA_theta_ds = torflux;
    P.dA_phi_ds = -torflux * io;
    // A_phi(s) = A_phi(nearest node below) - torflux * int_{s_node}^{s} iota  (Gauss-Legendre on the interpolant)
    int inode = std::max(0, std::min(ns - 1, (int)std::floor(s / hs)));
    const double a = inode * hs, b = s;
    static const double gx[3] = {-0.7745966692414834, 0.0, 0.7745966692414834}, gw[3] = {5.0 / 9, 8.0 / 9, 5.0 / 9};
    double integ = 0.0;
    for (int q = 0; q < 3; q++) {
      const double sq = 0.5 * (a + b) + 0.5 * (b - a) * gx[q];
      double wq[6], dq[6];
      int iq;
      weights(sq, 0, iq, wq, dq);
      double v = 0.0;
      for (int j = 0; j < 6; j++) v += wq[j] * iota[iq + j];
      integ += gw[q] * v;
    }
    integ *= 0.5 * (b - a);
    P.A_phi = aphi_node[inode] - torflux * integ;
  }
};

bool load_vmec(const std::string &path, Vmec &V, std::string &err)
{
  NcFile nc;
  if (!nc.open(path, err)) return false;
  std::vector<double> lmns, rmnc, zmns, xm, xn, iotaf, phi, nfp;
  std::vector<int64_t> shp;
  if (!nc.get("lmns", lmns, &shp, err) || !nc.get("rmnc", rmnc, nullptr, err) || !nc.get("zmns", zmns, nullptr, err) ||
      !nc.get("xm", xm, nullptr, err) || !nc.get("xn", xn, nullptr, err) || !nc.get("iotaf", iotaf, nullptr, err) ||
      !nc.get("phi", phi, nullptr, err) || !nc.get("nfp", nfp, nullptr, err))
    return false;
  if (shp.size() != 2) { err = "lmns must be 2-D (radius, mn_mode)"; return false; }
  const int ns = (int)shp[0], nm = (int)shp[1];  // Fortran lens(1)=nstrm=mn_mode, lens(2)=nsurfm=radius
  if (ns < 8) { err = "VMEC file has too few flux surfaces"; return false; }
  V.ns = ns; V.nm = nm; V.nfp = (int)nfp[0];
  V.hs = 1.0 / (ns - 1);
  const double fac_b = 1e4, fac_r = 1e2;
  // vmecin: phi/(2 pi), flux = phi(edge)*fac_b*fac_r^2  (truncated pi of the reference, vmecinm_m.f90:42)
  V.torflux = phi[ns - 1] / (2 * PI_MESH) * fac_b * fac_r * fac_r;
  V.m.resize(nm); V.n.resize(nm);
  for (int k = 0; k < nm; k++) {
    V.m[k] = (int)std::lround(xm[k]); V.n[k] = (int)std::lround(xn[k]);
    if (V.m[k] < 0 || V.m[k] > 63) { err = "poloidal mode number out of range"; return false; }
  }
  V.rmn.assign((size_t)nm * ns, 0.0); V.zmn.assign((size_t)nm * ns, 0.0); V.lmn.assign((size_t)nm * ns, 0.0);
  V.iota = iotaf;
  for (int k = 0; k < nm; k++) {
    for (int i = 0; i < ns; i++) {
      // lambda: half mesh -> full mesh (vmecinm_m.f90:70-77); file layout is [radius][mode]
      double al;
      if (i == 0) al = 0.0;
      else if (i < ns - 1) al = 0.5 * (lmns[(size_t)(i + 1) * nm + k] + lmns[(size_t)i * nm + k]);
      else al = lmns[(size_t)i * nm + k] + 0.5 * (lmns[(size_t)i * nm + k] - lmns[(size_t)(i - 1) * nm + k]);
      const double rho = std::sqrt(V.hs * i);
      const double div = (V.m[k] > 0 && i > 0) ? std::pow(rho, V.m[k]) : 1.0;
      V.rmn[(size_t)k * ns + i] = rmnc[(size_t)i * nm + k] * fac_r / div;
      V.zmn[(size_t)k * ns + i] = zmns[(size_t)i * nm + k] * fac_r / div;
      V.lmn[(size_t)k * ns + i] = al / div;
    }
  }
  // A_phi at the nodes: cumulative Gauss-Legendre integral of the iota interpolant
  V.aphi_node.assign(ns, 0.0);
  static const double gx[3] = {-0.7745966692414834, 0.0, 0.7745966692414834}, gw[3] = {5.0 / 9, 8.0 / 9, 5.0 / 9};
  for (int i = 1; i < ns; i++) {
    const double a = (i - 1) * V.hs, b = i * V.hs;
    double integ = 0.0;
    for (int q = 0; q < 3; q++) {
      const double sq = 0.5 * (a + b) + 0.5 * (b - a) * gx[q];
      double w[6], dw[6];
      int i0;
      V.weights(sq, 0, i0, w, dw);
      double v = 0.0;
      for (int j = 0; j < 6; j++) v += w[j] * V.iota[i0 + j];
      integ += gw[q] * v;
    }
    V.aphi_node[i] = V.aphi_node[i - 1] - V.torflux * 0.5 * (b - a) * integ;
  }
  return true;
}

// theta_sym_flux2theta_vmec (circular_mesh.f90:184-210): Newton on vartheta = theta + lambda(s, theta, varphi)
double theta_vmec_of(const Vmec &V, double s, double vartheta, double phi)
{
  double th = vartheta;
  Vmec::Point P;
  for (int it = 0; it < 100; it++) {
    V.eval(s, th, phi, P, true);
    const double d = (vartheta - th - P.lam) / (1.0 + P.dl_dt);
    th += d;
    if (std::fabs(d) < 1e-14) break;
  }
  return th;
}

// vmec_field (splint_vmec_data.f90:161-207) + vector_potential_sthetaphi_vmec (tetra_physics_mod.f90:1167-1210)
void vertex_field(const Vmec &V, double s, double theta_vmec, double phi, double bmod_multiplier,
                  double out[10] /*A1 A2 A3 h1 h2 h3 B sqg dRds dZds*/)
{
  Vmec::Point P;
  V.eval(s, theta_vmec, phi, P);
  double gV[3][3];
  gV[0][0] = P.dR_ds * P.dR_ds + P.dZ_ds * P.dZ_ds;
  gV[0][1] = gV[1][0] = P.dR_ds * P.dR_dt + P.dZ_ds * P.dZ_dt;
  gV[0][2] = gV[2][0] = P.dR_ds * P.dR_dp + P.dZ_ds * P.dZ_dp;
  gV[1][1] = P.dR_dt * P.dR_dt + P.dZ_dt * P.dZ_dt;
  gV[1][2] = gV[2][1] = P.dR_dt * P.dR_dp + P.dZ_dt * P.dZ_dp;
  gV[2][2] = P.R * P.R + P.dR_dp * P.dR_dp + P.dZ_dp * P.dZ_dp;
  const double sqgV = P.R * (P.dR_dt * P.dZ_ds - P.dR_ds * P.dZ_dt);
  const double cjac = 1.0 / (1.0 + P.dl_dt);
  const double sqg = sqgV * cjac;
  const double Bt = -P.dA_phi_ds / sqg, Bp = P.dA_theta_ds / sqg;  // contravariant vartheta, varphi
  double c[3][3] = {{1, 0, 0}, {-P.dl_ds * cjac, cjac, -P.dl_dp * cjac}, {0, 0, 1}};
  double t[3][3], g[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      t[i][j] = 0;
      for (int k = 0; k < 3; k++) t[i][j] += gV[i][k] * c[k][j];
    }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      g[i][j] = 0;
      for (int k = 0; k < 3; k++) g[i][j] += c[k][i] * t[k][j];
    }
  const double Bcov_r = g[0][1] * Bt + g[0][2] * Bp;
  const double Bcov_t = g[1][1] * Bt + g[1][2] * Bp;
  const double Bcov_p = g[2][1] * Bt + g[2][2] * Bp;
  const double bmod = std::sqrt(Bt * Bcov_t + Bp * Bcov_p) * bmod_multiplier;
  out[0] = 0.0; out[1] = P.A_theta; out[2] = P.A_phi;
  out[3] = Bcov_r / bmod; out[4] = Bcov_t / bmod; out[5] = Bcov_p / bmod;
  out[6] = bmod; out[7] = sqg; out[8] = P.dR_ds; out[9] = P.dZ_ds;
}

struct FaceKey {
  int32_t a, b, c;
  bool operator==(const FaceKey &o) const { return a == o.a && b == o.b && c == o.c; }
};
struct FaceHash {
  size_t operator()(const FaceKey &k) const
  {
    uint64_t h = (uint64_t)(uint32_t)k.a * 0x9E3779B97F4A7C15ull;
    h ^= ((uint64_t)(uint32_t)k.b + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
    h ^= ((uint64_t)(uint32_t)k.c + 0x165667B1ull) * 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    return (size_t)h;
  }
};

}  // namespace

// Topology of the field-aligned grid (all n1 rings with n3 vertices, inner ring repeated at s_min).
void make_field_aligned_topology(Mesh &m, int n1, int n2, int n3)
{
  const int64_t verts_per_slice = (int64_t)(n1 + 1) * n3;
  const int64_t tetras_per_slice = (int64_t)n1 * 2 * n3 * 3;
  m.nvert = verts_per_slice * n2;
  m.ntetr = tetras_per_slice * n2;
  m.tetra_grid.assign((size_t)m.ntetr * TG_N, 0);
  // corner codes: bit0 = +1 in theta, bit1 = outer ring, bit2 = next slice
  static const int CONF[2][3][4] = {{{0, 2, 3, 7}, {0, 2, 6, 7}, {0, 4, 6, 7}},   // prism with its top edge on the outer ring
                                    {{0, 1, 2, 6}, {0, 1, 5, 6}, {0, 4, 5, 6}}};  // mirrored prism
  for (int slice = 0; slice < n2; slice++) {
    int64_t tetra = (int64_t)slice * tetras_per_slice;  // 0-based running tetra index
    for (int ring = 1; ring <= n1; ring++) {
      const int64_t base = (int64_t)(ring - 1) * n3;  // 0-based first vertex of the lower ring in the slice
      for (int seg = 1; seg <= 2 * n3; seg++) {
        const int orient = (seg - 1) & 1, upper_off = seg / 2, lower_off = (seg - 1) / 2;
        for (int t = 0; t < 3; t++, tetra++) {
          int32_t *row = &m.tetra_grid[(size_t)tetra * TG_N];
          for (int v = 0; v < 4; v++) {
            const int code = CONF[orient][t][v];
            const int dth = code & 1, dr = (code >> 1) & 1, dsl = (code >> 2) & 1;
            const int segoff = (dth + (dr ? upper_off : lower_off)) % n3;
            const int sl = (slice + dsl) % n2;  // last slice wraps onto the first (wrap_idx_inplace)
            row[TG_KNOT + v] = (int32_t)(sl * verts_per_slice + base + (int64_t)dr * n3 + segoff + 1);
            row[TG_NEIGH + v] = -1;
            row[TG_NFACE + v] = -1;
          }
        }
        // periodic boundary in theta (circular_mesh.f90:404-429)
        int32_t *r0 = &m.tetra_grid[(size_t)(tetra - 3) * TG_N];
        if (orient == 0 && seg == 1) {
          r0[1 * TG_N + TG_PERTHETA + 3] = -1;
          r0[2 * TG_N + TG_PERTHETA + 3] = -1;
        } else if (orient == 1 && seg == 2 * n3) {
          r0[0 * TG_N + TG_PERTHETA + 0] = 1;
          r0[1 * TG_N + TG_PERTHETA + 0] = 1;
        }
        // periodic boundary in phi (:488-490): face 4 of the 1st tetra (first slice), face 1 of the 3rd (last slice)
        if (slice == 0) r0[0 * TG_N + TG_PERPHI + 3] = -1;
        if (slice == n2 - 1) r0[2 * TG_N + TG_PERPHI + 0] = 1;
      }
    }
  }
  // neighbours: the face opposite vertex f is face f; match faces by their vertex triple
  std::unordered_map<FaceKey, int64_t, FaceHash> faces;
  faces.reserve((size_t)m.ntetr * 2);
  for (int64_t t = 0; t < m.ntetr; t++) {
    int32_t *row = &m.tetra_grid[(size_t)t * TG_N];
    for (int f = 0; f < 4; f++) {
      int32_t v[3];
      int k = 0;
      for (int j = 0; j < 4; j++)
        if (j != f) v[k++] = row[TG_KNOT + j];
      std::sort(v, v + 3);
      FaceKey key{v[0], v[1], v[2]};
      auto it = faces.find(key);
      if (it == faces.end()) {
        faces.emplace(key, t * 4 + f);
      } else {
        const int64_t t2 = it->second / 4;
        const int f2 = (int)(it->second % 4);
        row[TG_NEIGH + f] = (int32_t)(t2 + 1);
        row[TG_NFACE + f] = f2 + 1;
        int32_t *row2 = &m.tetra_grid[(size_t)t2 * TG_N];
        row2[TG_NEIGH + f2] = (int32_t)(t + 1);
        row2[TG_NFACE + f2] = f + 1;
        faces.erase(it);
      }
    }
  }
}

int build_vmec(const gorilla_grid_settings &gs, const gorilla_settings &st, Mesh &m, std::string &err)
{
  if (st.coord_system != 2) { err = "grid_kind 3 requires coord_system = 2 (tetra_physics_mod.f90:248-250)"; return GORILLA_ERR_ARG; }
  if (!gs.netcdf_filename || !*gs.netcdf_filename) { err = "netcdf_filename is empty"; return GORILLA_ERR_ARG; }
  if (gs.n1 < 1 || gs.n2 < 3 || gs.n3 < 3) { err = "field-aligned grid needs n1 >= 1, n2 >= 3, n3 >= 3"; return GORILLA_ERR_ARG; }
  Vmec V;
  if (!load_vmec(gs.netcdf_filename, V, err)) return GORILLA_ERR_IO;
  const int n1 = gs.n1, n2 = gs.n2, n3 = gs.n3;
  m.grid_kind = 3;
  m.coord_system = 2;
  m.grid_size[0] = n1; m.grid_size[1] = n2; m.grid_size[2] = n3;
  m.n_field_periods = gs.boole_n_field_periods ? V.nfp : gs.n_field_periods_manual;
  m.sfc_s_min = gs.sfc_s_min;
  make_field_aligned_topology(m, n1, n2, n3);
  // vertices (points_2d.f90:171-252 + extrude_points): ring 0 at s_min, ring k at s_k; theta_j = j/n3 * 2 pi
  const int64_t vps = (int64_t)(n1 + 1) * n3;
  m.verts_sthetaphi.assign((size_t)m.nvert * 3, 0.0);
  m.verts_rphiz.assign((size_t)m.nvert * 3, 0.0);
  m.verts_theta_vmec.assign((size_t)m.nvert, 0.0);
  const double s_min = gs.sfc_s_min;
  const double sqrt_s_min = std::sqrt(s_min);
  VertexFields vf;
  vf.resize((size_t)m.nvert, true, true);
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t iv = 0; iv < m.nvert; iv++) {
    const int slice = (int)(iv / vps), ring = (int)((iv % vps) / n3), j = (int)(iv % n3);
    double s;
    if (ring == 0) s = s_min;
    else if (gs.i_radial_spacing == 2) {
      const double q = sqrt_s_min + (double)ring * (1.0 - sqrt_s_min) / (double)n1;
      s = q * q;
    } else {
      s = s_min + ((double)ring * (1.0 - s_min)) / (double)n1;
    }
    const double theta = ((double)j / (double)n3) * 2.0 * PI_MESH;
    const double phi = slice == 0 ? 0.0 : (2.0 * PI_MESH / m.n_field_periods * slice) / n2;
    double *vs = &m.verts_sthetaphi[3 * iv];
    vs[0] = s; vs[1] = theta; vs[2] = phi;
    const double thv = theta_vmec_of(V, s, theta, phi);
    m.verts_theta_vmec[iv] = thv;
    double f[10];
    vertex_field(V, s, thv, phi, m.bmod_multiplier, f);
    Vmec::Point P;
    V.eval(s, thv, phi, P);
    double *vr = &m.verts_rphiz[3 * iv];
    vr[0] = P.R; vr[1] = phi; vr[2] = P.Z;
    vf.A_x1[iv] = f[0]; vf.A_x2[iv] = f[1]; vf.A_x3[iv] = f[2];
    vf.h_x1[iv] = f[3]; vf.h_x2[iv] = f[4]; vf.h_x3[iv] = f[5];
    vf.bmod[iv] = f[6]; vf.sqg[iv] = f[7]; vf.dR_ds[iv] = f[8]; vf.dZ_ds[iv] = f[9];
    vf.phi_elec[iv] = vf.A_x2[iv] * st.eps_Phi;
  }
  // magnetic axis (tetra_physics_mod.f90:311-313): R, Z at s -> 0
  {
    Vmec::Point P;
    V.eval(1e-16, 0.1, 0.1, P);
    m.mag_axis_R0 = P.R;
    m.mag_axis_Z0 = P.Z;
  }
  m.Rmin = m.Rmax = m.Zmin = m.Zmax = 0.0;
  apply_vertex_noise(m, st, vf);
  linearise_tetrahedra(m, vf);
  check_tetra_overlaps(m);
  return GORILLA_OK;
}

}  // namespace gbhost
