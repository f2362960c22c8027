// gb_find.cuh -- initial localisation of a particle in the tetrahedral mesh + coordinate-domain check.
//
// Replaces (reference file:line):
//   find_tetra                 SRC/find_tetra_mod.f90:283-600  (boole_grid_for_find_tetra = .false.)
//   isinside                   SRC/tetra_physics_mod.f90:1038-1072
//   check_coordinate_domain    SRC/orbit_timestep_gorilla.f90:278-358
//   and the RK-module pieces find_tetra uses for starts that lie on a face:
//   initialize_pusher_tetra_rk_mod (ODE coefficients only) SRC/pusher_tetra_rk.f90:50-134,
//   rk4_step :840-896, normal_distances_func :2422, normal_velocity_func :2451.
//
// The reference scans the tetrahedra of one phi-slice in index order and takes the FIRST one that
// contains the point (tolerance eps*|dist_ref|).  The scan order is part of the result, so it is kept:
// one lane walks the slice; all lanes of a warp that sit in the same slice read the same 128-byte
// geometry records, which the L1/L2 serve as broadcasts.
#pragma once
#include "gb_rk.cuh"

namespace gb {

#define GB_EPS 1.e-10

// Fortran modulo(a,p) for reals as gfortran expands it (trans-intrinsic.cc, gfc_conv_intrinsic_mod): r = fmod(a,p), which is
// exact; r += p when r != 0 and its sign differs from p's; a zero result takes the sign of p.  (a - floor(a/p)*p rounds k*p
// and differs in the last bits once |a| >= 2p.)
GB_HD double f_modulo(double a, double p)
{
  double r = fmod(a, p);
  if (r != 0.0 && ((r < 0.0) != (p < 0.0))) r += p;
  if (r == 0.0) r = copysign(0.0, p);
  return r;
}

// returns 0 ok, 1 outside the computation domain (reference: print + stop)
GB_HD int check_coordinate_domain(const MeshDev &m, double *x, int boole_periodic_relocation)
{
  if (m.coord_system == 1) {
    if (boole_periodic_relocation) x[1] = f_modulo(x[1], m.period_phi);
    else if (x[1] < 0.0 || x[1] > m.period_phi) return 1;
  } else {
    if (x[0] < m.sfc_s_min || x[0] > 1.0) return 1;
    if (boole_periodic_relocation) {
      x[1] = f_modulo(x[1], m.period_theta);
      x[2] = f_modulo(x[2], m.period_phi);
    } else if (x[1] < 0.0 || x[1] > m.period_theta || x[2] < 0.0 || x[2] > m.period_phi) {
      return 1;
    }
  }
  return 0;
}

GB_HD bool isinside(const MeshDev &m, int64_t ind_tetr, const double *x, double *dist, double &dist_ref)
{
  const double *g = m.geom + (ind_tetr - 1) * GEOM_ND;
  double v[GEOM_ND];
#pragma unroll
  for (int i = 0; i < GEOM_ND; i += 2) ld2(g + i, v[i], v[i + 1]);
  dist_ref = v[3];
  const double dist_min = GB_EPS * fabs(dist_ref);
  const double d[3] = {x[0] - v[0], x[1] - v[1], x[2] - v[2]};
  bool all_ok = true;
#pragma unroll
  for (int f = 0; f < 4; f++) {
    double s = dot3(&v[4 + 3 * f], d);
    if (f == 0) s = s + dist_ref;
    dist[f] = s;
    if (!(s >= -dist_min)) all_ok = false;
  }
  return all_ok;
}

// Start point lies (within tolerance) on >= 1 face: hop through the neighbours until every converged
// face has inward normal velocity (find_tetra_mod.f90:478-581).
template <int PHI>
GB_HD_NOINLINE void find_tetra_on_face(const MeshDev *mp, double *x, double vpar, double vperp, int32_t &ind_tetr_out,
                                       int32_t &iface, int sign_t_step, const double *dist0, int n_plane_conv)
{
  const MeshDev &m = *mp;
  // reuse record loader + ODE coefficient builder (same formulas, :104-116 vs poly :1504-1517); EXT = 2 carries the
  // run-time hand-over kind (pusher_handover2neighbour honours handover_processing_kind = 2 here too, find_tetra_mod.f90:558)
  PolyPusher<1, PHI, 2> P;
  double stash[6];
  P.mp = mp;
  P.r.set_stash(stash, 1);
  int iface_new = 1;
  for (int f = 1; f < 4; f++)
    if (fabs(dist0[f]) < fabs(dist0[iface_new - 1])) iface_new = f + 1;
  int32_t tried[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  {
    Rec<PHI> r0;
    double stash0[6];
    r0.set_stash(stash0, 1);
    r0.load(m, ind_tetr_out);
    double z[3] = {x[0] - r0.x1[0], x[1] - r0.x1[1], x[2] - r0.x1[2]};
    P.perpinv = -0.5 * (vperp * vperp) / (r0.bmod1 + dot3(r0.gB, z));
  }
  for (int i_try = 1; i_try <= 2 * n_plane_conv; i_try++) {
    if (ind_tetr_out == -1) break;
    tried[i_try - 1] = ind_tetr_out;
    // ODE coefficients of this tetrahedron (b, amat, Bvec, spamat) for t_remain = sign_t_step
    P.init(ind_tetr_out, x, iface_new, vpar, (double)sign_t_step);
    P.template build_ode<true>();
    double z[4] = {P.z_init[0], P.z_init[1], P.z_init[2], vpar};
    double dist[4];
    bool conv[4];
    for (int f = 0; f < 4; f++) {
      dist[f] = P.normal_distance(z, f + 1);
      conv[f] = fabs(dist[f]) <= (GB_EPS * fabs(P.r.dist_ref));
    }
    // rk4_step(z, 0.d0, dzdtau): all stages collapse onto z; dzdtau = rhs at the last stage point
    double dydx[4], yt[4], dyt[4], dym[4];
    bm_vec_rk(dydx, P, z);
    for (int i = 0; i < 4; i++) yt[i] = z[i] + 0.0 * dydx[i];
    bm_vec_rk(dyt, P, yt);
    for (int i = 0; i < 4; i++) yt[i] = z[i] + 0.0 * dyt[i];
    bm_vec_rk(dym, P, yt);
    for (int i = 0; i < 4; i++) yt[i] = z[i] + 0.0 * dym[i];
    bm_vec_rk(dyt, P, yt);
    int counter_vnorm_pos = 0;
    double vn[4];
    for (int l = 0; l < 4; l++) {
      vn[l] = dot3(dyt, P.r.an[l]);
      if (m.newton_precalc) {   // normal_velocity_func of the RK module with boole_newton_precalc (pusher_tetra_rk.f90:2487-2505)
        const double *q = P.p4() + P4_AN_AMAT + 4 * l;
        double sacc = 0.0;
        for (int i = 0; i < 4; i++) sacc = sacc + (ldg(q + i) + P.perpinv * ldg(q + 16 + i)) * z[i];
        vn[l] = sacc * (double)P.sign_rhs + dot3(P.r.an[l], P.b);
      }
      if (conv[l] && vn[l] > 0.0) counter_vnorm_pos++;
    }
    if (counter_vnorm_pos == n_plane_conv) {
      iface = iface_new;
      break;
    }
    const int ind_tetr_save = ind_tetr_out, iface_new_save = iface_new;
    const double xs[3] = {x[0], x[1], x[2]};
    for (int l = 1; l <= 4; l++) {
      if (!conv[l - 1]) continue;
      if (vn[l - 1] > 0.0) continue;
      int32_t out, fout;
      P.handover(l, x, out, fout);
      ind_tetr_out = out;
      iface_new = fout;
      bool was_tried = false;
      for (int t = 0; t < 2 * n_plane_conv; t++)
        if (tried[t] == out) was_tried = true;
      if (was_tried || out == -1) {
        x[0] = xs[0]; x[1] = xs[1]; x[2] = xs[2];
        iface_new = iface_new_save;
      } else {
        break;
      }
    }
    (void)ind_tetr_save;
  }
}

template <int PHI>
GB_HD void find_tetra(const MeshDev *mp, double *x, double vpar, double vperp, int32_t &ind_tetr_out, int32_t &iface,
                      int sign_t_step)
{
  const MeshDev &m = *mp;
  const double PI = 3.141592653589793238462643383;
  int64_t indtetr_start, ntetr_searched;
  const int64_t ntetr = m.ntetr;
  int corr_plus = 0, corr_minus = 0;
  const int nphi = m.grid_size2;
  ind_tetr_out = -1;
  iface = -1;
  if (m.grid_kind == 1 || m.grid_kind == 5) {
    const int nr = m.grid_size1, nz = m.grid_size3;
    const double hr = (m.Rmax - m.Rmin) / nr, hphi = (2.0 * PI) / nphi, hz = (m.Zmax - m.Zmin) / nz;
    const int ir = (int)((x[0] - m.Rmin) / hr) + 1, iphi = (int)(x[1] / hphi) + 1, iz = (int)((x[2] - m.Zmin) / hz) + 1;
    if (ir < 1 || ir > nr || iphi < 1 || iphi > nphi || iz < 1 || iz > nz) return;
    indtetr_start = (int64_t)(((double)iz - 1.0) * 6.0 + 6.0 * (double)nz * ((double)ir - 1.0) +
                              6.0 * ((double)iphi - 1.0) * (double)nr * (double)nz + 1.0);
    ntetr_searched = 6;
  } else {
    const int ind_b = (m.coord_system == 2) ? 2 : 1;
    const int64_t ntetr_in_plane = ntetr / nphi;
    const double q = x[ind_b] * nphi / (2.0 * PI / m.n_field_periods);
    const int ind_plane = (int)q;
    if (fabs(q - (double)ind_plane) > (1.0 - GB_EPS)) corr_plus = 1;
    if (fabs(q - (double)ind_plane) < GB_EPS) corr_minus = 1;
    indtetr_start = (int64_t)ind_plane * ntetr_in_plane + 1;
    ntetr_searched = ntetr_in_plane * (1 + corr_plus + corr_minus);
  }
  // one candidate: the reference's loop body.  Returns true when the search is over.
  auto try_tetra = [&](int64_t ind) -> bool {
    double dist[4], dist_ref;
    if (!isinside(m, ind, x, dist, dist_ref)) return false;
    ind_tetr_out = (int32_t)ind;
    iface = 0;
    int n_plane_conv = 0;
#pragma unroll
    for (int f = 0; f < 4; f++)
      if (fabs(dist[f]) <= (GB_EPS * fabs(dist_ref))) n_plane_conv++;
    if (n_plane_conv > 0)
      find_tetra_on_face<PHI>(mp, x, vpar, vperp, ind_tetr_out, iface, sign_t_step, dist, n_plane_conv);
    if (ind_tetr_out == -1) {
      iface = -1;
      return false;
    }
    return true;
  };
  if (m.bin_start && !(m.grid_kind == 1 || m.grid_kind == 5)) {
    // binned search: same visiting order as the scan below (slice after slice, ascending index inside a slice), but only
    // the tetrahedra whose 2-D bounding box covers the point
    const int64_t tps = ntetr / nphi;
    const int nslices = 1 + corr_plus + corr_minus;
    // u, v are taken once: a failed start-on-a-face attempt can only leave x shifted by a whole period, which no tetrahedron
    // contains (isinside below always sees the current x)
    const double u = x[m.bin_c0], v = x[m.bin_c1];
    for (int k = 0; k < nslices; k++) {
      int64_t base = indtetr_start - 1 + (int64_t)(k - corr_minus) * tps;  // 0-based first tetrahedron of the slice
      if (base >= ntetr) base -= ntetr;
      if (base < 0) base += ntetr;
      const int iu = (int)((u - m.bin_u0) * m.bin_du_inv), iv = (int)((v - m.bin_v0) * m.bin_dv_inv);
      if (u < m.bin_u0 || v < m.bin_v0 || iu >= m.bin_nu || iv >= m.bin_nv) continue;
      const int b = iv * m.bin_nu + iu;
      for (int32_t q = m.bin_start[b]; q < m.bin_start[b + 1]; q++)
        if (try_tetra(base + m.bin_items[q] + 1)) return;
    }
    return;
  }
  for (int64_t i = 1; i <= ntetr_searched; i++) {
    int64_t ind = indtetr_start + i - 1 - (int64_t)corr_minus * (ntetr / nphi);
    if (ind > ntetr) ind -= ntetr;
    if (ind <= 0) ind += ntetr;
    if (try_tetra(ind)) break;
  }
}

// boole_full_orbit (SRC/gorilla_plot_mod.f90:553-579): what the reference writes next to the orbit point after a push,
// p_phi_func(vpar, z_save, ind_tetr_save) and energy_tot_func([z_save, vpar], perpinv, ind_tetr_save)
// (SRC/supporting_functions_mod.f90:377-408, 279-301), z = z_save relative to the first vertex of tetrahedron it (1-based)
GB_HD void orbit_point_invariants(const MeshDev &m, int32_t it, const double *z, double vpar, double perpinv, double &p_phi,
                                  double &e_tot)
{
  const double *pb = m.bpart + ((int64_t)it - 1) * BPART_ND;
  const double *pc = m.cold + ((int64_t)it - 1) * COLD_ND;
  const double gB[3] = {pb[B_GB], pb[B_GB + 1], pb[B_GB + 2]};
  double vperp = 0.0;  // vperp_func :321-338
  if (perpinv != 0.0) vperp = sqrt(2.0 * fabs(perpinv) * (pb[B_BMOD1] + dot3(gB, z)));
  double phi = 0.0;
  if (m.phi) {
    const double *pp = m.phi + ((int64_t)it - 1) * PHI_ND;
    const double gP[3] = {pp[P_GPHI], pp[P_GPHI + 1], pp[P_GPHI + 2]};
    phi = pp[P_PHI1] + dot3(gP, z);
  }
  e_tot = m.particle_mass / 2.0 * (vperp * vperp + vpar * vpar) + m.particle_charge * phi;
  const double *ps = m.se ? m.se + ((int64_t)it - 1) * SE_ND : nullptr;
  if (ps) {
    const double g2[3] = {ps[S_GV2EMOD], ps[S_GV2EMOD + 1], ps[S_GV2EMOD + 2]};
    e_tot = e_tot + 0.5 * m.particle_mass * (ps[S_V2EMOD1] + dot3(z, g2));
  }
  const double gh[3] = {pc[C_GHPHI], pc[C_GHPHI + 1], pc[C_GHPHI + 2]};
  const double gA[3] = {pc[C_GAPHI], pc[C_GAPHI + 1], pc[C_GAPHI + 2]};
  p_phi = m.particle_mass * vpar * (pc[C_HPHI1] + dot3(gh, z)) + m.particle_mass / m.cm_over_e * (pc[C_APHI1] + dot3(gA, z));
  if (ps) {
    const double gv[3] = {ps[S_GVE2], ps[S_GVE2 + 1], ps[S_GVE2 + 2]};
    p_phi = p_phi + m.particle_mass * (ps[S_VE2_1] + dot3(z, gv));
  }
}

} // namespace gb
