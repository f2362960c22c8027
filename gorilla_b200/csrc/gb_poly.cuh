// gb_poly.cuh -- polynomial tetrahedron pusher (orders K = 1..4), FP64, one particle per lane.
//
// Replaces (reference file:line), for i_precomp = 0, non-adaptive steps, handover_processing_kind = 1;
// i_time_tracing_option = 1 in the plain variant, 1 or 2 (Hamiltonian time) plus the optional quantities
// (t_hamiltonian, gyrophase, int v_par dt, int v_par^2 dt) in the EXT variant:
//   initialize_pusher_tetra_poly          SRC/pusher_tetra_poly.f90:125-178
//   pusher_tetra_poly                     :182-675
//   check_three_planes / _face_convergence / _velocity / _exit_time   :679-758
//   prolonged_trajectory                  :762-826
//   analytic_approx                       :1258-1482
//   analytic_coeff_without_precomp        :1486-1586
//   analytic_integration_without_precomp / set_integration_coef_manually   :2047-2113
//   normal_distance_func / normal_velocity_func / normal_v_func_from_trajectory   :2690-2775
//   physical_estimate_tau                 :2779-2831
//   trouble_shooting_polynomial_solver    :2835-2998
//   pusher_handover2neighbour             SRC/pusher_tetra_func_mod.f90:6-93
//   EXT: calc_t_hamiltonian / get_t_hamiltonian_root / calc_optional_quantities / z_series_coef /
//        poly_multiplication_coef / moment_integration   :2117-2536, 3000-3150 ; hamiltonian_time record
//        SRC/tetra_physics_mod.f90:105-114, 926-944
//
// Design (not a transliteration):
//   * The ODE matrix is block structured, A = [[a(3x3), c(3)], [0 0 0, s]].  Powers A^2..A^4 are
//     formed as block products (structural zeros skipped -- adding an exact 0 product never changes
//     an IEEE sum of finite terms), in the reference's accumulation order, so results are
//     bit-identical to matmul(amat, amat^k) while costing 39 instead of 112 flops per product.
//   * push_fast() is the branch-light common case (first attempt succeeds, particle leaves through a
//     face or stops inside); anything else makes it return false WITHOUT side effects and the caller
//     re-runs the push through push_full(), the complete fall-back ladder, which is a separate
//     non-inlined function so that its register/stack footprint does not tax the hot loop.
//   * The reference's THREADPRIVATE module state is the PolyPusher object; it lives in registers.
//   * The reference traps on FP exceptions (CMakeLists.txt:24-25); here a non-finite state simply
//     finds no valid exit time on its next push (every root test 0 < dtau < huge fails for NaN) and
//     the particle is removed like any unrecoverable push (ind_tetr = -1, iface = -1).
#pragma once
#include "gb_mesh.cuh"
#include "gb_roots.cuh"

namespace gb {

#define GB_CLIGHT 2.9979e10
#define GB_EPS_TAU 100.0

GB_HD double dot3(const double *a, const double *b) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }

// block matrix [[m, c],[0, s]]
struct BlockMat {
  double m[3][3], c[3], s;
};
// P = A * B (reference: matmul(A,B), k ascending)
GB_HD void bm_mul(BlockMat &P, const BlockMat &A, const BlockMat &B)
{
#pragma unroll
  for (int i = 0; i < 3; i++) {
#pragma unroll
    for (int j = 0; j < 3; j++) P.m[i][j] = (A.m[i][0] * B.m[0][j] + A.m[i][1] * B.m[1][j]) + A.m[i][2] * B.m[2][j];
    P.c[i] = ((A.m[i][0] * B.c[0] + A.m[i][1] * B.c[1]) + A.m[i][2] * B.c[2]) + A.c[i] * B.s;
  }
  P.s = A.s * B.s;
}
GB_HD void bm_vec(double *o, const BlockMat &A, const double *v)
{
#pragma unroll
  for (int i = 0; i < 3; i++) o[i] = ((A.m[i][0] * v[0] + A.m[i][1] * v[1]) + A.m[i][2] * v[2]) + A.c[i] * v[3];
  o[3] = A.s * v[3];
}

struct PushOut {
  double x[3], vpar, z_save[3], t_pass;
  int32_t ind_tetr, iface;
  int32_t finished;  // boole_t_finished
  int32_t z_save_set;
  int32_t fallback;  // bit0 2nd attempt, bit1 trouble shooting, bit2 prolonged, bit3 finish-outside, bit4 adaptive sub-steps
};

// an exit-time solve that has been set up but not executed (see PolyPusher::face_task)
struct SolveTask {
  double q[4], lambda, tau;
  int deg, kind;
};

// orbit events of one push (EXT = 2): per-particle state in/out and at most one event of each kind
struct EvRec {
  int kind, counter;   // 1 = toroidal mapping (phi = 0), 2 = banana tip (v_par = 0)
  double x[3], v[2];   // position; {p_phi, e_tot} resp. {J_par, e_tot}
};
struct EvState {
  int flags;           // bit0 boole_poincare_phi_0, bit1 boole_poincare_vpar_0, bit2 boole_J_par
  int nskip_p, nskip_v;
  double J;            // par_adiab_inv
  int cnt_v, cnt_p;    // counter_banana_mappings, counter_phi_0_mappings
  int n;
  EvRec e[2];
};

// EXT: 0 = plain (i_time_tracing_option = 1, no optional quantities); 1 = Hamiltonian time tracing, nothing else;
// 2 = time tracing option read at run time + optional quantities
template <int K, int PHI, int EXT = 0>
struct PolyPusher {
  const MeshDev *mp;
  Rec<PHI> r;
  double perpinv;
  int ind_tetr, iface_init, sign_rhs, nsteps, solver_iters, fallback;
  // EXT: tau_steps_list / intermediate_z0_list (:36-38; two entries in the non-adaptive scheme, :98-104) and the
  // optional quantities of this push
  double tau_list[2], z0_list[2][4], oq[4];
  unsigned oq_mask;  // bit0 boole_time_Hamiltonian, bit1 boole_gyrophase, bit2 boole_vpar_int, bit3 boole_vpar2_int
  double z0_last[4]; // EXT = 3: intermediate_z0_list(:, number_of_integration_steps), the only list entry the adaptive scheme reads
  int n_adaptive;    // EXT = 3: segments re-integrated in sub-steps during this push
  bool main_fc;      // EXT = 3: boole_face_correct of pusher_tetra_poly when the final processing starts: .false. after
                     // the third attempt (trouble shooting keeps its own copy), which disables the sub-stepping of the
                     // stop-inside segment (:564-568 with :893)
  // EXT = 5 = adaptive sub-stepping combined with the list consumers (Hamiltonian time, optional quantities, J_par): the
  // kernels of EXT = 2 and EXT = 3 in one, with tau_steps_list / intermediate_z0_list kept in full (3 * max_n_intermediate_steps
  // entries, :98-101) in a per-thread global scratch region
  static constexpr bool OPT = (EXT == 2 || EXT == 5);     // run-time options: optional quantities, events, hand-over kind 2
  static constexpr bool ADAPT = (EXT == 3 || EXT == 5);   // boole_adaptive_time_steps
  static constexpr bool LONG = (EXT == 5);
  double *lst;       // EXT = 5: [entry][5] = tau_steps_list(i), intermediate_z0_list(1:4, i)
  int lst_cap;
  int iper_phi;      // EXT = 2: toroidal period crossed by the hand-over (+1 / -1 / 0), for the phi = 0 mappings
  bool removed;      // EXT = 2: the push ended on one of the "remove particle" returns
  double dt_dtau_const, bmod0, vmod0, t_remain, z_init[4], k1, k3, dv2E;
  BlockMat A;
  double b[4];
  double Az[4], Ab[4], A2z[4], A2b[4], A3z[4], A3b[4], A4z[4];

  // ---- :125-178
  GB_HD void init(int ind_tetr_in, const double *x, int iface, double vpar, double t_remain_in)
  {
    t_remain = t_remain_in;
    ind_tetr = ind_tetr_in;
    r.load(*mp, ind_tetr);
    sign_rhs = mp->sign_sqg * (signbit(t_remain) ? -1 : 1);
#pragma unroll
    for (int i = 0; i < 3; i++) z_init[i] = x[i] - r.x1[i];
    z_init[3] = vpar;
    iface_init = iface;
    dt_dtau_const = r.dtdtau * (double)sign_rhs;
    bmod0 = r.bmod1 + dot3(r.gB, z_init);
    double vperp2 = -2.0 * perpinv * bmod0;
    double vpar2 = vpar * vpar;
    vmod0 = sqrt(vpar2 + vperp2);
    k1 = vperp2 + vpar2 + 2.0 * perpinv * r.bmod1;
    // strong electric field (:175): k1 += v_E^2(z) - v_E^2(x1); kept separately because the RK module adds it to b
    // as its own term (pusher_tetra_rk.f90:127-128)
    if (PHI == 2) dv2E = (r.v2Emod1 + dot3(z_init, r.gv2Emod)) - r.v2Emod1;
    if (PHI) {
      double phi_elec = r.Phi1 + dot3(r.gPhi, z_init);
      k3 = r.Phi1 - phi_elec;
    } else {
      k3 = 0.0;
    }
    nsteps = 0;
    if (ADAPT) {
      n_adaptive = 0;
      main_fc = true;
    }
    if (OPT) {
      oq[0] = oq[1] = oq[2] = oq[3] = 0.0;  // initialise_optional_quantities (:2117-2130)
      iper_phi = 0;
      removed = false;
    }
    if (EXT == 4) {
      // i_precomp = 2 never assigns the module variable b (it is only read by normal_velocity_func): it stays the
      // zero-initialised static it is in the reference build
      b[0] = b[1] = b[2] = b[3] = 0.0;
    }
  }

  // ==== EXT = 4: precomputed coefficients, i_precomp = 1, 2 (analytic_coeff_with_precomp :1590-1725,
  // analytic_integration_with_precomp :2530-2650, normal_velocity_func :2734-2738; record tetra_physics_poly4) ============
  GB_HD const double *p4() const { return mp->poly4 + ((int64_t)ind_tetr - 1) * P4_ND; }
  // index of the first of the ORD+1 matrices of order ORD among the 14: amat1_* 0, amat2_* 2, amat3_* 5, amat4_* 9
  GB_HD static constexpr int p4_first(int ord) { return ord == 1 ? 0 : ord == 2 ? 2 : ord == 3 ? 5 : 9; }
  // sum_i (M_0 + p M_1 + p^2 M_2 + ...)(i, n) * v(i) over the anorm_in_amat<ORD>_* columns of face n (ascending powers, :1633-1715)
  template <int ORD>
  GB_HD double p4_face_dot(const double *q, int n, const double *fac, const double *v) const
  {
    const double *base = q + P4_AN_AMAT + 16 * p4_first(ORD) + 4 * n;
    double sacc = 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      double e = ldg(base + i);
#pragma unroll
      for (int k = 1; k <= ORD; k++) e = e + fac[k] * ldg(base + 16 * k + i);
      sacc = sacc + e * v[i];
    }
    return sacc;
  }
  // element (i,j) of p^ORD amat<ORD>_<ORD> + ... + p amat<ORD>_1 + amat<ORD>_0 (DESCENDING powers, :2558-2640)
  template <int ORD>
  GB_HD double p4_op_elem(const double *q, int i, int j, const double *fac) const
  {
    const double *base = q + P4_AMAT + 16 * p4_first(ORD) + i + 4 * j;
    double e = fac[ORD] * ldg(base + 16 * ORD);
#pragma unroll
    for (int k = ORD - 1; k >= 1; k--) e = e + fac[k] * ldg(base + 16 * k);
    return e + ldg(base);
  }
  GB_HD void p4_factors(double *fac) const
  {
    const double perpinv2 = perpinv * perpinv;
    fac[0] = 1.0; fac[1] = perpinv; fac[2] = perpinv2; fac[3] = perpinv2 * perpinv; fac[4] = perpinv2 * perpinv2;
  }
  template <int ORD>
  GB_HD void face_coeffs_precomp(int n /*0-based face*/, const double *z, double *c) const
  {
    const double *q = p4();
    double fac[5];
    p4_factors(fac);
    {
      double nn[3];
      face_normal(n + 1, nn);
      c[0] = dot3(nn, z);
      if (n == 0) c[0] = c[0] + r.dist_ref;
      const bool one = mp->i_precomp == 1;
      if (ORD >= 1) {
        const double sz = p4_face_dot<1>(q, n, fac, z);
        if (one) c[1] = sz + dot3(nn, b);
        else c[1] = sz + ldg(q + P4_AN_B0 + n) + k1 * ldg(q + P4_AN_B0 + 4 + n) + perpinv * ldg(q + P4_AN_B0 + 8 + n) +
                    k3 * ldg(q + P4_AN_B0 + 12 + n);
      }
      if (ORD >= 2) {
        const double sz = p4_face_dot<2>(q, n, fac, z);
        if (one) {
          c[2] = sz + p4_face_dot<1>(q, n, fac, b);
        } else {
          const double *a0 = q + P4_AN_A10_B0 + n, *a1 = q + P4_AN_A11_B0 + n;   // [.. + 4 k]: b0..b3
          c[2] = sz + ldg(a0) + perpinv * ldg(a1) + k1 * (ldg(a0 + 4) + perpinv * ldg(a1 + 4)) +
                 perpinv * (ldg(a0 + 8) + perpinv * ldg(a1 + 8)) + k3 * (ldg(a0 + 12) + perpinv * ldg(a1 + 12));
        }
      }
      if (ORD >= 3) c[3] = p4_face_dot<3>(q, n, fac, z) + p4_face_dot<2>(q, n, fac, b);
      if (ORD >= 4) c[4] = p4_face_dot<4>(q, n, fac, z) + p4_face_dot<3>(q, n, fac, b);
    }
  }
  template <int ORD>
  GB_HD void integrate_precomp(double *z, double tau) const
  {
    if (ORD < 2) return;                       // no case(1) in the reference: z unchanged
    const bool one = mp->i_precomp == 1;
    if (!one && ORD != 2) return;              // i_precomp = 2 has the order-2 case only
    const double *q = p4();
    double fac[5];
    p4_factors(fac);
    const double tau2_half = tau * tau * 0.5, tau3_sixth = (tau * tau) * tau / 6.0;
    const double t2 = tau * tau, tau4_24 = (t2 * t2) / 24.0;
    double oz[4], ob[4];
#pragma unroll 1
    for (int i = 0; i < 4; i++) {
      double sz = 0.0, sb = 0.0;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const double c1 = p4_op_elem<1>(q, i, j, fac), c2 = p4_op_elem<2>(q, i, j, fac);
        double e = tau * c1 + tau2_half * c2;
        double f = tau * (i == j ? 1.0 : 0.0) + tau2_half * c1;
        if (ORD >= 3) {
          const double c3 = p4_op_elem<(ORD >= 3 ? 3 : 1)>(q, i, j, fac);
          e = e + tau3_sixth * c3;
          f = f + tau3_sixth * c2;
          if (ORD >= 4) {
            e = e + tau4_24 * p4_op_elem<(ORD >= 4 ? 4 : 1)>(q, i, j, fac);
            f = f + tau4_24 * c3;
          }
        }
        sz = sz + e * z[j];
        sb = sb + f * b[j];
      }
      oz[i] = sz;
      ob[i] = sb;
    }
    if (!one) {   // operator_b_in_b (:2578-2587)
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const double *a0 = q + P4_A10_B0 + i, *a1 = q + P4_A11_B0 + i;
        ob[i] = tau * (ldg(q + P4_B0 + i) + k1 * ldg(q + P4_B1 + i) + perpinv * ldg(q + P4_B2 + i) + k3 * ldg(q + P4_B3 + i)) +
                tau2_half * ((ldg(a0) + perpinv * ldg(a1)) + k1 * (ldg(a0 + 4) + perpinv * ldg(a1 + 4)) +
                             perpinv * (ldg(a0 + 8) + perpinv * ldg(a1 + 8)) + k3 * (ldg(a0 + 12) + perpinv * ldg(a1 + 12)));
      }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) z[i] = z[i] + ob[i] + oz[i];
  }
  GB_HD double normal_velocity_precomp(const double *z, int iface) const
  {
    const double *q = p4() + P4_AN_AMAT + 4 * (iface - 1);
    double n[3], sacc = 0.0;
    face_normal(iface, n);
#pragma unroll
    for (int i = 0; i < 4; i++) sacc = sacc + (ldg(q + i) + perpinv * ldg(q + 16 + i)) * z[i];
    return sacc * (double)sign_rhs + dot3(n, b);
  }

  // ---- ODE coefficients b, A  (:1503-1530; strong-electric-field terms :1519-1526).  RK = true forms b(1:3) the
  // way the RK module writes it (pusher_tetra_rk.f90:104-134: k1 spelled out, the v_E^2 part a separate term).
  template <bool RK = false>
  GB_HD void build_ode()
  {
    const double cm = mp->cm_over_e, sg = (double)sign_rhs;
    const double pc = perpinv * cm;
    const double k1_eff = (PHI == 2 && !RK) ? k1 + dv2E : k1;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      double t = (r.curlh[i] * k1_eff + perpinv * r.gBxh1[i]) * cm;
      if (PHI) t = t - GB_CLIGHT * (2.0 * k3 * r.curlh[i] + r.gPhixh1[i]);
      if (PHI == 2) {
        if (RK) t = t - 0.5 * cm * r.gv2Emodxh1[i] + cm * r.curlh[i] * dv2E;
        else t = t - 0.5 * cm * r.gv2Emodxh1[i];
      }
      b[i] = t * sg;
    }
    {
      double t = perpinv * r.gBxcurlA;
      if (PHI) t = t - GB_CLIGHT / cm * r.gPhixcurlA;
      if (PHI == 2)
        t = t + cm * perpinv * r.gBxcurlvE - GB_CLIGHT * r.gPhixcurlvE - 0.5 * cm * r.gv2EmodxcurlvE - 0.5 * r.gv2EmodxcurlA;
      b[3] = t * sg;
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
#pragma unroll
      for (int j = 0; j < 3; j++) {
        double t = pc * r.alp[i + 3 * j];
        if (PHI) t = t - GB_CLIGHT * r.bet[i + 3 * j];
        if (PHI == 2) t = t - 0.5 * cm * r.gam[i + 3 * j];
        A.m[i][j] = t * sg;
      }
      double c = r.curlA[i];
      if (PHI == 2) c = c + cm * r.curlvE[i];
      A.c[i] = c * sg;
    }
    {
      double t = pc * r.spalp;
      if (PHI) t = t - GB_CLIGHT * r.spbet;
      if (PHI == 2) t = t - 0.5 * cm * r.spgam;
      A.s = t * sg;
    }
  }

  // ---- ODE coefficients and the Taylor vectors A^k z, A^(k-1) b up to order ORD (:1503-1578, face independent)
  template <int ORD>
  GB_HD void prepare(const double *z)
  {
    if (EXT == 4) {
      if (mp->i_precomp == 1) {   // b WITHOUT sign_rhs (:1607-1614); no matrix, no Taylor vectors
        const double cm = mp->cm_over_e;
#pragma unroll
        for (int i = 0; i < 3; i++) {
          double t = (r.curlh[i] * k1 + perpinv * r.gBxh1[i]) * cm;
          if (PHI) t = t - GB_CLIGHT * (2.0 * k3 * r.curlh[i] + r.gPhixh1[i]);
          else t = t - GB_CLIGHT * (2.0 * k3 * r.curlh[i] + 0.0);
          b[i] = t;
        }
        double t = perpinv * r.gBxcurlA;
        if (PHI) t = t - GB_CLIGHT / cm * r.gPhixcurlA;
        else t = t - GB_CLIGHT / cm * 0.0;
        b[3] = t;
      }
      return;
    }
    build_ode();
    if (ORD >= 1) bm_vec(Az, A, z);
    if (ORD >= 2) {
      BlockMat A2;
      bm_mul(A2, A, A);
      bm_vec(A2z, A2, z);
      bm_vec(Ab, A, b);
      if (ORD >= 3) {
        BlockMat A3;
        bm_mul(A3, A, A2);
        bm_vec(A3z, A3, z);
        bm_vec(A2b, A2, b);
        if (ORD >= 4) {
          BlockMat A4;
          bm_mul(A4, A, A3);
          bm_vec(A4z, A4, z);
          bm_vec(A3b, A3, b);
        }
      }
    }
  }
  // ---- face-distance Taylor coefficients of one face with normal n (:1536-1584); c[k] = coef_mat(face,k+1)
  template <int ORD>
  GB_HD void face_coeffs(const double *n, bool is_face1, const double *z, double *c, int face0 = 0) const
  {
    if (EXT == 4) {
      face_coeffs_precomp<ORD>(face0, z, c);
      return;
    }
    c[0] = dot3(n, z);
    if (is_face1) c[0] = c[0] + r.dist_ref;  // coef - dist1, dist1 = -dist_ref
    if (ORD >= 1) c[1] = dot3(n, Az) + dot3(n, b);
    if (ORD >= 2) c[2] = dot3(n, A2z) + dot3(n, Ab);
    if (ORD >= 3) c[3] = dot3(n, A3z) + dot3(n, A2b);
    if (ORD >= 4) c[4] = dot3(n, A4z) + dot3(n, A3b);
  }
  template <int ORD>
  GB_HD void coeff(unsigned mask, const double *z, double cm[4][5])
  {
    prepare<ORD>(z);
#pragma unroll
    for (int f = 0; f < 4; f++)
      if (mask & (1u << f)) face_coeffs<ORD>(r.an[f], f == 0, z, cm[f], f);
  }

  // ---- :2087-2113 (A, b unchanged; matrix powers re-formed, which reproduces the stored ones)
  GB_HD void set_integration_coef_manually(const double *z0) { set_coef<K>(z0); }
  template <int ORD>
  GB_HD void set_coef(const double *z0)
  {
    if (EXT == 4) return;   // the precomputed-coefficient integration reads nothing of this
    if (ORD >= 1) bm_vec(Az, A, z0);
    if (ORD >= 2) {
      BlockMat A2;
      bm_mul(A2, A, A);
      bm_vec(A2z, A2, z0);
      bm_vec(Ab, A, b);
      if (ORD >= 3) {
        BlockMat A3;
        bm_mul(A3, A, A2);
        bm_vec(A3z, A3, z0);
        bm_vec(A2b, A2, b);
        if (ORD >= 4) {
          BlockMat A4;
          bm_mul(A4, A, A3);
          bm_vec(A4z, A4, z0);
          bm_vec(A3b, A3, b);
        }
      }
    }
  }

  // ---- one face: order reduction + solver selection (:1295-1467).  The iterative solves are not executed
  // here: they are described by (deg, q, lambda) so that a kernel can run them from a work queue.
  // kind 0: dtau = 0 (no valid root), 1: dtau = t.tau (closed form), 2: iterative solve pending.
  template <int ORD>
  GB_HD void face_task(const double *c, bool start_face, int i_scaling, SolveTask &t) const
  {
    int solver;
    double qa, qb, qc = 0.0, qd = 0.0, qe = 0.0;
    const bool reduced = start_face || (c[0] == 0.0);
    t.kind = 0;
    t.tau = 0.0;
    t.deg = 0;
    if (ORD == 1) {
      if (reduced) return;
      solver = 1; qa = c[1]; qb = c[0];
      if (qa == 0.0) return;
    } else if (ORD == 2) {
      if (reduced) {
        solver = 1; qa = c[2] / 2.0; qb = c[1];
        if (qa == 0.0) return;
      } else {
        solver = 2; qa = c[2]; qb = c[1]; qc = c[0];
        if (qa == 0.0) {
          if (qb != 0.0) { solver = 1; qa = qb; qb = qc; }
          else return;
        }
      }
    } else if (ORD == 3) {
      if (reduced) {
        solver = 2; qa = c[3] / 3.0; qb = c[2] / 2.0; qc = c[1];
        if (qa == 0.0) {
          if (qb != 0.0) { solver = 1; qa = qb; qb = qc; }
          else return;
        }
      } else {
        solver = 3; qa = c[3]; qb = c[2]; qc = c[1]; qd = c[0];
        if (qa == 0.0) {
          if (qb != 0.0) { solver = 2; qa = qb; qb = qc; qc = qd; }
          else if (qc != 0.0) { solver = 1; qa = qc; qb = qd; }
          else return;
        }
      }
    } else {
      if (reduced) {
        solver = 3; qa = c[4] / 4.0; qb = c[3] / 3.0; qc = c[2] / 2.0; qd = c[1];
        if (qa == 0.0) {
          if (qb != 0.0) { solver = 2; qa = qb; qb = qc; qc = qd; }
          else if (qc != 0.0) { solver = 1; qa = qc; qb = qd; }
          else return;
        }
      } else {
        solver = 4; qa = c[4]; qb = c[3]; qc = c[2]; qd = c[1]; qe = c[0];
        if (qa == 0.0) {
          if (qb != 0.0) { solver = 3; qa = qb; qb = qc; qc = qd; qd = qe; }
          else if (qc != 0.0) { solver = 2; qa = qc; qb = qd; qc = qe; }
          else if (qd != 0.0) { solver = 1; qa = qd; qb = qe; }
          else return;
        }
      }
    }
    if (solver == 1) {
      t.kind = 1;
      t.tau = linear_solver(qa, qb);
    } else if (solver == 2 && i_scaling == 0) {
      t.kind = 1;
      t.tau = quadratic_solver1(qa, qb, qc);
    } else {
      t.kind = 2;
      t.deg = solver;
      if (solver == 2) quadratic2_prepare(qa, qb, qc, t.q, t.lambda);
      else if (solver == 3) cubic_prepare(qa, qb, qc, qd, t.q, t.lambda);
      else quartic_prepare(i_scaling, qa, qb, qc, qd, qe, t.q, t.lambda);
    }
  }
  template <int ORD>
  GB_HD double face_root(const double *c, bool start_face, int i_scaling)
  {
    SolveTask t;
    face_task<ORD>(c, start_face, i_scaling, t);
    if (t.kind == 2) return solve_monic_min_positive(t.deg, t.q[0], t.q[1], t.q[2], t.q[3], t.lambda, solver_iters);
    return t.tau;
  }

  // ---- one face of an order-2 solve with the closed-form solver (i_scaling = 0): operands of the single
  // division that yields dtau; false = no valid root on this face (dtau = 0 or huge in the reference)
  GB_HD bool face_numden_ord2(const double *c, bool start_face, double &num, double &den) const
  {
    const bool reduced = start_face || (c[0] == 0.0);
    // reduced: Linear_Solver(c2/2, c1); degenerate quadratic (c2 == 0): Linear_Solver(c1, c0)
    const bool lin_deg = !reduced && (c[2] == 0.0);
    const double la = reduced ? c[2] / 2.0 : c[1];
    const double lb = reduced ? c[1] : c[0];
    double qn, qd;
    const bool qhas = quadratic_solver1_numden(c[2], c[1], c[0], qn, qd);
    const bool linear = reduced || lin_deg;
    const bool ok = linear ? (la != 0.0) : qhas;
    num = !ok ? 1.0 : linear ? -lb : qn;   // lanes without a root divide 1/1: stays on the fast division path
    den = !ok ? 1.0 : linear ? la : qd;
    return ok;
  }

  // ---- :1258-1482.  dtau/iface untouched when no valid root exists.  All four faces.
  template <int ORD>
  GB_HD bool analytic_approx(unsigned mask, int i_scaling, const double *z, int &iface_inout, double &dtau)
  {
    double cm[4][5];
    coeff<ORD>(mask, z, cm);
    return pick_exit<ORD>(cm, mask, i_scaling, iface_inout, dtau);
  }
  template <int ORD>
  GB_HD bool pick_exit(double cm[4][5], unsigned mask, int i_scaling, int &iface_inout, double &dtau)
  {
    const int iface = iface_inout;
    double best = GB_HUGE;
    int ibest = 0;
    if (ORD == 2 && i_scaling == 0) {
      // closed-form solver: branch free, the four faces are four independent instruction streams
#pragma unroll
      for (int i = 0; i < 4; i++) {
        if (!(mask & (1u << i))) continue;
        double num, den;
        const bool ok = face_numden_ord2(cm[i], (i + 1) == iface, num, den);
        const double d = num / den;
        if (ok && (d < GB_HUGE) && (d > 0.0) && (ibest == 0 || d < best)) {
          best = d;
          ibest = i + 1;
        }
      }
    } else {
      // iterative solvers: ONE call site, executed face after face by the whole warp
#pragma unroll 1
      for (int i = 0; i < 4; i++) {
        if (!(mask & (1u << i))) continue;
        const double d = face_root<ORD>(cm[i], (i + 1) == iface, i_scaling);
        // valid: 0 < d < huge ; minloc keeps the lowest face index on ties
        if ((d < GB_HUGE) && (d > 0.0) && (ibest == 0 || d < best)) {
          best = d;
          ibest = i + 1;
        }
      }
    }
    if (ibest == 0) return false;
    iface_inout = ibest;
    dtau = best;
    return true;
  }
  // Same for a single enabled face (boole_faces with only the guessed face set, :281-298).  The face is lane
  // data: its normal is selected, so that all lanes of a warp run ONE solve together instead of one per face.
  // `prepared`: prepare<ORD>(z) has already been called for this z.
  template <int ORD>
  GB_HD bool analytic_approx_single(int face, int i_scaling, const double *z, int &iface_inout, double &dtau,
                                    bool prepared)
  {
    if (!prepared) prepare<ORD>(z);
    double n[3], c[5];
    face_normal(face, n);
    face_coeffs<ORD>(n, face == 1, z, c, face - 1);
    const double d = face_root<ORD>(c, face == iface_inout, i_scaling);
    if (!((d < GB_HUGE) && (d > 0.0))) return false;
    iface_inout = face;
    dtau = d;
    return true;
  }

  // ---- :2047-2083
  template <int ORD>
  GB_HD void integrate(double *z, double tau)
  {
    if (EXT == 4) {   // analytic_integration dispatches (:2034-2039); the step book-keeping is in the other branch only
      integrate_precomp<ORD>(z, tau);
      return;
    }
    nsteps++;
    if (ADAPT) {
#pragma unroll
      for (int i = 0; i < 4; i++) z0_last[i] = z[i];
    }
    if (LONG) {
      if (nsteps <= lst_cap) {   // the reference's arrays hold 3 * max_n_intermediate_steps entries
        double *e = lst + 5 * (size_t)(nsteps - 1);
        e[0] = tau;
#pragma unroll
        for (int i = 0; i < 4; i++) e[1 + i] = z[i];
      }
    } else if (EXT && EXT != 3) {
      if (nsteps == 1) {
        tau_list[0] = tau;
#pragma unroll
        for (int i = 0; i < 4; i++) z0_list[0][i] = z[i];
      } else if (nsteps == 2) {
        tau_list[1] = tau;
#pragma unroll
        for (int i = 0; i < 4; i++) z0_list[1][i] = z[i];
      }
    }
    if (ORD >= 1) {
#pragma unroll
      for (int i = 0; i < 4; i++) z[i] = z[i] + tau * (b[i] + Az[i]);
    }
    if (ORD >= 2) {
      double tau2_half = tau * tau * 0.5;
#pragma unroll
      for (int i = 0; i < 4; i++) z[i] = z[i] + tau2_half * (Ab[i] + A2z[i]);
    }
    if (ORD >= 3) {
      double tau3_sixth = (tau * tau) * tau / 6.0;
#pragma unroll
      for (int i = 0; i < 4; i++) z[i] = z[i] + tau3_sixth * (A2b[i] + A3z[i]);
    }
    if (ORD >= 4) {
      double t2 = tau * tau;
      double tau4_24 = (t2 * t2) / 24.0;
#pragma unroll
      for (int i = 0; i < 4; i++) z[i] = z[i] + tau4_24 * (A3b[i] + A4z[i]);
    }
  }


  // ==== EXT: Hamiltonian time tracing and optional quantities ===========================================
  // z_series_coef (:2436-2474): Taylor coefficients of x(tau), vpar(tau) from the current Taylor vectors
  GB_HD void z_series(const double *z0, double xc[3][5], double vc[5]) const
  {
#pragma unroll
    for (int i = 0; i < 3; i++) xc[i][0] = z0[i];
    vc[0] = z0[3];
    if (K >= 1) {
#pragma unroll
      for (int i = 0; i < 3; i++) xc[i][1] = b[i] + Az[i];
      vc[1] = b[3] + Az[3];
    }
    if (K >= 2) {
#pragma unroll
      for (int i = 0; i < 3; i++) xc[i][2] = 0.5 * (Ab[i] + A2z[i]);
      vc[2] = 0.5 * (Ab[3] + A2z[3]);
    }
    if (K >= 3) {
#pragma unroll
      for (int i = 0; i < 3; i++) xc[i][3] = 1.0 / 6.0 * (A2b[i] + A3z[i]);
      vc[3] = 1.0 / 6.0 * (A2b[3] + A3z[3]);
    }
    if (K >= 4) {
#pragma unroll
      for (int i = 0; i < 3; i++) xc[i][4] = 1.0 / 24.0 * (A3b[i] + A4z[i]);
      vc[4] = 1.0 / 24.0 * (A3b[3] + A4z[3]);
    }
  }
  // poly_multiplication_coef (:2478-2520): operands of K+1 coefficients, terms above order K dropped; the
  // accumulation order (p1 outer, p2 inner, starting from 0) is the reference's
  GB_HD static void poly_mul(const double *p1, const double *p2, double *res)
  {
#pragma unroll
    for (int i = 0; i <= K; i++) res[i] = 0.0;
#pragma unroll
    for (int j = 0; j <= K; j++)
#pragma unroll
      for (int k = 0; k <= K; k++)
        if (j + k <= K) res[j + k] = res[j + k] + p1[j] * p2[k];
  }
  // moment_integration, scalar version (:3063-3091); x**3, x**4, x**5 as libgcc __powidf2 forms them
  GB_HD static double moment(double tau, const double *c)
  {
    double m = 0.0;
    const double t2 = tau * tau;
    if (K >= 1) m = c[0] * tau + t2 * 0.5 * c[1];
    if (K >= 2) m = m + (tau * t2) / 3.0 * c[2];
    if (K >= 3) m = m + (t2 * t2) / 4.0 * c[3];
    if (K >= 4) m = m + (tau * (t2 * t2)) / 5.0 * c[4];
    return m;
  }
  // hamiltonian_time(ind_tetr): h1_in_curlA, h1_in_curlh, vec_mismatch_der(3), vec_parcurr_der(3) -- two 32-byte sectors
  GB_HD void load_ham(double *h) const
  {
    const double *ph = mp->ham + ((int64_t)ind_tetr - 1) * HAM_ND;
    ld4(ph, h[0], h[1], h[2], h[3]);
    ld4(ph + 4, h[4], h[5], h[6], h[7]);
  }
  // calc_t_hamiltonian (:2214-2256): Hamiltonian time of one integration step (z0, tau)
  GB_HD double ham_delta(const double *z0, double tau, bool recompute)
  {
    if (recompute) set_integration_coef_manually(z0);
    double h[8], xc[3][5], vc[5], xv[3][5], mx[3], mxv[3];
    load_ham(h);
    z_series(z0, xc, vc);
#pragma unroll
    for (int i = 0; i < 3; i++) poly_mul(xc[i], vc, xv[i]);
#pragma unroll
    for (int i = 0; i < 3; i++) {
      mx[i] = moment(tau, xc[i]);
      mxv[i] = moment(tau, xv[i]);
    }
    const double cm = mp->cm_over_e;
    double d = h[0] * tau + cm * h[1] * moment(tau, vc) + dot3(&h[2], mx) + cm * dot3(&h[5], mxv);
    return d * (double)sign_rhs;
  }
  // get_t_hamiltonian_root (:2340-2432): tau at which the Hamiltonian time of the step reaches t_rem (5th order dropped)
  GB_HD double ham_root(const double *z0, double t_rem)
  {
    double h[8], xc[3][5], vc[5], xv[3][5];
    load_ham(h);
    z_series(z0, xc, vc);
#pragma unroll
    for (int i = 0; i < 3; i++) poly_mul(xc[i], vc, xv[i]);
    const double cm = mp->cm_over_e, sg = (double)sign_rhs;
    double co[4] = {0.0, 0.0, 0.0, 0.0};  // e, d, c, b
#pragma unroll
    for (int k = 0; k <= (K < 3 ? K : 3); k++) {
      const double col[3] = {xc[0][k], xc[1][k], xc[2][k]}, colv[3] = {xv[0][k], xv[1][k], xv[2][k]};
      double t = vc[k] * cm * h[1];
      if (k == 0) t = h[0] + t;
      t = t + dot3(col, &h[2]) + cm * dot3(colv, &h[5]);
      co[k] = k == 0 ? t : k == 1 ? 0.5 * t : k == 2 ? (1.0 / 3.0) * t : (1.0 / 4.0) * t;
    }
    if (K >= 1) co[1] = 2.0 * co[1];
    if (K >= 2) co[2] = 6.0 * co[2];
    if (K >= 3) co[3] = 24.0 * co[3];
#pragma unroll
    for (int k = 0; k <= (K < 3 ? K : 3); k++) co[k] = co[k] * sg;
    if (K == 1) return quadratic_solver2(co[1], co[0], -t_rem, solver_iters);
    if (K == 2) return cubic_solver(co[2], co[1], co[0], -t_rem, solver_iters);
    return quartic_solver(0, co[3], co[2], co[1], co[0], -t_rem, solver_iters);
  }
  // calc_optional_quantities (:2134-2210) of one integration step, accumulated into oq[]
  GB_HD void optional_step(const double *z0, double tau, bool recompute)
  {
    if (recompute) set_integration_coef_manually(z0);
    double h[8], xc[3][5], vc[5], xv[3][5], dtc[5], prod[5];
    load_ham(h);
    z_series(z0, xc, vc);
#pragma unroll
    for (int i = 0; i < 3; i++) poly_mul(vc, xc[i], xv[i]);
    const double cm = mp->cm_over_e, sg = (double)sign_rhs;
#pragma unroll
    for (int k = 0; k <= K; k++) {
      const double col[3] = {xc[0][k], xc[1][k], xc[2][k]}, colv[3] = {xv[0][k], xv[1][k], xv[2][k]};
      dtc[k] = dot3(&h[2], col) + cm * h[1] * vc[k] + cm * dot3(&h[5], colv);
    }
    dtc[0] = dtc[0] + h[0];
    if (oq_mask & 1u) oq[0] = oq[0] + moment(tau, dtc) * sg;
    if (oq_mask & 2u) {
      double om[5];
#pragma unroll
      for (int k = 0; k <= K; k++) {
        const double col[3] = {xc[0][k], xc[1][k], xc[2][k]};
        om[k] = 1.0 / cm * dot3(r.gB, col);
      }
      om[0] = om[0] + 1.0 / cm * r.bmod1;
      poly_mul(dtc, om, prod);
      oq[1] = oq[1] - sg * moment(tau, prod);
    }
    if (oq_mask & 4u) {
      poly_mul(dtc, vc, prod);
      oq[2] = oq[2] + sg * moment(tau, prod);
    }
    if (oq_mask & 8u) {
      double v2[5];
      poly_mul(vc, vc, v2);
      poly_mul(dtc, v2, prod);
      oq[3] = oq[3] + sg * moment(tau, prod);
    }
  }
  // the loop over number_of_integration_steps at the end of the pusher (:662-667)
  GB_HD void optional_all()
  {
    if (!OPT) return;
    if (!oq_mask) return;
    if (LONG) {
      const int nst = nsteps < lst_cap ? nsteps : lst_cap;
      for (int i = 0; i < nst; i++) {
        const double *e = lst + 5 * (size_t)i;
        const double z0[4] = {e[1], e[2], e[3], e[4]};
        optional_step(z0, e[0], nsteps > 1);
      }
      return;
    }
    if (nsteps >= 1) optional_step(z0_list[0], tau_list[0], nsteps > 1);
    if (nsteps >= 2) optional_step(z0_list[1], tau_list[1], true);
  }
  // Hamiltonian time summed over the integration steps of the push (:470-486, 621-631); thl = t_hamiltonian_list(2:3)
  // EXT = 5: the same over the long lists; i_root = the first step after which |t_hamiltonian| > |t_lim| (0: none),
  // th_before = t_hamiltonian_list(i_root) = the sum before that step, th_before_last = the sum before the last step
  GB_HD double ham_total_long(double t_lim, int &i_root, double &th_before, double &th_before_last)
  {
    double th = 0.0;
    i_root = 0;
    th_before = th_before_last = 0.0;
    const int nst = nsteps < lst_cap ? nsteps : lst_cap;
    for (int i = 0; i < nst; i++) {
      const double *e = lst + 5 * (size_t)i;
      const double z0[4] = {e[1], e[2], e[3], e[4]};
      const double prev = th;
      th = th + ham_delta(z0, e[0], nsteps > 1);
      th_before_last = prev;
      if (i_root == 0 && fabs(th) > fabs(t_lim)) {
        i_root = i + 1;
        th_before = prev;
      }
    }
    return th;
  }
  GB_HD double ham_total(double *thl)
  {
    double th = 0.0;
    thl[0] = thl[1] = 0.0;
    if (nsteps >= 1) {
      th = th + ham_delta(z0_list[0], tau_list[0], nsteps > 1);
      thl[0] = th;
    }
    if (nsteps >= 2) {
      th = th + ham_delta(z0_list[1], tau_list[1], true);
      thl[1] = th;
    }
    return th;
  }


  // ==== EXT = 2: orbit events -- J_par / banana tips / toroidal mappings =================================
  // module par_adiab_inv_poly_mod (:3156-3429) and the event part of gorilla_plot_orbit_integration
  // (SRC/gorilla_plot_mod.f90:585-638); events go to a device buffer instead of files.
  // par_adiab_tau (:3295-3322); a**3, a**4, x**5 as libgcc __powidf2 forms them
  GB_HD static double par_adiab_tau(double a44, double b4, double tau, double v)
  {
    const double t2 = tau * tau, v2 = v * v, a2 = a44 * a44, bb = b4 * b4;
    const double a3 = a44 * a2, t3 = tau * t2, t4 = t2 * t2;
    double r = tau * v2 + 0.5 * t2 * (2.0 * b4 * v + 2.0 * a44 * v2);
    if (K == 4) r = r + 1.0 / 3.0 * t3 * (bb + 3.0 * a44 * b4 * v + 2.0 * a44 * 2.0 * v2);  // (sic) :3316
    else r = r + 1.0 / 3.0 * t3 * (bb + 3.0 * a44 * b4 * v + 2.0 * a2 * v2);
    if (K >= 3) r = r + 1.0 / 4.0 * t4 * (a44 * bb + 7.0 / 3.0 * a2 * b4 * v + (4.0 * a3 * v2) / 3.0);
    if (K >= 4) {
      const double a4 = a2 * a2, t5 = tau * t4;
      r = r + 1.0 / 60.0 * t5 * (7.0 * a2 * bb + 15.0 * a3 * b4 * v + 8.0 * a4 * v2);
    }
    return r;
  }
  // tau_vpar_root (:3326-3374)
  GB_HD double tau_vpar_root(double a44, double b4, double v)
  {
    const double a2 = a44 * a44, a3 = a44 * a2, a4 = a2 * a2;
    const double c1 = v, c2 = b4 + a44 * v, c3 = a44 * b4 + a2 * v;
    if (K <= 2) return quadratic_solver2(c3, c2, c1, solver_iters);
    const double c4 = a2 * b4 + a3 * v;
    if (K == 3) return cubic_solver(c4, c3, c2, c1, solver_iters);
    const double c5 = a3 * b4 + a4 * v;
    return quartic_solver(0, c5, c4, c3, c2, c1, solver_iters);
  }
  // energy_tot_func / p_phi_func (SRC/supporting_functions_mod.f90:279-301, 377-408) in the current tetrahedron
  GB_HD double energy_tot(const double *z /*[4]*/) const
  {
    const double vperp = sqrt(2.0 * fabs(perpinv) * (r.bmod1 + dot3(r.gB, z)));
    double phi = 0.0;
    if (PHI) phi = r.Phi1 + dot3(r.gPhi, z);
    double e = mp->particle_mass / 2.0 * (vperp * vperp + z[3] * z[3]) + mp->particle_charge * phi;
    if (PHI == 2) e = e + 0.5 * mp->particle_mass * (r.v2Emod1 + dot3(z, r.gv2Emod));
    return e;
  }
  GB_HD double p_phi(double vpar, const double *z /*[3]*/) const
  {
    const double *pc = mp->cold + ((int64_t)ind_tetr - 1) * COLD_ND;
    const double gh[3] = {ldg(pc + C_GHPHI), ldg(pc + C_GHPHI + 1), ldg(pc + C_GHPHI + 2)};
    const double gA[3] = {ldg(pc + C_GAPHI), ldg(pc + C_GAPHI + 1), ldg(pc + C_GAPHI + 2)};
    double p = mp->particle_mass * vpar * (ldg(pc + C_HPHI1) + dot3(gh, z)) +
               mp->particle_mass / mp->cm_over_e * (ldg(pc + C_APHI1) + dot3(gA, z));
    if (PHI == 2) {
      const double *ps = mp->se + ((int64_t)ind_tetr - 1) * SE_ND;
      const double gv[3] = {ldg(ps + S_GVE2), ldg(ps + S_GVE2 + 1), ldg(ps + S_GVE2 + 2)};
      p = p + mp->particle_mass * (ldg(ps + S_VE2_1) + dot3(z, gv));
    }
    return p;
  }
  // what gorilla_plot_orbit_integration does after a push that did not end the time step (:585-638)
  GB_HD void events_after_push(double vpar_in, const PushOut &o, EvState &es)
  {
    es.n = 0;
    if (LONG && (es.flags & 6) && !removed) {   // par_adiab_inv_tetra_poly (:3173-3291) over the long lists
      const double a44 = A.s, b4 = b[3], vpar_end = o.vpar;
      const int nst = nsteps < lst_cap ? nsteps : lst_cap;
      if ((vpar_end > 0.0) && (vpar_in < 0.0)) {
        int turning_index = -1;   // findloc(intermediate_z0_list(4, 1:n) > 0) - 1
        for (int i = 1; i <= nst; i++)
          if (lst[5 * (size_t)(i - 1) + 4] > 0.0) {
            turning_index = i - 1;
            break;
          }
        if (turning_index != 0) {   // 0: reference stops with an error (cannot happen: z0(4,1) = vpar_in < 0)
          if (turning_index == -1) turning_index = nst;
          const double *et = lst + 5 * (size_t)(turning_index - 1);
          const double tau_part1 = tau_vpar_root(a44, b4, et[4]);
          for (int i = 1; i <= turning_index - 1; i++) {
            const double *e = lst + 5 * (size_t)(i - 1);
            es.J = es.J + par_adiab_tau(a44, b4, e[0], e[4]) * dt_dtau_const;
          }
          es.J = es.J + par_adiab_tau(a44, b4, tau_part1, et[4]) * dt_dtau_const;
          if (es.cnt_v > 1 && (es.cnt_v / es.nskip_v * es.nskip_v == es.cnt_v)) {
            double z[4] = {et[1], et[2], et[3], et[4]};
            set_integration_coef_manually(z);
            const int keep = nsteps;
            const double k0 = lst[0], k1z = lst[1], k2z = lst[2], k3z = lst[3], k4z = lst[4];
            nsteps = 0;                   // analytic_integration_external (:3378-3408) has no step book-keeping:
            integrate<K>(z, tau_part1);   // entry 1 of the list is overwritten by integrate() and restored
            lst[0] = k0; lst[1] = k1z; lst[2] = k2z; lst[3] = k3z; lst[4] = k4z;
            nsteps = keep;
            EvRec &e = es.e[es.n++];
            e.kind = 2;
            e.counter = es.cnt_v;
#pragma unroll
            for (int i = 0; i < 3; i++) e.x[i] = z[i] + r.x1s(i);
            e.v[0] = es.J;
            e.v[1] = energy_tot(z);
          }
          es.cnt_v = es.cnt_v + 1;
          es.J = 0.0;
          es.J = es.J + par_adiab_tau(a44, b4, et[0] - tau_part1, 0.0) * dt_dtau_const;
          for (int i = turning_index + 1; i <= nst; i++) {
            const double *e = lst + 5 * (size_t)(i - 1);
            es.J = es.J + par_adiab_tau(a44, b4, e[0], e[4]) * dt_dtau_const;
          }
        }
      } else {
        for (int i = 1; i <= nst; i++) {
          const double *e = lst + 5 * (size_t)(i - 1);
          es.J = es.J + par_adiab_tau(a44, b4, e[0], e[4]) * dt_dtau_const;
        }
      }
    } else if ((es.flags & 6) && !removed) {   // par_adiab_inv_tetra_poly (:3173-3291)
      const double a44 = A.s, b4 = b[3], vpar_end = o.vpar;
      const double v1 = z0_list[0][3], v2 = z0_list[1][3], tau1 = tau_list[0], tau2 = tau_list[1];
      if ((vpar_end > 0.0) && (vpar_in < 0.0)) {
        // turning_index = findloc(z0(4,1:nsteps) > 0) - 1, "not found" -> nsteps (index 0 cannot happen: z0(4,1) = vpar_in < 0)
        const bool turn2 = (nsteps >= 2) && !(v2 > 0.0);   // the bounce lies in the second integration step
        const double v_turn = turn2 ? v2 : v1, tau_turn = turn2 ? tau2 : tau1;
        const double tau_part1 = tau_vpar_root(a44, b4, v_turn);
        if (turn2) es.J = es.J + par_adiab_tau(a44, b4, tau1, v1) * dt_dtau_const;
        es.J = es.J + par_adiab_tau(a44, b4, tau_part1, v_turn) * dt_dtau_const;
        if (es.cnt_v > 1 && (es.cnt_v / es.nskip_v * es.nskip_v == es.cnt_v)) {
          double z[4];
#pragma unroll
          for (int i = 0; i < 4; i++) z[i] = turn2 ? z0_list[1][i] : z0_list[0][i];
          set_integration_coef_manually(z);
          // analytic_integration_external (:3378-3408) = the arithmetic of integrate(); the step lists are not needed
          // any more (v1, v2, tau1, tau2 were read above), so its book-keeping is harmless
          const int keep = nsteps;
          integrate<K>(z, tau_part1);
          nsteps = keep;
          EvRec &e = es.e[es.n++];
          e.kind = 2;
          e.counter = es.cnt_v;
#pragma unroll
          for (int i = 0; i < 3; i++) e.x[i] = z[i] + r.x1s(i);
          e.v[0] = es.J;
          e.v[1] = energy_tot(z);
        }
        es.cnt_v = es.cnt_v + 1;
        es.J = 0.0;
        es.J = es.J + par_adiab_tau(a44, b4, tau_turn - tau_part1, 0.0) * dt_dtau_const;
        if (!turn2 && nsteps >= 2) es.J = es.J + par_adiab_tau(a44, b4, tau2, v2) * dt_dtau_const;
      } else {
        if (nsteps >= 1) es.J = es.J + par_adiab_tau(a44, b4, tau1, v1) * dt_dtau_const;
        if (nsteps >= 2) es.J = es.J + par_adiab_tau(a44, b4, tau2, v2) * dt_dtau_const;
      }
    }
    if (iper_phi != 0) {   // toroidal mappings (:601-636)
      es.cnt_p = es.cnt_p + iper_phi;
      if ((es.flags & 1) && (es.cnt_p / es.nskip_p * es.nskip_p == es.cnt_p)) {
        const double zv[4] = {o.z_save[0], o.z_save[1], o.z_save[2], o.vpar};
        EvRec &e = es.e[es.n++];
        e.kind = 1;
        e.counter = es.cnt_p;
#pragma unroll
        for (int i = 0; i < 3; i++) e.x[i] = o.x[i];
        e.v[0] = p_phi(o.vpar, o.z_save);
        e.v[1] = energy_tot(zv);
      }
    }
  }


  // ==== EXT = 3: adaptive energy-controlled sub-stepping (boole_adaptive_time_steps, :830-1254) =========
  // Only the last entry of intermediate_z0_list is read by the scheme (z0_last); the list consumers (Hamiltonian time,
  // optional quantities, J_par) are not combined with it.  Where the reference reads uninitialised locals:
  // eta_minimum / tau_minimum start as the unsplit step, iface_out keeps the incoming face if no exit time was solved.
  // adaptive_time_steps_update_eta (:1167-1213); 1E-15 there is a default-real literal
  GB_HD void update_eta(int ORD, double delta, int &eta) const
  {
    const double min_step_error = (double)1E-15f;
    const int max_n = mp->max_n_intermediate_steps;
    double sf = pow(delta / mp->desired_delta_energy, 1.0 / ORD);
    const double msf = pow(delta / (min_step_error * eta), 1.0 / (ORD + 1));
    if ((sf > 1.0) && (msf > 1.0)) {
      sf = sf < msf ? sf : msf;
      const int c = (int)ceil(eta * sf);
      eta = c < max_n ? c : max_n;
    } else if (sf > 1.0) {
      const int c = (int)ceil(eta * sf);
      eta = c < max_n ? c : max_n;
    } else {
      eta = (int)(eta + 1.0);
    }
  }
  // adaptive_time_steps_exit_time (:1217-1252)
  template <int ORD>
  GB_HD bool adaptive_exit_time(int i_scaling, const double *z, bool bga, int &iface, double &tau_exit)
  {
    if (bga && ORD > 2) {
      const bool ok2 = analytic_approx<2>(0xFu, i_scaling, z, iface, tau_exit);
      const int guess = iface;
      iface = 0;
      if (ok2) return analytic_approx_single<ORD>(guess, i_scaling, z, iface, tau_exit, false);
    }
    return analytic_approx<ORD>(0xFu, i_scaling, z, iface, tau_exit);
  }
  GB_HD bool any_outside(const double *z) const
  {
    double d[4];
    normal_distances(z, d);
    return (d[0] < 0.0) || (d[1] < 0.0) || (d[2] < 0.0) || (d[3] < 0.0);
  }
  // adaptive_time_steps_equidistant (:937-1163)
  template <int ORD>
  GB_HD void adaptive_equidistant(int i_scaling, bool bga, bool bp, double delta, int &iface_out, double &tau, double *z,
                                  bool &bfc)
  {
    const int max_n = mp->max_n_intermediate_steps;
    double z_start[4];
#pragma unroll
    for (int i = 0; i < 4; i++) z_start[i] = z0_last[i];
    const int nsteps_start = nsteps - 1;
    const double energy_start = energy_tot(z_start);
    int eta = 1, eta_extended = 1, eta_minimum = 1;
    bool reached_minimum = false;
    double delta_minimum = delta, tau_minimum = tau, tau_exit = 0.0;
    int iface_new = iface_out;
    n_adaptive++;
    fallback |= 16;
    // (the partition loop of the reference has no iteration bound and can cycle when two partitions give exactly the same
    // energy error; it is ended after max_n + 64 partitions here and in the oracle)
    int n_partitions = 0;
    while (eta < max_n) {
      if (++n_partitions > max_n + 64) break;
      update_eta(ORD, delta, eta);
      if (reached_minimum) {
        eta = eta_minimum;
        tau = tau_minimum;
      }
#pragma unroll
      for (int i = 0; i < 4; i++) z[i] = z_start[i];
      nsteps = nsteps_start;
      double tau_prime = tau / eta, tau_collected = 0.0;
      bfc = true;
      bool exit_tetra = false;
      const int eta_limit = bp ? (int)ceil(max_n * 1.1) : eta;
      for (int i = 1; i <= eta_limit - 1; i++) {
        set_coef<ORD>(z);
        integrate<ORD>(z, tau_prime);
        if (any_outside(z)) {
          if (i == 1) {
            bfc = false;
            return;
          }
          // step back to the start of this sub-step (its z0 is the last list entry)
#pragma unroll
          for (int q = 0; q < 4; q++) z[q] = z0_last[q];
          nsteps = nsteps - 1;
          exit_tetra = true;
          break;
        }
        tau_collected = tau_collected + tau_prime;
        eta_extended = i;
        if (!bga) {
          iface_new = 0;
          if (!adaptive_exit_time<ORD>(i_scaling, z, false, iface_new, tau_exit)) {
            bfc = false;
            return;
          }
          if (tau_exit <= tau_prime) {
            tau_prime = tau_exit;
            break;
          }
        }
      }
      if (exit_tetra) {
        iface_new = 0;
        if (!adaptive_exit_time<ORD>(i_scaling, z, bga, iface_new, tau_exit)) {
          bfc = false;
          return;
        }
        tau_prime = tau_exit;
      } else {
        set_coef<ORD>(z);
      }
      integrate<ORD>(z, tau_prime);
      tau_collected = tau_collected + tau_prime;
      delta = fabs(1 - energy_tot(z) / energy_start);
      const int eta_buffer = eta;
      const double tau_buffer = tau;
      eta = eta_extended + 1;
      tau = tau_collected;
      if (reached_minimum) {
        break;
      } else if (delta < delta_minimum) {
        delta_minimum = delta;
        eta_minimum = eta_buffer;
        tau_minimum = tau_buffer;
        if (delta <= mp->desired_delta_energy) break;
      } else if (delta > delta_minimum) {
        reached_minimum = true;
      }
    }
    iface_out = iface_new;
  }
  // the energy test of overhead_adaptive_time_steps (:896-901) for the step that was just integrated from z0_last to z
  GB_HD double adaptive_delta_energy(const double *z) const
  {
    const double energy_start = energy_tot(z0_last), energy_current = energy_tot(z);
    return fabs(1 - energy_current / energy_start);
  }
  // overhead_adaptive_time_steps (:830-933)
  template <int ORD>
  GB_HD void overhead_adaptive(int i_scaling, bool bga, bool bp, int &iface, double &tau, double *z, bool &bfc)
  {
    if (bp) {
      if (!three_planes_ok(z, iface)) bfc = false;
      if (!face_converged(z, iface)) bfc = false;
    } else {
      // check_three_planes(z, 0): faces modulo(0 + j - 1, 4) + 1 = 1, 2, 3 -- the fourth face is not looked at (:689-691)
      double d[4];
      normal_distances(z, d);
      if (d[0] < 0.0 || d[1] < 0.0 || d[2] < 0.0) bfc = false;
    }
    if (!bfc) return;
    const double delta = adaptive_delta_energy(z);
    if (delta > mp->desired_delta_energy) adaptive_equidistant<ORD>(i_scaling, bga, bp, delta, iface, tau, z, bfc);
  }

  // all four normal distances (:2690-2705); static indexing keeps r.an in registers
  GB_HD void normal_distances(const double *z, double *d) const
  {
#pragma unroll
    for (int f = 0; f < 4; f++) d[f] = dot3(z, r.an[f]);
    d[0] = d[0] + r.dist_ref;
  }
  GB_HD double normal_distance(const double *z, int iface /*1-based*/) const
  {
    double d[4];
    normal_distances(z, d);
    return iface == 1 ? d[0] : iface == 2 ? d[1] : iface == 3 ? d[2] : d[3];
  }
  // anorm(:,iface) by selection instead of a dynamically indexed register array
  GB_HD void face_normal(int iface, double *n) const
  {
#pragma unroll
    for (int i = 0; i < 3; i++)
      n[i] = iface == 1 ? r.an[0][i] : iface == 2 ? r.an[1][i] : iface == 3 ? r.an[2][i] : r.an[3][i];
  }
  // ---- :2741-2775
  GB_HD double normal_v_from_trajectory(int iface, double tau) const
  {
    double n[3];
    face_normal(iface, n);
    double v = 0.0;
    if (K >= 1) v = (n[0] * (b[0] + Az[0]) + n[1] * (b[1] + Az[1])) + n[2] * (b[2] + Az[2]);
    if (K >= 2) v = v + ((n[0] * tau * (Ab[0] + A2z[0]) + n[1] * tau * (Ab[1] + A2z[1])) + n[2] * tau * (Ab[2] + A2z[2]));
    if (K >= 3) {
      double h = tau * tau * 0.5;
      v = v + ((n[0] * h * (A2b[0] + A3z[0]) + n[1] * h * (A2b[1] + A3z[1])) + n[2] * h * (A2b[2] + A3z[2]));
    }
    if (K >= 4) {
      double s = (tau * tau) * tau / 6.0;
      v = v + ((n[0] * s * (A3b[0] + A4z[0]) + n[1] * s * (A3b[1] + A4z[1])) + n[2] * s * (A3b[2] + A4z[2]));
    }
    return v;
  }
  // ---- :2709-2739, poly1 quantities (n.alpha, n.beta, n.curlA) formed on the fly
  GB_HD double normal_velocity(const double *z, int iface) const
  {
    if (EXT == 4) return normal_velocity_precomp(z, iface);
    double n[3];
    face_normal(iface, n);
    const double pc = perpinv * mp->cm_over_e;
    double t[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
      double in_alp = dot3(n, &r.alp[3 * j]);
      if (PHI) {
        double in_bet = dot3(n, &r.bet[3 * j]);
        t[j] = (-GB_CLIGHT * in_bet + pc * in_alp) * z[j];
      } else {
        t[j] = (pc * in_alp) * z[j];
      }
    }
    double in_betvec = dot3(n, r.curlA);
    double v = ((t[0] + t[1]) + t[2] + in_betvec * z[3]) * (double)sign_rhs + dot3(n, b);
    if (PHI == 2) {  // :2728-2732
      double g[3];
#pragma unroll
      for (int j = 0; j < 3; j++) g[j] = -0.5 * mp->cm_over_e * dot3(n, &r.gam[3 * j]) * z[j];
      const double in_gamvec = dot3(n, r.curlvE);
      v = v + ((g[0] + g[1]) + g[2] + mp->cm_over_e * in_gamvec * z[3]) * (double)sign_rhs;
    }
    return v;
  }
  // check_three_planes (:679-699): the three faces other than the exit face must have distance >= 0
  GB_HD bool three_planes_ok(const double *z, int iface_new) const
  {
    double d[4];
    normal_distances(z, d);
    bool ok = true;
#pragma unroll
    for (int f = 0; f < 4; f++)
      if (f + 1 != iface_new && d[f] < 0.0) ok = false;
    return ok;
  }
  // check_face_convergence (:703-718)
  GB_HD bool face_converged(const double *z, int iface_new) const
  {
    return !(fabs(normal_distance(z, iface_new)) > 1.e-11);
  }
  // both of the above from one evaluation of the four distances
  GB_HD bool exit_point_ok(const double *z, int iface_new) const
  {
    double d[4];
    normal_distances(z, d);
    bool ok = true;
#pragma unroll
    for (int f = 0; f < 4; f++) {
      if (f + 1 != iface_new) {
        if (d[f] < 0.0) ok = false;
      } else {
        if (fabs(d[f]) > 1.e-11) ok = false;
      }
    }
    return ok;
  }
  // ---- :2779-2831
  GB_HD double physical_estimate_tau() const
  {
    const double *cold = mp->cold + ((int64_t)ind_tetr - 1) * COLD_ND;
    double tetra_dist_ref = fabs(ldg(cold + C_TETRA_DIST_REF));
    double R1 = ldg(cold + C_R1), Er_mod = ldg(cold + C_ER_MOD);
    double vperp2 = -2.0 * perpinv * bmod0;
    double vd_ExB = (PHI == 2) ? r.vE_mod_avg : fabs(GB_CLIGHT / bmod0 * Er_mod);
    double tau_est = fabs(tetra_dist_ref / z_init[3]);
    if (!(vperp2 == 0.0)) {
      double c2 = sqrt(tetra_dist_ref * vmod0 * R1 / (vperp2 * (double)mp->grid_size2 * 0.1));
      if (c2 < tau_est) tau_est = c2;
    }
    if (!(vd_ExB == 0.0)) {
      double c3 = tetra_dist_ref / vd_ExB;
      if (c3 < tau_est) tau_est = c3;
    }
    return fabs(tau_est / dt_dtau_const);
  }

  // ---- :762-826 ; returns boole_analytical_approx
  GB_HD bool prolonged_trajectory(int i_scaling, double *z, double &tau, int &iface_new, bool &face_correct)
  {
    double tau_save = tau, tau_max = 0.0;
    int iface_new_save = iface_new;
    bool approx = true;
    if (K > 2) {
      approx = analytic_approx<2>(0xFu, i_scaling, z, iface_new, tau);
      tau_max = tau * GB_EPS_TAU;
    }
    iface_new = iface_new_save;
    approx = analytic_approx<K>(0xFu, i_scaling, z, iface_new, tau);
    if (!approx) return false;
    integrate<K>(z, tau);
    if (ADAPT) overhead_adaptive<K>(i_scaling, false, true, iface_new, tau, z, face_correct);  // :811-814
    if (K > 2 && tau > tau_max) face_correct = false;
    if (!three_planes_ok(z, iface_new)) face_correct = false;
    if (normal_velocity(z, iface_new) > 0.0) face_correct = false;
    if (!face_converged(z, iface_new)) face_correct = false;
    tau = tau + tau_save;
    fallback |= 4;   // counted when the second segment was actually integrated
    return true;
  }

  GB_HD bool ts_checks(const double *z, int iface_new, double tau, double tau_max) const
  {
    bool ok = true;
    if (!three_planes_ok(z, iface_new)) ok = false;
    if (!face_converged(z, iface_new)) ok = false;
    if (normal_velocity(z, iface_new) > 0.0) ok = false;
    if (tau > tau_max) {
      double tau_max_est = physical_estimate_tau() * GB_EPS_TAU;
      if (tau > tau_max_est) ok = false;
    }
    return ok;
  }
  // ---- :2835-2998 ; returns boole_trouble_shooting
  GB_HD bool trouble_shooting(double *z, double &tau, int &iface_new)
  {
    fallback |= 2;
    bool face_correct = false, approx;
    iface_new = iface_init;
#pragma unroll
    for (int i = 0; i < 4; i++) z[i] = z_init[i];
    approx = analytic_approx<2>(0xFu, 0, z, iface_new, tau);
    (void)approx;
    double tau_max = tau * GB_EPS_TAU;
    if (K == 4) {
      for (int i = 1; i <= 6 && !face_correct; i++) {
        iface_new = iface_init;
#pragma unroll
        for (int q = 0; q < 4; q++) z[q] = z_init[q];
        nsteps = 0;
        if (!analytic_approx<K>(0xFu, i, z, iface_new, tau)) return false;
        integrate<K>(z, tau);
        face_correct = ts_checks(z, iface_new, tau, tau_max);
      }
    }
    if (!face_correct) {
      iface_new = iface_init;
#pragma unroll
      for (int q = 0; q < 4; q++) z[q] = z_init[q];
      nsteps = 0;
      if (K == 4) {
        // order reduced to 3, i_scaling 0; trajectory integrated with the ORIGINAL order (:2961),
        // using the order-4 terms left over from the loop above (same z_init => same values)
        if (!analytic_approx<3>(0xFu, 0, z, iface_new, tau)) return false;
      } else if (K == 1) {
        // no case(1) in the reference (:2936-2947); keep order, i_scaling = 0
        if (!analytic_approx<K>(0xFu, 0, z, iface_new, tau)) return false;
      } else {
        if (!analytic_approx<K>(0xFu, 1, z, iface_new, tau)) return false;
      }
      integrate<K>(z, tau);
      bool fc = true;
      // :2966-2971, with the REDUCED order and its i_scaling.  (The call inside the loop above, :2897-2902, is made with
      // boole_face_correct = .false. and returns at once, :893.)
      if (ADAPT) overhead_adaptive<(K == 4 ? 3 : K)>((K == 2 || K == 3) ? 1 : 0, false, true, iface_new, tau, z, fc);
      if (!fc || !ts_checks(z, iface_new, tau, tau_max)) return false;
    }
    return true;
  }

  // ---- pusher_handover2neighbour (kind 1)
  GB_HD void handover(int iface_exit, double *x, int32_t &ind_out, int32_t &iface_out) const
  {
    const int f = iface_exit - 1;
    const uint32_t flags = r.flags();
    ind_out = r.nb(f);
    iface_out = topo_face(flags, f);
    const int iper_phi = topo_perphi(flags, f);
    if (OPT) const_cast<PolyPusher *>(this)->iper_phi = iper_phi;
    if (OPT && mp->skew != nullptr) {
      // handover_processing_kind = 2: position exchange via Cartesian skew coordinates (pusher_tetra_func_mod.f90:59-89).
      // A particle leaving the domain keeps its exit position (the reference indexes tetra_skew_coord(-1) there).
      if (ind_out < 1) return;
      double blk[24], bv[3], tv[3], xc[3];
      const double *pl = mp->skew + ((int64_t)ind_tetr - 1) * SKEW_ND + 48 * f;
#pragma unroll
      for (int i = 0; i < 24; i += 4) ld4(pl + i, blk[i], blk[i + 1], blk[i + 2], blk[i + 3]);
#pragma unroll
      for (int i = 0; i < 3; i++) bv[i] = x[i] - blk[i];
#pragma unroll
      for (int i = 0; i < 3; i++) tv[i] = ((0.0 + blk[3 + i] * bv[0]) + blk[6 + i] * bv[1]) + blk[9 + i] * bv[2];
#pragma unroll
      for (int i = 0; i < 3; i++) xc[i] = (((0.0 + blk[12 + i] * tv[0]) + blk[15 + i] * tv[1]) + blk[18 + i] * tv[2]) + blk[21 + i];
      const double *pe = mp->skew + ((int64_t)ind_out - 1) * SKEW_ND + 48 * (iface_out - 1) + 24;
#pragma unroll
      for (int i = 0; i < 24; i += 4) ld4(pe + i, blk[i], blk[i + 1], blk[i + 2], blk[i + 3]);
#pragma unroll
      for (int i = 0; i < 3; i++) bv[i] = xc[i] - blk[i];
#pragma unroll
      for (int i = 0; i < 3; i++) tv[i] = ((0.0 + blk[3 + i] * bv[0]) + blk[6 + i] * bv[1]) + blk[9 + i] * bv[2];
#pragma unroll
      for (int i = 0; i < 3; i++) x[i] = (((0.0 + blk[12 + i] * tv[0]) + blk[15 + i] * tv[1]) + blk[18 + i] * tv[2]) + blk[21 + i];
      return;
    }
    if (mp->coord_system == 1) {
      if (iper_phi == 1) x[1] = x[1] - mp->period_phi;
      else if (iper_phi == -1) x[1] = x[1] + mp->period_phi;
    } else {
      const int iper_theta = topo_pertheta(flags, f);
      if (iper_phi == 1) x[2] = x[2] - mp->period_phi;
      else if (iper_phi == -1) x[2] = x[2] + mp->period_phi;
      if (iper_theta == 1) x[1] = x[1] - mp->period_theta;
      else if (iper_theta == -1) x[1] = x[1] + mp->period_theta;
    }
  }

  GB_HD void set_removed(PushOut &o) const
  {
    if (OPT) const_cast<PolyPusher *>(this)->removed = true;
    o.ind_tetr = -1;
    o.iface = -1;
    o.finished = 0;
    o.z_save_set = 0;
    o.fallback = fallback;
  }

  // ---- final processing shared by push_fast / push_full (:459-487, 497-667).
  // Returns false if (fast mode) the stop-inside case needs the fall-back ladder.
  template <bool FAST>
  GB_HD bool finish(double *z, double tau, int iface_new, PushOut &o)
  {
    // x, vpar are updated BEFORE the stop-inside test (:459-460); a removal further down keeps them
#pragma unroll
    for (int i = 0; i < 3; i++) o.x[i] = z[i] + r.x1s(i);
    o.vpar = z[3];
    const bool tt2 = (EXT == 1) || (OPT && mp->time_tracing == 2);
    double thl[2] = {0.0, 0.0};
    int i_root = 0;                              // EXT = 5 (long lists)
    double th_before = 0.0, th_before_last = 0.0;
    double t_pass;
    if (tt2) t_pass = LONG ? ham_total_long(t_remain, i_root, th_before, th_before_last) : ham_total(thl);
    else t_pass = tau * dt_dtau_const;
    o.t_pass = t_pass; // (:466) assigned before the stop-inside test; kept if the particle is removed below
    if (fabs(t_pass) >= fabs(t_remain)) {
      if (FAST && (tt2 || nsteps > 1)) return false;
      if (!tt2) {
#pragma unroll
        for (int i = 0; i < 4; i++) z[i] = z_init[i];
        if (nsteps > 1) {
          iface_new = iface_init;
          set_integration_coef_manually(z);
        }
        nsteps = 0;
        tau = t_remain / dt_dtau_const;
      } else if (LONG) {
        // :523-541 over the long lists
        const int nst = nsteps < lst_cap ? nsteps : lst_cap;
        if (i_root == 0) {
          i_root = nst;
          th_before = th_before_last;
        }
        const double *e = lst + 5 * (size_t)(i_root - 1);
#pragma unroll
        for (int i = 0; i < 4; i++) z[i] = e[1 + i];
        set_integration_coef_manually(z);
        const double tau_step = e[0];
        const double t_remain_new = t_remain - th_before;
        nsteps = i_root - 1;
        iface_new = iface_init;
        tau = ham_root(z, t_remain_new);
        if (tau > tau_step) tau = t_remain_new / dt_dtau_const;
      } else {
        // :523-541  step in which the Hamiltonian time exceeds t_remain (findloc over t_hamiltonian_list; an exact tie,
        // which the reference does not handle, takes the last step)
        const bool second = (nsteps > 1) && !(fabs(thl[0]) > fabs(t_remain));
#pragma unroll
        for (int i = 0; i < 4; i++) z[i] = second ? z0_list[1][i] : z0_list[0][i];
        set_integration_coef_manually(z);
        const double tau_step = second ? tau_list[1] : tau_list[0];
        const double t_remain_new = t_remain - (second ? thl[0] : 0.0);
        nsteps = second ? 1 : 0;
        iface_new = iface_init;
        tau = ham_root(z, t_remain_new);
        if (tau > tau_step) tau = t_remain_new / dt_dtau_const;
      }
      integrate<K>(z, tau);
      if (ADAPT) {  // :564-568
        if (FAST) {
          double d[4];
          normal_distances(z, d);
          if (!(d[0] < 0.0 || d[1] < 0.0 || d[2] < 0.0) && adaptive_delta_energy(z) > mp->desired_delta_energy) return false;
        } else {
          bool fc = main_fc;
          overhead_adaptive<K>(0, false, false, iface_new, tau, z, fc);
        }
      }
      bool inside = true;
      {
        double d[4];
        normal_distances(z, d);
#pragma unroll
        for (int f = 0; f < 4; f++)
          if (d[f] < 0.0) inside = false;
      }
      if (inside) {
        o.ind_tetr = ind_tetr;
        o.iface = 0;
        o.finished = 1;
#pragma unroll
        for (int i = 0; i < 3; i++) {
          o.z_save[i] = z[i];
          o.x[i] = z[i] + r.x1s(i);
        }
        o.z_save_set = 1;
        o.vpar = z[3];
        o.t_pass = t_remain;
        o.fallback = fallback;
        optional_all();
        return true;
      }
      if (FAST) return false;
      fallback |= 8;
      if (!trouble_shooting(z, tau, iface_new)) {
        set_removed(o);
        return true;
      }
      if (tt2) t_pass = LONG ? ham_total_long(t_remain, i_root, th_before, th_before_last) : ham_total(thl);
      else t_pass = tau * dt_dtau_const;
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
      o.z_save[i] = z[i];
      o.x[i] = z[i] + r.x1s(i);
    }
    o.z_save_set = 1;
    o.vpar = z[3];
    o.t_pass = t_pass;
    o.finished = 0;
    handover(iface_new, o.x, o.ind_tetr, o.iface);
    o.fallback = fallback;
    optional_all();
    return true;
  }

  // ---- common case only; false => nothing decided, call push_full.
  // push_fast = fast_begin (set-up, order-2 guess, order-K solve described as a task) + the task's solve +
  // fast_end (integration, checks, hand-over).  Kernels that run the solves from a work queue call the two halves
  // separately; fast_end re-forms the Taylor vectors when `prepared` is false (bit-identical, same inputs).
  GB_HD bool fast_begin(int ind_tetr_in, int iface, const double *x, double vpar, double t_remain_in, SolveTask &t,
                        int &iface_new, double &tau_max)
  {
    init(ind_tetr_in, x, iface, vpar, t_remain_in);
    solver_iters = 0;
    fallback = 0;
    double tau = 0.0;
    iface_new = iface_init;
    // b, A and all Taylor vectors up to order K once; the order-2 guess reads the first three coefficients
    prepare<(K > 2 ? K : 2)>(z_init);
    {
      double cm[4][5];
#pragma unroll
      for (int f = 0; f < 4; f++) face_coeffs<2>(r.an[f], f == 0, z_init, cm[f], f);
      const bool have_exit = pick_exit<2>(cm, 0xFu, 0, iface_new, tau);
      // GATHER kernels: neighbour's record -> shared memory.  Called by every lane (the cooperative form is a warp operation)
      r.prefetch_next(*mp, have_exit ? r.nb(iface_new - 1) : 0);
      if (!have_exit) return false;
    }
    tau_max = tau * GB_EPS_TAU;
    if (mp->prefetch) prefetch_record<PHI>(*mp, r.nb(iface_new - 1), r.gmode != 0);   // the exit face is (almost always) this one
    t.kind = 1;
    t.tau = tau;
    t.deg = 0;
    if (K > 2) {
      if (!mp->boole_guess) return false;  // all-face order-K solves go through the complete path
      double n[3], c[5];
      face_normal(iface_new, n);
      face_coeffs<K>(n, iface_new == 1, z_init, c, iface_new - 1);
      face_task<K>(c, iface_new == iface_init, 0, t);
    }
    return true;
  }
  GB_HD bool fast_end(double tau, int iface_new, double tau_max, bool prepared, PushOut &o)
  {
    if (!((tau < GB_HUGE) && (tau > 0.0))) return false;
    if (!prepared) prepare<(K > 2 ? K : 2)>(z_init);
    double z[4] = {z_init[0], z_init[1], z_init[2], z_init[3]};
    integrate<K>(z, tau);
    if (!exit_point_ok(z, iface_new)) return false;
    if (ADAPT && adaptive_delta_energy(z) > mp->desired_delta_energy) return false;  // sub-stepping: complete path
    if (K > 2 && tau > tau_max) return false;
    if ((EXT == 4 ? normal_velocity(z, iface_new) : normal_v_from_trajectory(iface_new, tau)) > 0.0) return false;   // :328-332
    return finish<true>(z, tau, iface_new, o);
  }
  // t_remain_reload (kernel only): where the caller keeps t_remain; re-reading it after the solve -- same value --
  // ends its register live range at init(), so that it is not carried (spilled) across the solve
  GB_HD bool push_fast(int ind_tetr_in, int iface, const double *x, double vpar, double t_remain_in, PushOut &o,
                       const volatile double *t_remain_reload = nullptr)
  {
    SolveTask t;
    int iface_new;
    double tau_max;
    if (K > 2 && !mp->boole_guess) {
      // no face guess: order-K roots of all four faces
      init(ind_tetr_in, x, iface, vpar, t_remain_in);
      solver_iters = 0;
      fallback = 0;
      double z[4] = {z_init[0], z_init[1], z_init[2], z_init[3]}, tau = 0.0;
      iface_new = iface_init;
      prepare<K>(z);
      double cm[4][5];
#pragma unroll
      for (int f = 0; f < 4; f++) face_coeffs<K>(r.an[f], f == 0, z, cm[f], f);
      if (!pick_exit<2>(cm, 0xFu, 0, iface_new, tau)) return false;
      tau_max = tau * GB_EPS_TAU;
      iface_new = iface_init;
      if (!pick_exit<K>(cm, 0xFu, 0, iface_new, tau)) return false;
      if (t_remain_reload) t_remain = *t_remain_reload;
      return fast_end(tau, iface_new, tau_max, true, o);
    }
    if (!fast_begin(ind_tetr_in, iface, x, vpar, t_remain_in, t, iface_new, tau_max)) return false;
    double tau = t.tau;
    if (t.kind == 0) return false;
    if (t.kind == 2) tau = solve_monic_min_positive(t.deg, t.q[0], t.q[1], t.q[2], t.q[3], t.lambda, solver_iters);
    if (t_remain_reload) t_remain = *t_remain_reload;
    return fast_end(tau, iface_new, tau_max, true, o);
  }

  // ---- the complete ladder (:182-675)
  GB_HD void push_full(int ind_tetr_in, int iface, const double *x, double vpar, double t_remain_in, PushOut &o)
  {
    init(ind_tetr_in, x, iface, vpar, t_remain_in);
    solver_iters = 0;
    fallback = 0;
    double z[4] = {z_init[0], z_init[1], z_init[2], z_init[3]};
    double tau = 0.0;
    int iface_new = iface_init;
    bool approx = analytic_approx<2>(0xFu, 0, z, iface_new, tau);
    const double tau_max = tau * GB_EPS_TAU;
    if (K > 2) {
      const int guess = iface_new;
      iface_new = iface_init;
      if (mp->boole_guess && approx) approx = analytic_approx_single<K>(guess, 0, z, iface_new, tau, false);
      else approx = analytic_approx<K>(0xFu, 0, z, iface_new, tau);
    }
    bool face_correct = approx;
    if (face_correct) {
      integrate<K>(z, tau);
      if (ADAPT) overhead_adaptive<K>(0, mp->boole_guess != 0, true, iface_new, tau, z, face_correct);  // :316-319
      if (!three_planes_ok(z, iface_new)) face_correct = false;
      if (!face_converged(z, iface_new)) face_correct = false;
      if (K > 2 && tau > tau_max) face_correct = false;
      if (face_correct) {
        if ((EXT == 4 ? normal_velocity(z, iface_new) : normal_v_from_trajectory(iface_new, tau)) > 0.0) {
          if (K > 2) {
            face_correct = false;
          } else {
            if (!prolonged_trajectory(0, z, tau, iface_new, face_correct)) {
              set_removed(o);
              return;
            }
          }
        }
      }
    }
    if (!face_correct) {
      fallback |= 1;
      face_correct = true;
      iface_new = iface_init;
#pragma unroll
      for (int i = 0; i < 4; i++) z[i] = z_init[i];
      nsteps = 0;
      approx = analytic_approx<K>(0xFu, (K == 2) ? 1 : 0, z, iface_new, tau);
      if (!approx) {
        set_removed(o);
        return;
      }
      integrate<K>(z, tau);
      if (ADAPT) overhead_adaptive<K>((K == 2) ? 1 : 0, false, true, iface_new, tau, z, face_correct);  // :391-399
      if (K > 2 && tau > tau_max) face_correct = false;
      if (!three_planes_ok(z, iface_new)) face_correct = false;
      if (!face_converged(z, iface_new)) face_correct = false;
      if (face_correct) {
        if (normal_velocity(z, iface_new) > 0.0) {
          if (!prolonged_trajectory(0, z, tau, iface_new, face_correct)) {
            set_removed(o);
            return;
          }
        }
      }
      if (!face_correct) {
        if (ADAPT) main_fc = false;
        if (!trouble_shooting(z, tau, iface_new)) {
          set_removed(o);
          return;
        }
      }
    }
    finish<false>(z, tau, iface_new, o);
  }
};

// Non-inlined complete push: by-value in, by-value out, so that no hot-loop variable has its address taken.
template <int K, int PHI, int EXT = 0>
GB_HD_NOINLINE PushOut push_full_call(const MeshDev *mp, double perpinv, int ind_tetr, int iface, double x0,
                                      double x1, double x2, double vpar, double t_remain)
{
  PolyPusher<K, PHI, EXT> P;
  double stash[6];
  P.mp = mp;
  P.perpinv = perpinv;
  P.r.set_stash(stash, 1);
  PushOut o;
  double x[3] = {x0, x1, x2};
  o.x[0] = x0; o.x[1] = x1; o.x[2] = x2; o.vpar = vpar;
  o.z_save[0] = o.z_save[1] = o.z_save[2] = 0.0;
  o.t_pass = 0.0; // undefined in the reference when the particle is removed in attempts 1-3
  P.push_full(ind_tetr, iface, x, vpar, t_remain, o);
  return o;
}

// EXT = 2: the push result plus the optional quantities of the push
struct PushOutX {
  PushOut o;
  double oq[4];
  EvState es;
};
template <int K, int PHI, int EXT = 2>
GB_HD_NOINLINE PushOutX push_full_call_x(const MeshDev *mp, double perpinv, int ind_tetr, int iface, double x0, double x1,
                                         double x2, double vpar, double t_remain, unsigned oq_mask, int ev_flags,
                                         int nskip_p, int nskip_v, double J, int cnt_v, int cnt_p, double *lst = nullptr,
                                         int lst_cap = 0)
{
  PolyPusher<K, PHI, EXT> P;
  double stash[6];
  P.mp = mp;
  P.perpinv = perpinv;
  P.oq_mask = oq_mask;
  P.lst = lst;
  P.lst_cap = lst_cap;
  P.r.set_stash(stash, 1);
  PushOutX ox;
  PushOut &o = ox.o;
  double x[3] = {x0, x1, x2};
  o.x[0] = x0; o.x[1] = x1; o.x[2] = x2; o.vpar = vpar;
  o.z_save[0] = o.z_save[1] = o.z_save[2] = 0.0;
  o.t_pass = 0.0;
  P.push_full(ind_tetr, iface, x, vpar, t_remain, o);
  // a removed particle returns before the optional quantities are formed (:435-441): P.oq is still zero then
#pragma unroll
  for (int q = 0; q < 4; q++) ox.oq[q] = P.oq[q];
  ox.es.flags = ev_flags; ox.es.nskip_p = nskip_p; ox.es.nskip_v = nskip_v;
  ox.es.J = J; ox.es.cnt_v = cnt_v; ox.es.cnt_p = cnt_p; ox.es.n = 0;
  if (K >= 2 && ev_flags && !o.finished) P.events_after_push(vpar, o, ox.es);
  return ox;
}

// ---- small field helpers (SRC/supporting_functions_mod.f90:279-408) on the device layout -------------
template <int PHI>
GB_HD double bmod_at(const MeshDev &m, int ind_tetr, const double *z)
{
  const double *pb = m.bpart + ((int64_t)ind_tetr - 1) * BPART_ND;
  double g[3] = {ldg(pb + B_GB), ldg(pb + B_GB + 1), ldg(pb + B_GB + 2)};
  return ldg(pb + B_BMOD1) + dot3(g, z);
}

} // namespace gb
