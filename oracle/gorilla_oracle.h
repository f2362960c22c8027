/*
 * gorilla_oracle.h -- CPU ORACLE for the GORILLA orbit-pusher hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (gorilla_b200/) never links, imports or falls back to anything in oracle/.
 *
 * It is an op-for-op C99 restatement of the reference Fortran (itpplasma/GORILLA):
 *   SRC/orbit_timestep_gorilla.f90:19-147,278-358
 *   SRC/pusher_tetra_poly.f90:125-826,1258-1586,1767-2113,2690-2998
 *   SRC/pusher_tetra_func_mod.f90:6-93
 *   SRC/find_tetra_mod.f90:283-600   (+ the RK-module pieces it uses,
 *                                     SRC/pusher_tetra_rk.f90:50-193,840-896,2422-2467)
 *   SRC/supporting_functions_mod.f90:279-408
 *   SRC/contrib/Polynomial234RootSolvers.f90, SRC/contrib/cmplx_roots_sg.f90:83-201,558-731,906-1362
 *
 * PARITY STATUS: "parity unpinned".  The reference holds no golden vectors for this path
 * (both pusher tests are @disable'd, SRC/TESTS/test_orbit_timestep_gorilla.f90:26-42) and it
 * cannot be compiled in this image (no gfortran / NetCDF-Fortran / LAPACK).  What IS pinned:
 * the complex-arithmetic lowering (gcc -fcx-fortran-rules, the flag gfortran sets) and the
 * libm routines gfortran calls (glibc cabs/csqrt/cexp) are the real ones -- this file is
 * compiled with native `double _Complex` and those flags -- and the solver chain is checked
 * against constructed-root known-answer tests (tests/test_oracle_*.py).
 *
 * Data layout = the reference's own `sequence` derived types, flattened:
 *   tetra_physics : double [ntetr][142]  (SRC/tetra_physics_mod.f90:9-83)
 *   tetra_grid    : int32  [ntetr][20]   (SRC/tetra_grid_mod.f90:6-15)
 * All tetra / face indices are 1-based as in the reference; 0 = inside cell, -1 = lost.
 */
#ifndef GORILLA_ORACLE_H
#define GORILLA_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* offsets (in doubles) into one tetrahedron_physics record */
enum {
  TP_X1 = 0, TP_DIST_REF = 3, TP_DIST_REF_VEC = 4, TP_TETRA_DIST_REF = 8, TP_ANORM = 9,
  TP_CURLA = 21, TP_BMOD1 = 24, TP_ATHETA1 = 25, TP_APHI1 = 26, TP_H1_1 = 27, TP_H2_1 = 28,
  TP_H3_1 = 29, TP_PHI1 = 30, TP_R1 = 31, TP_Z1 = 32, TP_VE1_1 = 33, TP_VE2_1 = 34, TP_VE3_1 = 35,
  TP_V2EMOD_1 = 36, TP_ER_MOD = 37, TP_VE_MOD_AVG = 38, TP_SQG1 = 39, TP_DT_DTAU_CONST = 40,
  TP_GBXCURLA = 41, TP_GPHIXCURLA = 42, TP_GV2EMODXCURLA = 43, TP_GBXCURLVE = 44,
  TP_GPHIXCURLVE = 45, TP_GV2EMODXCURLVE = 46, TP_SPALPMAT = 47, TP_SPBETMAT = 48,
  TP_SPGAMMAT = 49, TP_GBXH1 = 50, TP_GPHIXH1 = 53, TP_GV2EMODXH1 = 56, TP_GB = 59, TP_GPHI = 62,
  TP_GR = 65, TP_GZ = 68, TP_GSQG = 71, TP_GATHETA = 74, TP_GAPHI = 77, TP_GH1 = 80, TP_GH2 = 83,
  TP_GH3 = 86, TP_CURLH = 89, TP_GVE1 = 92, TP_GVE2 = 95, TP_GVE3 = 98, TP_CURLVE = 101,
  TP_GV2EMOD = 104, TP_ALPMAT = 107, TP_BETMAT = 116, TP_GAMMAT = 125, TP_ACOEF_PRE = 134,
  TP_ACOEF_PRE_SE = 138, TP_NDOUBLES = 142
};
/* offsets (in doubles) into one tetrahedron_physics_precomp_poly4 record (SRC/tetra_physics_poly_precomp_mod.f90:21-45);
 * 4x4 matrices in Fortran column-major order, amat(i,j) at [i-1 + 4*(j-1)]; anorm_in_amat*(:,n) is column n */
enum {
  P4_AMAT1_0 = 0, P4_AMAT1_1 = 16, P4_AMAT2_0 = 32, P4_AMAT2_1 = 48, P4_AMAT2_2 = 64, P4_AMAT3_0 = 80, P4_AMAT3_1 = 96,
  P4_AMAT3_2 = 112, P4_AMAT3_3 = 128, P4_AMAT4_0 = 144, P4_AMAT4_1 = 160, P4_AMAT4_2 = 176, P4_AMAT4_3 = 192,
  P4_AMAT4_4 = 208, P4_AN_AMAT1_0 = 224, P4_AN_AMAT1_1 = 240, P4_AN_AMAT2_0 = 256, P4_AN_AMAT2_1 = 272,
  P4_AN_AMAT2_2 = 288, P4_AN_AMAT3_0 = 304, P4_AN_AMAT3_1 = 320, P4_AN_AMAT3_2 = 336, P4_AN_AMAT3_3 = 352,
  P4_AN_AMAT4_0 = 368, P4_AN_AMAT4_1 = 384, P4_AN_AMAT4_2 = 400, P4_AN_AMAT4_3 = 416, P4_AN_AMAT4_4 = 432,
  P4_B0 = 448, P4_B1 = 452, P4_B2 = 456, P4_B3 = 460, P4_A10_B0 = 464, P4_A10_B1 = 468, P4_A10_B2 = 472, P4_A10_B3 = 476,
  P4_A11_B0 = 480, P4_A11_B1 = 484, P4_A11_B2 = 488, P4_A11_B3 = 492, P4_AN_B0 = 496, P4_AN_B1 = 500, P4_AN_B2 = 504,
  P4_AN_B3 = 508, P4_AN_A10_B0 = 512, P4_AN_A10_B1 = 516, P4_AN_A10_B2 = 520, P4_AN_A10_B3 = 524, P4_AN_A11_B0 = 528,
  P4_AN_A11_B1 = 532, P4_AN_A11_B2 = 536, P4_AN_A11_B3 = 540, P4_NDOUBLES = 544
};
enum { TG_IND_KNOT = 0, TG_NEIGHBOUR_TETR = 4, TG_NEIGHBOUR_FACE = 8, TG_PERBOU_PHI = 12,
       TG_PERBOU_THETA = 16, TG_NINTS = 20 };

typedef struct {
  int64_t ntetr;
  const double *tetra_physics;   /* [ntetr][142] */
  const int32_t *tetra_grid;     /* [ntetr][20]  */
  /* scalars of tetra_physics_mod / tetra_grid_settings_mod / gorilla_settings_mod */
  double cm_over_e, particle_mass, particle_charge;
  int32_t sign_sqg, coord_system, n_field_periods, grid_kind;
  int32_t grid_size[3];
  double Rmin, Rmax, Zmin, Zmax;          /* rectangular grids only (find_tetra) */
  double sfc_s_min;
  int32_t ipusher;                        /* 1 = RK4, 2 = polynomial */
  int32_t poly_order;                     /* 1..4 */
  int32_t boole_guess;
  int32_t boole_strong_electric_field;
  int32_t boole_periodic_relocation;
  int32_t boole_dt_dtau;                  /* RK only */
  int32_t i_time_tracing_option;          /* 1 = dt/dtau constant per cell, 2 = Hamiltonian time (polynomial pusher only) */
  /* optional quantities of pusher_tetra_poly (gorilla_settings_mod.f90:51-55) */
  int32_t boole_time_hamiltonian, boole_gyrophase, boole_vpar_int, boole_vpar2_int;
  /* adaptive energy-controlled sub-stepping (gorilla_settings_mod.f90:75-77) */
  int32_t boole_adaptive_time_steps, max_n_intermediate_steps;
  double desired_delta_energy;
  /* handover_processing_kind = 2: tetra_skew_coord [ntetr][168] (tetra_physics_mod.f90:89-99), else NULL / 1 */
  const double *tetra_skew_coord;
  int32_t handover_processing_kind;
  /* precomputed-coefficient modes (SRC/tetra_physics_poly_precomp_mod.f90:160-476): i_precomp = 1, 2 of the polynomial
   * pusher, boole_newton_precalc of the RK pusher; tetra_physics_poly4 [ntetr][544] from gor_make_precomp_poly4 */
  int32_t i_precomp, boole_newton_precalc;
  const double *tetra_physics_poly4;
  /* RK pusher with adaptive RKF45 steps (boole_pusher_ode45, rel_err_ode45; SRC/pusher_tetra_rk.f90:2549-2581,
   * SRC/odeint_rkf45.f90, SRC/contrib/rkf45.f90) */
  int32_t boole_pusher_ode45, pad_ode45;
  double rel_err_ode45;
} gor_mesh;

/* make_precomp_poly4 (SRC/tetra_physics_poly_precomp_mod.f90:160-476): out = [ntetr][544] */
void gor_make_precomp_poly4(const gor_mesh *m, double *out);

/* optional per-particle trace of the visited (ind_tetr, iface) sequence */
typedef struct {
  int64_t n_pushes;        /* pusher invocations (the metric's "tetra crossings") */
  int64_t cap;             /* capacity of the arrays below (0 = do not record) */
  int32_t *ind_tetr;       /* state AFTER push k */
  int32_t *iface;
  int64_t n_fallback[4];   /* [0] 2nd attempt, [1] trouble shooting, [2] prolonged, [3] finish-outside */
  int64_t n_solver_iters;  /* Laguerre/SG/Newton iterations incl. polish */
  int64_t n_solver_calls;
  /* sum over the pushes of the time step of pusher_tetra_poly's optional_quantities
   * {t_hamiltonian, gyrophase, vpar_int, vpar2_int} (type optional_quantities_type, gorilla_settings_mod.f90:9-15) */
  double optional_quantities[4];
  int64_t n_adaptive;      /* pushes / segments that were re-integrated in sub-steps (adaptive scheme) */
} gor_trace;

/* orbit events (gorilla_plot_mod.f90:585-638, par_adiab_inv_poly_mod pusher_tetra_poly.f90:3156-3429) */
enum { GOR_EVENT_PHI_0 = 1, GOR_EVENT_VPAR_0 = 2, GOR_EVENT_FULL_ORBIT = 3 };
typedef struct {
  int64_t particle;   /* index given by the caller */
  int32_t kind;       /* GOR_EVENT_PHI_0: toroidal mapping, value = {p_phi, e_tot}; GOR_EVENT_VPAR_0: banana tip, value = {J_par, e_tot} */
  int32_t counter;    /* counter_phi_0_mappings / counter_banana_mappings at the event */
  int64_t push;       /* index of the push within this call (0-based) */
  double x[3];        /* position written to poincare_plot_phi_0 / poincare_plot_vpar_0 */
  double value[2];
  double t;           /* t_step - t_remain after the push of the event (gorilla_plot_mod.f90:564,572) */
} gor_event;
typedef struct {
  int32_t boole_poincare_phi_0, n_skip_phi_0, boole_poincare_vpar_0, boole_J_par, n_skip_vpar_0;
  /* boole_full_orbit (gorilla_plot_mod.f90:553-579): position, p_phi and E_tot after every n_skip_full_orbit-th push */
  int32_t boole_full_orbit, n_skip_full_orbit, reserved;
} gor_event_settings;

/* return codes */
enum { GOR_OK = 0, GOR_ERR_DOMAIN = 1, GOR_ERR_CONFIG = 2 };

int gor_check_coordinate_domain(const gor_mesh *m, double x[3]);
void gor_find_tetra(const gor_mesh *m, double x[3], double vpar, double vperp,
                    int32_t *ind_tetr, int32_t *iface, int sign_t_step);
int gor_orbit_timestep(const gor_mesh *m, double x[3], double *vpar, double *vperp, double t_step,
                       int32_t *boole_initialized, int32_t *ind_tetr, int32_t *iface,
                       double *t_remain_out, gor_trace *trace);
int gor_orbit_timestep_events(const gor_mesh *m, double x[3], double *vpar, double *vperp, double t_step,
                              int32_t *boole_initialized, int32_t *ind_tetr, int32_t *iface, double *t_remain_out,
                              gor_trace *trace, const gor_event_settings *cfg, double *par_adiab_inv /* inout */,
                              int32_t *counter_vpar_0 /* inout */, int32_t *counter_phi_0 /* inout */, int64_t particle,
                              gor_event *events, int64_t cap, int64_t *n_events /* inout: events so far (may exceed cap) */);
/* OpenMP batch driver used as the CPU baseline ("one particle per thread", README.md:181) */
int64_t gor_orbit_timestep_batch(const gor_mesh *m, int64_t n, double *x /*[n][3]*/, double *vpar,
                                 double *vperp, double t_step, int32_t *boole_initialized,
                                 int32_t *ind_tetr, int32_t *iface, double *t_remain_out,
                                 int64_t *n_pushes /*[n] or NULL*/, int nthreads);
/* same, also returning the optional quantities summed along each particle's time step: [n][4] or NULL */
int64_t gor_orbit_timestep_batch_opt(const gor_mesh *m, int64_t n, double *x, double *vpar, double *vperp,
                                     double t_step, int32_t *boole_initialized, int32_t *ind_tetr, int32_t *iface,
                                     double *t_remain_out, int64_t *n_pushes, double *optional_quantities,
                                     int nthreads);

double gor_energy_tot(const gor_mesh *m, const double z[4], double perpinv, int32_t ind_tetr);
double gor_p_phi(const gor_mesh *m, double vpar, const double z[3], int32_t ind_tetr);
double gor_bmod(const gor_mesh *m, const double z[3], int32_t ind_tetr);

/* solver chain exposed for known-answer tests; root is [n][2] column-major as in Fortran: root(i,1)=Re */
void gor_quadratic_roots(double q1, double q0, int *nreal, double root[4]);
void gor_cubic_roots(double c2, double c1, double c0, int *nreal, double root[6]);
void gor_quartic_roots(double q3, double q2, double q1, double q0, int *nreal, double root[8]);
/* raw complex roots (re,im interleaved), same call as cmplx_roots_gen(roots,poly,deg,.true.,.false.) */
void gor_cmplx_roots_gen(int degree, const double *poly_re_im, double *roots_re_im);
double gor_quadratic_solver1(double a, double b, double c);
double gor_quadratic_solver2(double a, double b, double c);
double gor_cubic_solver(double a, double b, double c, double d);
double gor_quartic_solver(int i_scaling, double a, double b, double c, double d, double e);
/* cexp(i*2*pi*FRAC_JUMPS[k]) table entries as glibc computes them (for the device constant table) */
void gor_frac_jump_phase(int k, double out[2]);

/* probes of gcc -fcx-fortran-rules complex lowering and glibc cabs/csqrt (pin the device restatements) */
void gor_probe_csqrt(double re, double im, double out[2]);
double gor_probe_cabs(double re, double im);
void gor_probe_cdiv(double ar, double ai, double br, double bi, double out[2]);
void gor_probe_cmul(double ar, double ai, double br, double bi, double out[2]);
void gor_probe_rmul(double r0, double br, double bi, double out[2]);

#ifdef __cplusplus
}
#endif
#endif
