/*
 * gorilla_b200.h -- C ABI of the B200-native GORILLA orbit pusher (libgorilla_b200.so).
 *
 * This is the drop-in boundary for the reference's particle-parallel hot path.  GORILLA itself is
 * Fortran 90 with no C ABI; the entry points below are what its module procedures for this path bind
 * to through ISO_C_BINDING (see INTEGRATION.md and gorilla_b200/fortran/orbit_timestep_gorilla_b200_mod.f90).
 * Every function cites the reference interface it replaces (paths relative to the GORILLA tree).
 *
 * Conventions (identical to the reference):
 *   - tetrahedron and face indices are 1-based; iface = 0 means "inside the cell";
 *     ind_tetr = -1 (and iface = -1) means the particle left the domain / was removed.
 *   - x is x(3,n) in Fortran order, i.e. C double[n][3]: (R,phi,Z) for coord_system 1, (s,theta,phi) for 2.
 *   - Gaussian CGS units: cm, s, cm/s, Gauss, statvolt.
 *   - per-particle logicals are int32 (0 = .false., non-zero = .true.).
 * All functions return GORILLA_OK (0) or a GORILLA_ERR_* code; nothing calls exit()/stop.
 * There is no CPU fall-back: every compute entry point runs CUDA kernels on the current device and
 * returns GORILLA_ERR_CUDA if that is impossible.
 */
#ifndef GORILLA_B200_H
#define GORILLA_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
  GORILLA_OK = 0,
  GORILLA_ERR_ARG = 1,          /* null pointer / bad size */
  GORILLA_ERR_UNSUPPORTED = 2,  /* option of gorilla.inp that this path does not implement (yet) */
  GORILLA_ERR_CUDA = 3,         /* CUDA runtime failure (gorilla_b200_last_error() has the text) */
  GORILLA_ERR_DOMAIN = 4,       /* a start position is outside the computation domain
                                   (reference: print + stop, orbit_timestep_gorilla.f90:299-352) */
  GORILLA_ERR_IO = 5
};

#define GORILLA_TETRA_PHYSICS_NDOUBLES 142 /* sizeof(type tetrahedron_physics)/8, tetra_physics_mod.f90:9-83 */
#define GORILLA_TETRA_GRID_NINTS 20        /* sizeof(type tetrahedron_grid)/4,    tetra_grid_mod.f90:6-15   */
#define GORILLA_TETRA_SKEW_NDOUBLES 168     /* sizeof(type tetrahedron_skew_coord)/8, tetra_physics_mod.f90:89-99 */

/* The subset of namelist GORILLANML (gorilla_settings_mod.f90:94-105) that the hot path reads. */
typedef struct gorilla_settings {
  double eps_Phi;
  int32_t coord_system;              /* 1 (R,phi,Z) | 2 (s,theta,phi) */
  int32_t ispecies;                  /* 1 e-, 2 D+, 3 alpha, 4 W74+ (orbit_timestep_gorilla.f90:204-249) */
  int32_t boole_periodic_relocation;
  int32_t ipusher;                   /* 1 RK4 | 2 polynomial */
  int32_t boole_pusher_ode45;        /* ipusher = 1: every integration step with the adaptive RKF45 integrator (rel_err_ode45)
                                        instead of one RK4 step (pusher_tetra_rk.f90:2549-2581, contrib/rkf45.f90) */
  int32_t boole_dt_dtau;             /* must be 1 */
  int32_t boole_newton_precalc;      /* ipusher = 1: normal velocity / acceleration and the quadratic start guess from the
                                        tetra_physics_poly4 records (pusher_tetra_rk.f90:579-632, 2487-2527) */
  int32_t poly_order;                /* 1..4 */
  int32_t i_precomp;                 /* 0 | 1 (poly_order 2..4) | 2 (poly_order 2): precomputed coefficients, the library
                                        forms the tetra_physics_poly4 records (tetra_physics_poly_precomp_mod.f90:160-476,
                                        4352 bytes per tetrahedron) itself at init; polynomial pusher, not combined with
                                        Hamiltonian time / optional quantities / adaptive steps / strong E / hand-over kind 2 */
  int32_t boole_guess;
  int32_t i_time_tracing_option;     /* 1 dt/dtau constant per cell | 2 Hamiltonian time (ipusher = 2 only,
                                        gorilla_settings_mod.f90:124-129) */
  int32_t handover_processing_kind;  /* 1 periodic shifts | 2 position exchange via Cartesian skew coordinates
                                        (pusher_tetra_func_mod.f90:59-89; needs gorilla_mesh_desc.tetra_skew_coord) */
  int32_t boole_adaptive_time_steps; /* energy-controlled sub-stepping (pusher_tetra_poly.f90:830-1254); polynomial pusher.
                                        Combined with Hamiltonian time tracing / optional quantities / events the step
                                        lists are kept in full (3 * max_n_intermediate_steps entries of 40 bytes per
                                        device thread, allocated on first use; refused above 64 GB) */
  int32_t boole_strong_electric_field; /* ExB-drift terms of order v_E^2; cylindrical grids (coord_system 1) only */
  int32_t boole_grid_for_find_tetra; /* ignored: the device scan does not need the box accelerator */
  /* optional quantities of pusher_tetra_poly (gorilla_settings_mod.f90:51-55; ipusher = 2 only); boole_gyrophase
   * requires boole_time_Hamiltonian (:132-135) */
  int32_t boole_time_Hamiltonian;
  int32_t boole_gyrophase;
  int32_t boole_vpar_int;
  int32_t boole_vpar2_int;
  int32_t max_n_intermediate_steps;  /* adaptive scheme: >= 2 (INPUT/gorilla.inp:151) */
  double desired_delta_energy;       /* adaptive scheme: > 0, relative energy error per tetrahedron (gorilla.inp:147) */
  double rel_err_ode45;              /* boole_pusher_ode45: relative error of the RKF45 integrator (gorilla.inp:43, 1e-8) */
  /* analytical helical perturbation of the equilibrium (gorilla.inp:118-125; vector_potential_sthetaphi,
   * tetra_physics_mod.f90:1158-1161): A_phi += A_phi * eps * cos(m * theta + n * phi) at the vertices.  Read by
   * gorilla_mesh_build for grid_kind 2 only (as in the reference); ignored by gorilla_b200_init (the records carry it). */
  double helical_pert_eps_Aphi;
  int32_t boole_helical_pert;
  int32_t helical_pert_m_fourier;
  int32_t helical_pert_n_fourier;
  int32_t reserved0;
  /* random noise on the vertex values of a built mesh, for testing how the integrator copes with a rough field
   * (gorilla.inp:84-109; tetra_physics_mod.f90:256-261,400-415,441-444): A_k += A_k * eps * r with r uniform in [0, 1) --
   * one r per vertex of a poloidal plane repeated in every plane (axisymmetric, vector potential and / or electrostatic
   * potential) or three fresh r per vertex (non-axisymmetric, vector potential).  Read by gorilla_mesh_build, all grid kinds.
   * The reference draws from gfortran's random_number; this library from its own xoshiro256** stream seeded with noise_seed
   * (0 = a fixed default): the same noise statistically, not the same numbers. */
  double axi_noise_eps_A, axi_noise_eps_Phi, non_axi_noise_eps_A;
  int32_t boole_axi_noise_vector_pot, boole_axi_noise_elec_pot, boole_non_axi_noise_vector_pot;
  int32_t noise_seed;
} gorilla_settings;

/* Everything initialize_gorilla() (orbit_timestep_gorilla.f90:151-274) leaves in module variables that
 * the hot path reads afterwards.  A Fortran caller fills it from tetra_physics_mod / tetra_grid_mod /
 * tetra_grid_settings_mod with c_loc() -- both derived types are `sequence` types of a single kind, so
 * the arrays are plain [ntetr][142] doubles and [ntetr][20] int32. */
typedef struct gorilla_mesh_desc {
  int64_t ntetr;
  const double *tetra_physics;  /* tetra_physics(1:ntetr)  */
  const int32_t *tetra_grid;    /* tetra_grid(1:ntetr)     */
  double cm_over_e;             /* tetra_physics_mod: cm_over_e       */
  double particle_mass;         /*                    particle_mass   */
  double particle_charge;       /*                    particle_charge */
  int32_t sign_sqg;             /* tetra_physics_mod: sign_sqg        */
  int32_t coord_system;
  int32_t n_field_periods;      /* tetra_grid_settings_mod */
  int32_t grid_kind;
  int32_t grid_size[3];
  int32_t pad0;
  double Rmin, Rmax, Zmin, Zmax; /* tetra_grid_mod (rectangular grids; unused otherwise) */
  double sfc_s_min;              /* tetra_grid_settings_mod */
  const double *tetra_skew_coord; /* tetra_skew_coord(1:ntetr) (`sequence` type, 168 doubles, tetra_physics_mod.f90:89-99);
                                     NULL unless handover_processing_kind = 2 */
} gorilla_mesh_desc;

typedef struct gorilla_b200_handle gorilla_b200_handle;

/* Aggregates of one orbit_timestep call (the reference keeps such counters in gorilla_plot_mod.f90:
 * counter_tetrahedron_passes :550, lost-particle counter :290-294). */
typedef struct gorilla_counters {
  int64_t n_particles;
  int64_t n_pushes;          /* pusher invocations = "tetra crossings" of the metric */
  int64_t n_lost;            /* ind_tetr == -1 after the call */
  int64_t n_finished;        /* boole_t_finished */
  int64_t n_fallback[4];     /* pushes that needed: 2nd attempt, trouble shooting, prolonged trajectory,
                                stop-inside landing outside the cell */
  int64_t n_domain_errors;
  double kernel_ms;          /* device time of the push kernel (CUDA events on the launch stream) */
  double find_ms;            /* device time of the localisation kernel, 0 if not run */
  int64_t n_adaptive;        /* pushes in which the adaptive scheme re-integrated a segment in sub-steps */
  int64_t n_lost_inner;      /* lost IN this call through the inner boundary s = sfc_s_min (flux-coordinate grids) */
  int64_t n_failed;          /* lost IN this call: removed by the pusher inside the domain (no valid exit time / trouble shooting
                                failed, pusher_tetra_poly.f90:434-440,609-615) */
} gorilla_counters;

/* ---- lifetime ---------------------------------------------------------------------------------- */

/* Uploads the mesh (repacked into sub-record SoA) to the CURRENT CUDA device.
 * Replaces the "library owns the mesh in module arrays" half of initialize_gorilla. */
int gorilla_b200_init(const gorilla_mesh_desc *mesh, const gorilla_settings *settings,
                      gorilla_b200_handle **out);
void gorilla_b200_free(gorilla_b200_handle *h);
const char *gorilla_b200_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py "gpu_launches") */
int64_t gorilla_b200_launch_count(void);

/* ---- the hot path ------------------------------------------------------------------------------ */

/* Batched orbit_timestep_gorilla(x,vpar,vperp,t_step,boole_initialized,ind_tetr,iface,t_remain_out)
 * (orbit_timestep_gorilla.f90:19-147) for n independent particles; HOST pointers, copies included.
 * t_remain_out and n_pushes may be NULL.  n = 1 is exactly the reference's scalar call. */
int gorilla_b200_orbit_timestep(gorilla_b200_handle *h, int64_t n, double *x, double *vpar, double *vperp,
                                double t_step, int32_t *boole_initialized, int32_t *ind_tetr, int32_t *iface,
                                double *t_remain_out, int64_t *n_pushes);

/* Same, all pointers are DEVICE pointers on the handle's device; runs on `stream` (a cudaStream_t,
 * NULL = default stream) and does not synchronise.  Every call gets its own device counter block and work queue (a ring of
 * 8 per handle; a 9th call in flight waits for the oldest), so calls on different streams of one handle do not interfere.
 * The handle's device is made current for the duration of every entry point. */
int gorilla_b200_orbit_timestep_dev(gorilla_b200_handle *h, int64_t n, double *x, double *vpar, double *vperp,
                                    double t_step, int32_t *boole_initialized, int32_t *ind_tetr, int32_t *iface,
                                    double *t_remain_out, int64_t *n_pushes, void *stream);

/* As gorilla_b200_orbit_timestep, additionally recording the (ind_tetr, iface) state after each of the
 * first trace_cap pushes of every particle: trace_* are HOST int32 [n][trace_cap], unused slots = 0.
 * Used by the parity tests ("visited tetra sequence bit-exact for the first N crossings"). */
int gorilla_b200_orbit_timestep_trace(gorilla_b200_handle *h, int64_t n, double *x, double *vpar, double *vperp,
                                      double t_step, int32_t *boole_initialized, int32_t *ind_tetr,
                                      int32_t *iface, double *t_remain_out, int64_t *n_pushes,
                                      int32_t trace_cap, int32_t *trace_ind_tetr, int32_t *trace_iface);

/* As gorilla_b200_orbit_timestep, additionally returning the optional quantities of pusher_tetra_poly
 * (type optional_quantities_type, gorilla_settings_mod.f90:9-15; pusher_tetra_poly.f90:204,228-229,662-667,2134-2210):
 * optional_quantities is HOST double [n][4] = { t_hamiltonian, gyrophase, vpar_int, vpar2_int }, each the SUM over the
 * pushes of this time step of the value the reference pusher returns per push (what a caller of the pusher accumulates
 * along the orbit).  Only the quantities switched on in gorilla_settings (boole_time_Hamiltonian, boole_gyrophase,
 * boole_vpar_int, boole_vpar2_int) are formed, the others are 0.  Polynomial pusher only (GORILLA_ERR_UNSUPPORTED for
 * ipusher = 1, as in the reference where pusher_tetra_rk has no such argument). */
int gorilla_b200_orbit_timestep_optional(gorilla_b200_handle *h, int64_t n, double *x, double *vpar, double *vperp,
                                         double t_step, int32_t *boole_initialized, int32_t *ind_tetr, int32_t *iface,
                                         double *t_remain_out, int64_t *n_pushes, double *optional_quantities);
/* Same with DEVICE pointers on `stream`, no synchronisation. */
int gorilla_b200_orbit_timestep_optional_dev(gorilla_b200_handle *h, int64_t n, double *x, double *vpar, double *vperp,
                                             double t_step, int32_t *boole_initialized, int32_t *ind_tetr,
                                             int32_t *iface, double *t_remain_out, int64_t *n_pushes,
                                             double *optional_quantities, void *stream);

/* ---- orbit events: J_par / banana tips / toroidal mappings ---------------------------------------------
 * The event capture of gorilla_plot_orbit_integration (gorilla_plot_mod.f90:433-658: after every push that does not end
 * the time step, par_adiab_inv_tetra_poly :585-596 and the phi = 0 mappings :601-638) with module par_adiab_inv_poly_mod
 * (pusher_tetra_poly.f90:3156-3429), as a device-side event buffer instead of the reference's text files
 * (poincare_plot_phi_0 / poincare_plot_vpar_0 / J_par / e_tot / p_phi, gorilla_plot.inp). */
enum { GORILLA_EVENT_PHI_0 = 1, GORILLA_EVENT_VPAR_0 = 2, GORILLA_EVENT_FULL_ORBIT = 3 };
typedef struct gorilla_event {
  int64_t particle;  /* index into the batch */
  int32_t kind;      /* GORILLA_EVENT_PHI_0: toroidal mapping, value = { p_phi_func, energy_tot_func };
                        GORILLA_EVENT_VPAR_0: banana tip (v_par = 0), value = { J_par of the completed bounce, energy_tot_func };
                        GORILLA_EVENT_FULL_ORBIT: the orbit point after a push (boole_full_orbit, gorilla_plot_mod.f90:553-579:
                        the full_orbit_plot / p_phi / e_tot files), value = { p_phi_func, energy_tot_func } in the tetrahedron
                        the push ran in */
  int32_t counter;   /* counter_phi_0_mappings resp. counter_banana_mappings at the event; full orbit: the number of pushes of
                        this call so far (counter_tetrahedron_passes), a multiple of n_skip_full_orbit */
  int64_t push;      /* index of the push within this call (0-based) */
  double x[3];       /* the position the reference writes to poincare_plot_phi_0_* / poincare_plot_vpar_0_* */
  double value[2];
  double t;          /* elapsed part of the time step after this push, t_step - t_remain (what the reference writes next to
                        p_phi / e_tot); set for every kind */
} gorilla_event;
typedef struct gorilla_event_settings { /* namelist GORILLA_PLOT_NML (gorilla_plot_mod.f90) */
  int32_t boole_poincare_phi_0, n_skip_phi_0;
  int32_t boole_poincare_vpar_0, boole_J_par, n_skip_vpar_0;
  int32_t boole_full_orbit, n_skip_full_orbit; /* one GORILLA_EVENT_FULL_ORBIT record after every n_skip_full_orbit-th push,
                                                  the one that ends the time step included (:553-556) */
  int32_t reserved;
} gorilla_event_settings;
/* As gorilla_b200_orbit_timestep, with event capture.  par_adiab_inv / counter_vpar_0 / counter_phi_0 are HOST [n]
 * in/out arrays holding the per-particle state of par_adiab_inv_poly_mod and of the mapping counters between calls (zero
 * them before the first call).  events: HOST buffer of event_cap records, filled in no particular order (sort by
 * particle, push); *n_events returns the number of events that occurred, which may exceed event_cap (the surplus is
 * dropped).  Polynomial pusher of order 2..4 (par_adiab_tau, pusher_tetra_poly.f90:3302-3320, has no case for order 1) or
 * the RK pusher (module par_adiab_inv_rk_mod, pusher_tetra_rk.f90:2589-2798: J_par as a fifth RKF45 equation integrated
 * with rel_err_ode45). */
int gorilla_b200_orbit_timestep_events(gorilla_b200_handle *h, int64_t n, double *x, double *vpar, double *vperp,
                                       double t_step, int32_t *boole_initialized, int32_t *ind_tetr, int32_t *iface,
                                       double *t_remain_out, int64_t *n_pushes, const gorilla_event_settings *cfg,
                                       double *par_adiab_inv, int32_t *counter_vpar_0, int32_t *counter_phi_0,
                                       gorilla_event *events, int64_t event_cap, int64_t *n_events);
/* Same with DEVICE pointers (n_events: DEVICE uint64 counter, zeroed by the caller) on `stream`, no synchronisation. */
int gorilla_b200_orbit_timestep_events_dev(gorilla_b200_handle *h, int64_t n, double *x, double *vpar, double *vperp,
                                           double t_step, int32_t *boole_initialized, int32_t *ind_tetr, int32_t *iface,
                                           double *t_remain_out, int64_t *n_pushes, const gorilla_event_settings *cfg,
                                           double *par_adiab_inv, int32_t *counter_vpar_0, int32_t *counter_phi_0,
                                           gorilla_event *events, int64_t event_cap, uint64_t *n_events, void *stream);

/* check_coordinate_domain + find_tetra(x,vpar,vperp,ind_tetr,iface,sign_t_step)
 * (orbit_timestep_gorilla.f90:278-358, find_tetra_mod.f90:283-600); HOST pointers. x may be modified
 * (periodic relocation, start points lying on a face). */
int gorilla_b200_find_tetra(gorilla_b200_handle *h, int64_t n, double *x, const double *vpar, const double *vperp,
                            int32_t *ind_tetr, int32_t *iface, int32_t sign_t_step);

/* ---- diagnostics ------------------------------------------------------------------------------- */

/* Per-particle invariants, HOST pointers, any output may be NULL:
 *   energy  = energy_tot_func   (supporting_functions_mod.f90:279-301)
 *   p_phi   = p_phi_func        (:377-408)
 *   perpinv = -vperp^2/(2 |B|)  (orbit_timestep_gorilla.f90:77), the conserved magnetic-moment proxy
 * Particles with ind_tetr < 1 get NaN. */
int gorilla_b200_invariants(gorilla_b200_handle *h, int64_t n, const double *x, const double *vpar,
                            const double *vperp, const int32_t *ind_tetr, double *energy, double *p_phi,
                            double *perpinv);
int gorilla_b200_invariants_dev(gorilla_b200_handle *h, int64_t n, const double *x, const double *vpar,
                                const double *vperp, const int32_t *ind_tetr, double *energy, double *p_phi,
                                double *perpinv, void *stream);

/* Counters of the most recent orbit_timestep* call on this handle (waits for that call to complete). */
int gorilla_b200_get_counters(gorilla_b200_handle *h, gorilla_counters *out);

/* ---- periodic particle re-sorting by tetrahedron index (gather locality; north_star) -------------------------- */

/* Fills perm (DEVICE int64[n]) with the permutation that orders particles by ind_tetr (lost particles last). */
int gorilla_b200_sort_permutation_dev(gorilla_b200_handle *h, int64_t n, const int32_t *ind_tetr, int64_t *perm,
                                      void *stream);
/* Re-sorts a resident batch in place: computes that permutation and applies it to the six state arrays (DEVICE pointers;
 * boole_initialized may be NULL) and to n_extra further DEVICE double[n] arrays the caller keeps per particle (e.g. the
 * reference invariants of gorilla_b200_diag_reduce_dev); perm_out (DEVICE int64[n], may be NULL) receives the permutation,
 * new[i] = old[perm[i]].  One pass over the data per array; does not synchronise. */
int gorilla_b200_resort_dev(gorilla_b200_handle *h, int64_t n, double *x, double *vpar, double *vperp,
                            int32_t *boole_initialized, int32_t *ind_tetr, int32_t *iface, int32_t n_extra,
                            double *const *extra, int64_t *perm_out, void *stream);
/* on != 0: the HOST-pointer entry point gorilla_b200_orbit_timestep sorts each uploaded batch by tetrahedron on the device
 * before the push and restores the caller's order before the download (results are identical: particles are independent).
 * Off by default; ignored by the trace / optional-quantity / event variants. */
int gorilla_b200_set_host_resort(gorilla_b200_handle *h, int32_t on);

/* ---- diagnostics reduction and multi-GPU ---------------------------------------------------------------------
 * The reference is one OpenMP process; its counters live in gorilla_plot_mod (counter_tetrahedron_passes :550, lost
 * particles :290-294,488-491) and it writes E_tot / p_phi per time step for the user to compare (:603).  Here particles shard
 * over the GPUs of one box with the mesh replicated on each (one process and one handle per GPU); the only exchange of the
 * path is the reduction below.  NCCL is loaded at run time (libnccl.so.2; GORILLA_B200_NCCL_LIB overrides the name). */
typedef struct gorilla_diag {
  int64_t n_particles;       /* particles in the reduced batches, all ranks */
  int64_t n_pushes;          /* tetra crossings since gorilla_b200_diag_reset (or init), all ranks */
  int64_t n_lost;            /* particles lost since the reset = n_lost_outer + n_lost_inner + n_failed */
  int64_t n_lost_outer;      /* left through the outer boundary (flux grids: s = 1; cylindrical grids: any boundary face) */
  int64_t n_lost_inner;      /* left through the inner boundary s = sfc_s_min (flux-coordinate grids) */
  int64_t n_failed;          /* removed by the pusher inside the domain */
  int64_t n_finished;        /* completed time steps (boole_t_finished) */
  int64_t n_fallback[4];     /* as gorilla_counters */
  int64_t n_adaptive;
  int64_t n_sampled;         /* particles that entered the drift statistics (still in the domain, finite reference) */
  /* conservation since the reference values were taken: max and rms over the sampled particles of
   * |E/E_ref - 1| (energy_tot_func), |perpinv/perpinv_ref - 1| (magnetic moment) and |p_phi/p_phi_ref - 1| (p_phi_func) */
  double max_delta_energy, rms_delta_energy;
  double max_delta_perpinv, rms_delta_perpinv;
  double max_delta_p_phi, rms_delta_p_phi;
  int32_t nranks, reserved;
} gorilla_diag;

#define GORILLA_COMM_ID_BYTES 128 /* sizeof(ncclUniqueId) */
/* Rank 0 creates the id (ncclGetUniqueId) and hands the bytes to the other ranks by whatever means the host program has
 * (MPI_Bcast, a file, torch.distributed); then every rank joins with its handle (ncclCommInitRank on the handle's device). */
int gorilla_b200_comm_unique_id(void *id_out /* GORILLA_COMM_ID_BYTES */);
int gorilla_b200_comm_init(gorilla_b200_handle *h, const void *id, int32_t rank, int32_t nranks);
int gorilla_b200_comm_free(gorilla_b200_handle *h);
/* In-place all-reduce of a small DEVICE double buffer over the handle's communicator (op: 0 sum, 1 max, 2 min); a handle
 * without communicator (single GPU) leaves the buffer as it is.  For what a driver reduces besides the diagnostics (timings). */
int gorilla_b200_comm_allreduce_f64(gorilla_b200_handle *h, double *buf, int64_t count, int32_t op, void *stream);
/* The contiguous shard [first, first+count) of n_total particles that belongs to `rank` of `nranks`
 * ([r N/G, (r+1) N/G), BASELINE config 5). */
int gorilla_b200_shard_range(int64_t n_total, int32_t rank, int32_t nranks, int64_t *first, int64_t *count);

/* Zeroes the counters that the orbit_timestep* calls of this handle accumulate for gorilla_b200_diag_reduce_dev. */
int gorilla_b200_diag_reset(gorilla_b200_handle *h, void *stream);
/* One device reduction kernel over the batch (DEVICE pointers) + one grouped NCCL all-reduce when the handle has a
 * communicator; *out (HOST) is complete on return (synchronises `stream`).  energy_ref / p_phi_ref / perpinv_ref: DEVICE
 * double[n] reference values per particle as gorilla_b200_invariants_dev returned them earlier (each may be NULL: that drift
 * is then reported as 0).  Collective: every rank of the communicator must call it. */
int gorilla_b200_diag_reduce_dev(gorilla_b200_handle *h, int64_t n, const double *x, const double *vpar,
                                 const double *vperp, const int32_t *ind_tetr, const double *energy_ref,
                                 const double *p_phi_ref, const double *perpinv_ref, gorilla_diag *out, void *stream);
/* Same with HOST pointers (copies included), for callers whose particle arrays live on the host. */
int gorilla_b200_diag_reduce(gorilla_b200_handle *h, int64_t n, const double *x, const double *vpar, const double *vperp,
                             const int32_t *ind_tetr, const double *energy_ref, const double *p_phi_ref,
                             const double *perpinv_ref, gorilla_diag *out);

/* FP64 issue-rate micro-benchmark on the current device: thread-level instructions per second for DFMA and
 * for DMUL+DADD pairs (the strict build issues the latter).  Roofline denominator of the FP64-bound orders. */
int gorilla_b200_fp64_peak(double *dfma_inst_per_s, double *dmul_dadd_inst_per_s);

/* Neighbour-record prefetch of the push kernels: once the exit face of a push is known the records of the tetrahedron
 * behind it are requested into the L2, overlapping the gather latency with the rest of the push.  It pays when the records
 * a batch touches do not fit the L2 and costs when they do, so it is a run-time option: mode 1 on, 0 off, -1 auto (on when
 * the hot records of the mesh are more than four times the L2 capacity; the default after gorilla_b200_init).  Results are
 * identical either way. */
int gorilla_b200_set_prefetch(gorilla_b200_handle *h, int32_t mode);

/* How the push kernels of order 2 and of the RK4 pusher gather the tetrahedron records: 0 = per-lane vector loads through
 * the L1 (best when the records a batch touches are L2 resident), 1 = per-lane bulk copies of the geometry / magnetic
 * sub-records (cp.async.bulk into a shared-memory slot, issued one push ahead), 2 = warp-cooperative cp.async copies of the 32
 * records a warp needs next into the same kind of slots (four cache lines per instruction; with Phi / strong E the whole
 * record is staged; best on meshes much larger than the L2), -1 = auto by mesh size, field content and pusher (the default
 * after gorilla_b200_init).  Results are identical in every mode. */
int gorilla_b200_set_gather(gorilla_b200_handle *h, int32_t mode);
/* The gather mode in effect (0, 1 or 2; what -1 resolved to). */
int gorilla_b200_get_gather(gorilla_b200_handle *h, int32_t *mode);

/* Tuning knobs (0 = keep default): CTAs per SM and threads per CTA of the persistent push kernel. */
int gorilla_b200_set_launch_config(gorilla_b200_handle *h, int32_t ctas_per_sm, int32_t threads_per_cta);

/* ---- host-side mesh construction (runs once; stays on the host per north_star) ----------------- */

typedef struct gorilla_mesh gorilla_mesh;

/* namelist TETRA_GRID_NML (tetra_grid_settings_mod.f90:70-75) */
typedef struct gorilla_grid_settings {
  int32_t grid_kind;             /* 1 rect/EFIT, 2 field-aligned EFIT, 3 field-aligned VMEC, 4 SOLEDGE3X, 5 analytic */
  int32_t n1, n2, n3;
  int32_t boole_n_field_periods; /* 1 = take from the equilibrium */
  int32_t n_field_periods_manual;
  int32_t i_radial_spacing;
  int32_t theta_geom_flux;
  double sfc_s_min;
  double theta0_at_xpoint;       /* grid_kind 2: non-zero = theta = 0 on the axis -> X-point ray (.true. in the namelist) */
  double R0_analytic_circ, a_analytic_circ, B0_analytic_circ, q0_analytic_circ, q1_analytic_circ;
  const char *g_file_filename;
  const char *convex_wall_filename;
  const char *netcdf_filename;
  const char *knots_SOLEDGE3X_EIRENE_filename;
  const char *triangles_SOLEDGE3X_EIRENE_filename;
  double bmod_multiplier;        /* optional argument of initialize_gorilla / make_tetra_physics (orbit_timestep_gorilla.f90:151,
                                    tetra_physics_mod.f90:281-286): |B| at the vertices is multiplied by it before h = B / |B| and
                                    the records are formed (large values give field-line following); 0 = not given = 1 */
  int32_t nwindow_r, nwindow_z;  /* field_divB0.inp: half-widths of the moving-average filter of the psi(R, Z) table over R and
                                    over Z before it is splined (bdivfree.f90:1144-1164, window_filter
                                    utils_bdivfree.f90:859-872); 0 0 = no filtering, as in every input the reference ships */
} gorilla_grid_settings;

/* make_tetra_grid + make_tetra_physics + check_tetra_overlaps of initialize_gorilla
 * (orbit_timestep_gorilla.f90:151-274; tetra_grid_mod.f90:27-211; tetra_physics_mod.f90:127-1034,1291-1336).
 * Implemented grid kinds: 1 (EFIT g-file, rectangular grid), 2 (EFIT g-file, field aligned, symmetry flux coordinates
 * constructed by field-line integration; theta_geom_flux = 1 flux angle | 2 geometrical angle, points_2d.f90:139-149), 3 (VMEC, field aligned), 4 (SOLEDGE3X-EIRENE triangle mesh
 * extruded toroidally, WEST equilibrium table), 5 (analytic circular tokamak, rectangular grid). */
int gorilla_mesh_build(const gorilla_grid_settings *grid, const gorilla_settings *settings, gorilla_mesh **out);
/* sizeof of the structs of this header as the library was compiled, in the order gorilla_settings, gorilla_mesh_desc,
 * gorilla_counters, gorilla_diag, gorilla_grid_settings, gorilla_event, gorilla_event_settings: a binding in another
 * language (the bind(C) types of the Fortran module, a ctypes mirror) checks its own layout against it once at start-up
 * instead of finding out through a corrupted field. */
#define GORILLA_ABI_N_STRUCTS 7
int gorilla_b200_abi_struct_sizes(int64_t sizes_out[GORILLA_ABI_N_STRUCTS]);
int gorilla_mesh_get_desc(const gorilla_mesh *mesh, gorilla_mesh_desc *out);
/* vertices: nvert, verts_rphiz[nvert][3], verts_sthetaphi[nvert][3] or NULL */
int gorilla_mesh_get_vertices(const gorilla_mesh *mesh, int64_t *nvert, const double **verts_rphiz,
                              const double **verts_sthetaphi);
void gorilla_mesh_free(gorilla_mesh *mesh);

/* .gmesh: versioned on-disk form of a host mesh (header with magic, format version, byte-order tag, record sizes, sizes,
 * the module scalars of gorilla_mesh_desc and an FNV-1a checksum; payload = tetra_physics, tetra_grid, vertex tables as
 * they sit in memory).  The reference has no mesh file -- initialize_gorilla rebuilds the mesh in every run
 * (orbit_timestep_gorilla.f90:151-274); a Fortran caller can dump its own arrays through gorilla_mesh_save and any later
 * run can start from gorilla_mesh_load + gorilla_mesh_get_desc + gorilla_b200_init.  verts_* may be NULL (nvert = 0).
 * Wrong magic / version / byte order / sizes, truncation and corruption give GORILLA_ERR_IO. */
int gorilla_mesh_save(const gorilla_mesh_desc *mesh, int64_t nvert, const double *verts_rphiz,
                      const double *verts_sthetaphi, const char *path);
int gorilla_mesh_load(const char *path, gorilla_mesh **out);

#ifdef __cplusplus
}
#endif
#endif /* GORILLA_B200_H */
