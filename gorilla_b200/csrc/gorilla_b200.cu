// gorilla_b200.cu -- CUDA kernels (sm_100a) and the C ABI of include/gorilla_b200.h.
//
// Kernels:
//   orbit_kernel<K,PHI>   persistent, one particle per lane, lane refill from a global queue:
//                         the whole `do ... enddo` of orbit_timestep_gorilla (orbit_timestep_gorilla.f90:100-139)
//                         runs on the device; a lane whose particle finished or got lost immediately pulls
//                         the next one, so warps stay full until the queue is empty.  K = 1, 2: polynomial orders,
//                         K = 0: RK4; PHI = 0 / 1 / 2: magnetic only / + electrostatic group / + strong-E group.
//   orbit_kernel_g<K,PHI> orders 3 and 4: the same push in 16-warp CTAs whose sub-partition groups run the root-solver
//                         iterations in lock step (instruction-cache sharing), see gb_internal.cuh.
//   find_kernel<PHI>      check_coordinate_domain + find_tetra (binned for the slice-wise grids) for particles that
//                         are not initialised yet.
//   invariants_kernel     energy / p_phi / perpinv per particle.
//   sort keys             radix sort (CUB) of particle indices by tetra index.
// Compile with --fmad=false: the reference ISA has no FMA and the visited-tetra sequence is only
// reproducible with separately rounded multiplies and adds.
#include <string.h>
#include <math.h>
#include <mutex>
#include <vector>
#include "gb_internal.cuh"
#include "gb_repack.hpp"

// ----------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;
static std::atomic<int64_t> g_launch_count{0};
namespace gbint {
void set_error(const char *msg) { g_last_error = msg; }
void count_launch(int n) { g_launch_count += n; }
int fail(int code, const char *msg)
{
  g_last_error = msg;
  return code;
}
}
using gbint::fail;
namespace gbhost {
void set_last_error(const std::string &s) { g_last_error = s; }
}

// ----------------------------------------------------------------------------------------------------
template <int PHI>
__global__ void __launch_bounds__(128) find_kernel(const __grid_constant__ MeshDev m, const Batch bt)
{
  unsigned long long dom_err = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < bt.n; i += (int64_t)gridDim.x * blockDim.x) {
    if (bt.init && bt.init[i]) continue;
    double x[3] = {bt.x[3 * i], bt.x[3 * i + 1], bt.x[3 * i + 2]};
    int32_t it = -1, ifc = -1;
    if (check_coordinate_domain(m, x, bt.boole_periodic_relocation) != 0) {
      dom_err++;
    } else {
      find_tetra<PHI>(&m, x, bt.vpar[i], bt.vperp[i], it, ifc, bt.sign_t_step);
      bt.x[3 * i] = x[0];
      bt.x[3 * i + 1] = x[1];
      bt.x[3 * i + 2] = x[2];
    }
    bt.ind_tetr[i] = it;
    bt.iface[i] = ifc;
    if (bt.init && it != -1) bt.init[i] = 1;
  }
  if (dom_err) atomicAdd(bt.ctr + CTR_DOMAIN, dom_err);
}

// ----------------------------------------------------------------------------------------------------
// FP64 issue-rate micro-benchmark (roofline denominator for the FP64-bound orders; MEASURED_PEAKS.json has
// no FP64 figure).  8 independent dependency chains per thread; MODE 0: DFMA, MODE 1: DMUL + DADD pairs
// (what the strict --fmad=false build issues).
template <int MODE>
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double seed)
{
  double a[8];
#pragma unroll
  for (int k = 0; k < 8; k++) a[k] = seed + 1e-3 * (threadIdx.x + k);
  const double m = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (MODE == 0) a[k] = __fma_rn(a[k], m, c);
      else a[k] = __dadd_rn(__dmul_rn(a[k], m), c);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 8; k++) s += a[k];
  if (s == 123.456) out[0] = s;  // keep the chains alive
}

extern "C" int gorilla_b200_fp64_peak(double *dfma_inst_per_s, double *dmul_dadd_inst_per_s)
{
  int dev = 0, sms = 0;
  GB_CUDA(cudaGetDevice(&dev));
  GB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double *d_out = nullptr;
  GB_CUDA(cudaMalloc((void **)&d_out, sizeof(double)));
  cudaEvent_t e0, e1;
  GB_CUDA(cudaEventCreate(&e0));
  GB_CUDA(cudaEventCreate(&e1));
  const int iters = 1 << 15, grid = sms * 8, block = 256;
  double best[2] = {0.0, 0.0};
  for (int mode = 0; mode < 2; mode++) {
    for (int rep = 0; rep < 4; rep++) {
      GB_CUDA(cudaEventRecord(e0, 0));
      if (mode == 0) fp64_peak_kernel<0><<<grid, block>>>(d_out, iters, 1.0 + rep);
      else fp64_peak_kernel<1><<<grid, block>>>(d_out, iters, 1.0 + rep);
      g_launch_count++;
      GB_CUDA(cudaEventRecord(e1, 0));
      GB_CUDA(cudaEventSynchronize(e1));
      float ms = 0.f;
      GB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      const double inst = (double)grid * block * (double)iters * 8.0 * (mode == 0 ? 1.0 : 2.0);
      const double rate = inst / (ms * 1e-3);
      if (rep > 0 && rate > best[mode]) best[mode] = rate;
    }
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  if (dfma_inst_per_s) *dfma_inst_per_s = best[0];
  if (dmul_dadd_inst_per_s) *dmul_dadd_inst_per_s = best[1];
  return GORILLA_OK;
}

// ----------------------------------------------------------------------------------------------------
__global__ void invariants_kernel(const __grid_constant__ MeshDev m, int64_t n, const double *x, const double *vpar,
                                  const double *vperp, const int32_t *ind_tetr, double *energy, double *p_phi,
                                  double *perpinv_out)
{
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t it = ind_tetr[i];
    double e = NAN, p = NAN, mu = NAN;
    if (it >= 1) particle_invariants(m, it, &x[3 * i], vpar[i], vperp[i], e, p, mu);
    if (energy) energy[i] = e;
    if (p_phi) p_phi[i] = p;
    if (perpinv_out) perpinv_out[i] = mu;
  }
}

// adds the counters of a finished call to the handle's accumulators (gorilla_b200_diag_reduce_dev)
__global__ void accumulate_counters_kernel(const unsigned long long *ctr, unsigned long long *acc)
{
  const int k = threadIdx.x;
  if (k < CTR_N && k != CTR_QUEUE) acc[k] += ctr[k];
}

static int check_settings(const gorilla_settings *s)
{
  if (s->ipusher != 1 && s->ipusher != 2) return fail(GORILLA_ERR_ARG, "ipusher must be 1 (RK4) or 2 (polynomial)");
  if (s->ipusher == 2 && (s->poly_order < 1 || s->poly_order > 4)) return fail(GORILLA_ERR_ARG, "poly_order must be 1..4");
  if (s->ipusher == 1 && !s->boole_dt_dtau) return fail(GORILLA_ERR_UNSUPPORTED, "ipusher = 1 requires boole_dt_dtau = .true.");
  if (s->i_precomp < 0 || s->i_precomp > 2) return fail(GORILLA_ERR_UNSUPPORTED, "i_precomp must be 0, 1 or 2 (3 is not implemented in the reference either)");
  if (s->i_precomp != 0 && s->ipusher == 2) {
    // analytic_integration_with_precomp has no case(1); i_precomp = 2 assigns the coefficients of orders <= 2 only
    if (s->poly_order < 2) return fail(GORILLA_ERR_UNSUPPORTED, "i_precomp = 1, 2 exist for poly_order >= 2");
    if (s->i_precomp == 2 && s->poly_order != 2) return fail(GORILLA_ERR_UNSUPPORTED, "i_precomp = 2 exists for poly_order = 2 only");
    // the step lists that Hamiltonian time / optional quantities / J_par / the adaptive scheme read are only kept by the
    // i_precomp = 0 integration (pusher_tetra_poly.f90:2047-2083)
    if (s->i_time_tracing_option != 1 || s->boole_time_Hamiltonian || s->boole_gyrophase || s->boole_vpar_int ||
        s->boole_vpar2_int || s->boole_adaptive_time_steps)
      return fail(GORILLA_ERR_UNSUPPORTED, "i_precomp = 1, 2 is not combined with Hamiltonian time / optional quantities / adaptive steps");
    if (s->handover_processing_kind != 1) return fail(GORILLA_ERR_UNSUPPORTED, "i_precomp = 1, 2 is not combined with handover_processing_kind = 2");
  }
  if (s->i_time_tracing_option != 1 && s->i_time_tracing_option != 2)
    return fail(GORILLA_ERR_ARG, "i_time_tracing_option must be 1 or 2");
  // gorilla_settings_mod.f90:124-135
  if (s->i_time_tracing_option == 2 && s->ipusher != 2)
    return fail(GORILLA_ERR_ARG, "Hamiltonian time tracing (i_time_tracing_option = 2) requires ipusher = 2");
  if (s->boole_gyrophase && !s->boole_time_Hamiltonian)
    return fail(GORILLA_ERR_ARG, "boole_gyrophase requires boole_time_Hamiltonian = .true.");
  if (s->handover_processing_kind != 1 && s->handover_processing_kind != 2)
    return fail(GORILLA_ERR_ARG, "handover_processing_kind must be 1 or 2");
  if (s->handover_processing_kind == 2 && s->boole_adaptive_time_steps)
    return fail(GORILLA_ERR_UNSUPPORTED, "handover_processing_kind = 2 is not combined with adaptive sub-stepping");
  if (s->boole_adaptive_time_steps) {
    if (s->ipusher != 2) return fail(GORILLA_ERR_ARG, "boole_adaptive_time_steps exists for the polynomial pusher only");
    // pusher_tetra_poly.f90:868-874
    if (!(s->desired_delta_energy > 0.0)) return fail(GORILLA_ERR_ARG, "desired_delta_energy must be > 0");
    if (s->max_n_intermediate_steps < 2) return fail(GORILLA_ERR_ARG, "max_n_intermediate_steps must be >= 2");
    // with Hamiltonian time tracing / optional quantities / events the step lists are kept in full (EXT = 5 kernels)
  }
  // gorilla_settings_mod.f90:139-144 (coord_system is checked against the mesh in gorilla_b200_init)
  if (s->boole_strong_electric_field && (s->i_precomp != 0 || s->boole_newton_precalc))
    return fail(GORILLA_ERR_ARG, "boole_strong_electric_field requires i_precomp = 0 and boole_newton_precalc = .false.");
  if (s->boole_pusher_ode45 && s->ipusher == 1 && !(s->rel_err_ode45 >= 0.0))
    return fail(GORILLA_ERR_ARG, "rel_err_ode45 must be >= 0");
  return GORILLA_OK;
}

static cudaError_t init_slots(gorilla_b200_handle *h)
{
  cudaError_t e;
  for (int k = 0; k < GB_NSLOTS; k++) {
    CallSlot &c = h->slots[k];
    if ((e = cudaMalloc((void **)&c.d_ctr, CTR_N * sizeof(unsigned long long))) != cudaSuccess ||
        (e = cudaEventCreate(&c.ev0)) != cudaSuccess || (e = cudaEventCreate(&c.ev1)) != cudaSuccess ||
        (e = cudaEventCreate(&c.ev2)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&c.done, cudaEventDisableTiming)) != cudaSuccess)
      return e;
  }
  return cudaSuccess;
}
// the next call slot of the ring; waits if the call that used it last is still running
static int acquire_slot(gorilla_b200_handle *h, CallSlot **out)
{
  const int k = (h->cur_slot + 1) % GB_NSLOTS;
  CallSlot &c = h->slots[k];
  if (c.used) GB_CUDA(cudaEventSynchronize(c.done));
  c.used = true;
  c.have_find_time = c.have_push_time = false;
  h->cur_slot = k;
  *out = &c;
  return GORILLA_OK;
}

extern "C" const char *gorilla_b200_last_error(void) { return g_last_error.c_str(); }
extern "C" int64_t gorilla_b200_launch_count(void) { return g_launch_count.load(); }

extern "C" int gorilla_b200_init(const gorilla_mesh_desc *md, const gorilla_settings *st, gorilla_b200_handle **out)
{
  if (!md || !st || !out || !md->tetra_physics || !md->tetra_grid || md->ntetr < 1)
    return fail(GORILLA_ERR_ARG, "gorilla_b200_init: null argument or empty mesh");
  int rc = check_settings(st);
  if (rc) return rc;
  if (md->coord_system != 1 && md->coord_system != 2) return fail(GORILLA_ERR_ARG, "coord_system must be 1 or 2");
  if (st->boole_strong_electric_field && md->coord_system != 1)
    return fail(GORILLA_ERR_ARG, "boole_strong_electric_field requires coord_system = 1 (gorilla_settings_mod.f90:139)");
  if (st->handover_processing_kind == 2 && !md->tetra_skew_coord)
    return fail(GORILLA_ERR_ARG, "handover_processing_kind = 2 needs gorilla_mesh_desc.tetra_skew_coord");
  int dev = 0;
  GB_CUDA(cudaGetDevice(&dev));
  gorilla_b200_handle *h = new gorilla_b200_handle();
  h->device = dev;
  GB_CUDA(cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, dev));
  h->settings = *st;
  {
    int l2 = 0;
    GB_CUDA(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev));
    h->l2_bytes = l2;
  }

  const int64_t nt = md->ntetr;
  std::vector<double> geom, bpart, phi, cold, se;
  bool has_phi = false;
  const bool strong = st->boole_strong_electric_field != 0;
  if (!gb::repack_mesh(md, geom, bpart, phi, cold, has_phi, strong ? &se : nullptr)) {
    delete h;
    return fail(GORILLA_ERR_ARG, "tetra_grid: neighbour_face / perbou value out of range");
  }
  auto up = [&](double **d, const std::vector<double> &v) -> cudaError_t {
    cudaError_t e = cudaMalloc((void **)d, v.size() * sizeof(double));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*d, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice);
  };
  cudaError_t e;
  if ((e = up(&h->d_geom, geom)) != cudaSuccess || (e = up(&h->d_bpart, bpart)) != cudaSuccess ||
      (e = up(&h->d_cold, cold)) != cudaSuccess || ((has_phi || strong) && (e = up(&h->d_phi, phi)) != cudaSuccess) ||
      (strong && (e = up(&h->d_se, se)) != cudaSuccess) ||
      (e = cudaMalloc((void **)&h->d_acc, CTR_N * sizeof(unsigned long long))) != cudaSuccess ||
      (e = cudaMemset(h->d_acc, 0, CTR_N * sizeof(unsigned long long))) != cudaSuccess ||
      (e = cudaMalloc((void **)&h->d_diag, GB_DIAG_ND * sizeof(double))) != cudaSuccess ||
      (e = cudaMallocHost(&h->h_diag, GB_DIAG_ND * sizeof(double))) != cudaSuccess ||
      (e = cudaEventCreateWithFlags(&h->sort_done, cudaEventDisableTiming)) != cudaSuccess ||
      (e = init_slots(h)) != cudaSuccess) {
    g_last_error = std::string("gorilla_b200_init: ") + cudaGetErrorString(e);
    gorilla_b200_free(h);
    return GORILLA_ERR_CUDA;
  }
  // hamiltonian_time records: only the EXT kernels (Hamiltonian time tracing / optional quantities) read them
  h->oq_mask = (st->boole_time_Hamiltonian ? 1u : 0u) | (st->boole_gyrophase ? 2u : 0u) | (st->boole_vpar_int ? 4u : 0u) |
               (st->boole_vpar2_int ? 8u : 0u);
  if (st->ipusher == 2 && (st->i_time_tracing_option == 2 || h->oq_mask)) {
    std::vector<double> ham;
    gb::repack_hamiltonian_time(md, ham);
    if ((e = up(&h->d_ham, ham)) != cudaSuccess) {
      g_last_error = std::string("gorilla_b200_init: ") + cudaGetErrorString(e);
      gorilla_b200_free(h);
      return GORILLA_ERR_CUDA;
    }
  }
  if ((st->ipusher == 2 && st->i_precomp != 0) || (st->ipusher == 1 && st->boole_newton_precalc)) {
    std::vector<double> p4;
    gb::make_precomp_poly4(md, p4);
    if ((e = up(&h->d_poly4, p4)) != cudaSuccess) {
      g_last_error = std::string("gorilla_b200_init: ") + cudaGetErrorString(e);
      gorilla_b200_free(h);
      return GORILLA_ERR_CUDA;
    }
  }
  if (st->handover_processing_kind == 2) {
    std::vector<double> skew;
    gb::repack_skew(md, skew);
    if ((e = up(&h->d_skew, skew)) != cudaSuccess) {
      g_last_error = std::string("gorilla_b200_init: ") + cudaGetErrorString(e);
      gorilla_b200_free(h);
      return GORILLA_ERR_CUDA;
    }
  }
  MeshDev &m = h->mesh;
  m.ntetr = nt;
  h->hot_bytes = (int64_t)nt * 8 * (GEOM_ND + BPART_ND + ((has_phi || strong) ? PHI_ND : 0) + (strong ? SE_ND : 0));
  m.prefetch = 0;   // L2 prefetch of the next record: measured slower on every mesh; gorilla_b200_set_prefetch(1) switches it on
  m.pad_prefetch = 0;
  m.skew = h->d_skew;
  m.ham = h->d_ham;
  m.poly4 = h->d_poly4;
  m.rec44 = nullptr;   // made by gorilla_b200_set_gather
  m.i_precomp = (st->ipusher == 2) ? st->i_precomp : 0;
  m.newton_precalc = (st->ipusher == 1 && st->boole_newton_precalc) ? 1 : 0;
  m.ode45 = (st->ipusher == 1 && st->boole_pusher_ode45) ? 1 : 0;
  m.pad_ode45 = 0;
  m.rel_err_ode45 = st->rel_err_ode45;
  m.time_tracing = st->i_time_tracing_option;
  m.desired_delta_energy = st->desired_delta_energy;
  m.max_n_intermediate_steps = st->max_n_intermediate_steps;
  m.geom = h->d_geom;
  m.bpart = h->d_bpart;
  m.phi = (has_phi || strong) ? h->d_phi : nullptr;
  m.se = strong ? h->d_se : nullptr;
  {  // find_tetra bins for the slice-wise grids (kinds 2, 3, 4); without them find_tetra scans the whole slice
    gb::FindBins fb;
    m.bin_start = m.bin_items = nullptr;
    if (gb::build_find_bins(md, fb)) {
      cudaError_t eb;
      if ((eb = cudaMalloc((void **)&h->d_bin_start, fb.start.size() * sizeof(int32_t))) != cudaSuccess ||
          (eb = cudaMalloc((void **)&h->d_bin_items, (fb.items.size() + 1) * sizeof(int32_t))) != cudaSuccess ||
          (eb = cudaMemcpy(h->d_bin_start, fb.start.data(), fb.start.size() * sizeof(int32_t), cudaMemcpyHostToDevice)) != cudaSuccess ||
          (eb = cudaMemcpy(h->d_bin_items, fb.items.data(), fb.items.size() * sizeof(int32_t), cudaMemcpyHostToDevice)) != cudaSuccess) {
        g_last_error = std::string("gorilla_b200_init: ") + cudaGetErrorString(eb);
        gorilla_b200_free(h);
        return GORILLA_ERR_CUDA;
      }
      m.bin_start = h->d_bin_start; m.bin_items = h->d_bin_items;
      m.bin_nu = fb.nu; m.bin_nv = fb.nv; m.bin_c0 = fb.c0; m.bin_c1 = fb.c1;
      m.bin_u0 = fb.u0; m.bin_v0 = fb.v0; m.bin_du_inv = fb.du_inv; m.bin_dv_inv = fb.dv_inv;
    }
  }
  m.cold = h->d_cold;
  m.cm_over_e = md->cm_over_e;
  m.particle_mass = md->particle_mass;
  m.particle_charge = md->particle_charge;
  const double PI = 3.141592653589793238462643383;
  m.period_phi = 2.0 * PI / md->n_field_periods;
  m.period_theta = 2.0 * PI;
  m.sign_sqg = md->sign_sqg;
  m.coord_system = md->coord_system;
  m.grid_size1 = md->grid_size[0];
  m.grid_size2 = md->grid_size[1];
  m.grid_size3 = md->grid_size[2];
  m.boole_guess = st->boole_guess;
  m.grid_kind = md->grid_kind;
  m.n_field_periods = md->n_field_periods;
  m.Rmin = md->Rmin;
  m.Rmax = md->Rmax;
  m.Zmin = md->Zmin;
  m.Zmax = md->Zmax;
  m.sfc_s_min = md->sfc_s_min;
  if ((rc = gorilla_b200_set_gather(h, -1)) != GORILLA_OK) {   // bulk-copy gather where the mesh is much larger than the L2
    gorilla_b200_free(h);
    return rc;
  }
  *out = h;
  return GORILLA_OK;
}

extern "C" void gorilla_b200_free(gorilla_b200_handle *h)
{
  if (!h) return;
  DeviceGuard dg(h->device);
  gorilla_b200_comm_free(h);
  cudaDeviceSynchronize();
  cudaFree(h->d_geom); cudaFree(h->d_bpart); cudaFree(h->d_phi); cudaFree(h->d_cold); cudaFree(h->d_se); cudaFree(h->d_ham); cudaFree(h->d_skew); cudaFree(h->s_oq); cudaFree(h->d_bin_start); cudaFree(h->d_bin_items);
  cudaFree(h->d_poly4); cudaFree(h->d_rec44); cudaFree(h->d_lst);
  cudaFree(h->d_acc); cudaFree(h->d_diag); cudaFree(h->g_d); cudaFree(h->g_i); cudaFree(h->sort_perm);
  cudaFree(h->s_J); cudaFree(h->s_cv); cudaFree(h->s_cp); cudaFree(h->s_ev); cudaFree(h->s_nev);
  if (h->h_diag) cudaFreeHost(h->h_diag);
  if (h->sort_done) cudaEventDestroy(h->sort_done);
  if (h->lst_done) cudaEventDestroy(h->lst_done);
  for (int k = 0; k < GB_NSLOTS; k++) {
    CallSlot &c = h->slots[k];
    cudaFree(c.d_ctr);
    if (c.ev0) cudaEventDestroy(c.ev0);
    if (c.ev1) cudaEventDestroy(c.ev1);
    if (c.ev2) cudaEventDestroy(c.ev2);
    if (c.done) cudaEventDestroy(c.done);
  }
  cudaFree(h->s_x); cudaFree(h->s_vpar); cudaFree(h->s_vperp); cudaFree(h->s_tro); cudaFree(h->s_e);
  cudaFree(h->s_p); cudaFree(h->s_mu); cudaFree(h->s_init); cudaFree(h->s_ind); cudaFree(h->s_iface);
  cudaFree(h->s_np); cudaFree(h->s_tr_t); cudaFree(h->s_tr_f); cudaFree(h->sort_tmp);
  cudaFree(h->sort_keys_in); cudaFree(h->sort_keys_out); cudaFree(h->sort_vals_in);
  delete h;
}

extern "C" int gorilla_b200_set_launch_config(gorilla_b200_handle *h, int32_t ctas_per_sm, int32_t threads_per_cta)
{
  if (!h) return fail(GORILLA_ERR_ARG, "null handle");
  if (ctas_per_sm > 0) h->ctas_per_sm = ctas_per_sm;
  if (threads_per_cta > 0) {
    if (threads_per_cta % 32 || threads_per_cta > 128) return fail(GORILLA_ERR_ARG, "threads_per_cta must be 32..128, multiple of 32");
    h->threads_per_cta = threads_per_cta;
  }
  return GORILLA_OK;
}
// geom[t][16] + bpart[t][28] (+ phi[t][20] (+ the hot doubles of se[t][32])) -> rec[t][nd]: nd = 44 (bulk-copy gather), 48
// (warp-cooperative gather: records of three 128-byte lines), 64 (with Phi: four lines) or 96 (with the strong-electric-field
// terms: six lines) -- everything a push reads; the doubles behind the last sub-record are padding
__global__ void interleave_rec44_kernel(int64_t ntetr, const double *geom, const double *bpart, const double *phi, const double *se,
                                        double *rec, int nd)
{
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ntetr * nd; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = i / nd;
    int k = (int)(i - t * nd);
    double v = 0.0;
    if (k < GEOM_ND) v = geom[t * GEOM_ND + k];
    else if ((k -= GEOM_ND) < BPART_ND) v = bpart[t * BPART_ND + k];
    else if (nd >= 64 && (k -= BPART_ND) < PHI_ND) v = phi[t * PHI_ND + k];
    else if (nd == 96 && (k -= PHI_ND) < S_HOT_ND) v = se[t * SE_ND + k];
    rec[i] = v;
  }
}
extern "C" int gorilla_b200_set_gather(gorilla_b200_handle *h, int32_t mode)
{
  if (!h || mode < -1 || mode > 2)
    return fail(GORILLA_ERR_ARG, "gorilla_b200_set_gather: mode must be -1 (auto), 0 (loads), 1 (bulk copy) or 2 (warp-cooperative copy)");
  GB_ENTER(h);
  // auto: records staged in shared memory one push ahead when the hot records of the mesh exceed the L2 by a wide margin.
  // Measured, order 2, 3.84 M-tetrahedron EFIT mesh: vector loads 5.56e9, per-lane bulk copies 8.24e9, warp-cooperative copies
  // 1.19e10 crossings/s (RK4: 3.95e9 bulk, 4.23e9 cooperative); 4.24 M-tetrahedron WEST mesh with strong E (the cooperative form
  // stages the Phi / strong-E sub-records as well): order 2 3.88e9 bulk, 5.37e9 cooperative, RK4 2.37e9 / 2.71e9.  On the
  // L2-resident 0.96 M-tetrahedron meshes the vector loads stay ahead (VMEC order 2: 1.42e10 against 1.38e10 cooperative /
  // 0.94e10 bulk).  With Phi alone (whole record staged as well, two CTAs per SM): EFIT mesh, order 2 4.03e9 loads / 7.30e9 bulk /
  // 8.35e9 cooperative, RK4 3.49e9 bulk / 3.42e9 cooperative -- the one case where the bulk copies stay the choice.
  const bool has_bulk_kernel = h->settings.ipusher == 1 || h->settings.poly_order == 2;   // launch_orbit_t: EXT = 0, K = 2 or RK4
  const int phi_kind = h->mesh.se ? 2 : h->mesh.phi ? 1 : 0;
  const int staged = (phi_kind == 1 && h->settings.ipusher == 1) ? 1 : 2;
  const int want = mode >= 0 ? mode : ((h->hot_bytes > 4 * h->l2_bytes && has_bulk_kernel) ? staged : 0);
  const int nd = want == 2 ? coop_nd(phi_kind) : 44;   // phi_kind = the launcher's PHI
  if (want && (!h->d_rec44 || h->rec_nd != nd)) {
    GB_CUDA(cudaDeviceSynchronize());   // a launch may still be reading the other layout
    if (h->d_rec44) GB_CUDA(cudaFree(h->d_rec44));
    h->d_rec44 = nullptr;
    h->mesh.rec44 = nullptr;
    GB_CUDA(cudaMalloc((void **)&h->d_rec44, (size_t)h->mesh.ntetr * nd * sizeof(double)));
    interleave_rec44_kernel<<<h->num_sms * 8, 256>>>(h->mesh.ntetr, h->d_geom, h->d_bpart, h->d_phi, h->d_se, h->d_rec44, nd);
    g_launch_count++;
    GB_CUDA(cudaGetLastError());
    GB_CUDA(cudaDeviceSynchronize());
    h->mesh.rec44 = h->d_rec44;
    h->rec_nd = nd;
  }
  h->bulk_gather = want;
  return GORILLA_OK;
}
extern "C" int gorilla_b200_get_gather(gorilla_b200_handle *h, int32_t *mode)
{
  if (!h || !mode) return fail(GORILLA_ERR_ARG, "gorilla_b200_get_gather: null argument");
  *mode = h->bulk_gather;
  return GORILLA_OK;
}
extern "C" int gorilla_b200_set_prefetch(gorilla_b200_handle *h, int32_t mode)
{
  if (!h || mode < -1 || mode > 1) return fail(GORILLA_ERR_ARG, "gorilla_b200_set_prefetch: mode must be -1 (auto), 0 or 1");
  // auto = off: prefetch.global.L2 of the neighbour's record after the exit face is known was measured slower on the
  // L2-resident meshes (-21 %) and on the DRAM-resident ones (-25 %): the record is needed ~1 us later and the extra
  // request only competes with the demand loads
  h->mesh.prefetch = mode > 0 ? 1 : 0;
  return GORILLA_OK;
}
// test/tuning hook (not in the public header): 0 = one-particle-per-lane kernel of 4-warp CTAs also for orders 3/4
extern "C" int gorilla_b200_debug_use_group(gorilla_b200_handle *h, int32_t on)
{
  if (!h) return GORILLA_ERR_ARG;
  h->use_group = on;
  return GORILLA_OK;
}
// test hook (not in the public header): 0 = find_tetra scans the whole phi slice like the reference, 1 = binned search
extern "C" int gorilla_b200_debug_find_bins(gorilla_b200_handle *h, int32_t on)
{
  if (!h) return GORILLA_ERR_ARG;
  h->mesh.bin_start = (on && h->d_bin_start) ? h->d_bin_start : nullptr;
  return GORILLA_OK;
}
// test hook (not in the public header): route every push through the complete fall-back ladder
extern "C" int gorilla_b200_debug_force_full(gorilla_b200_handle *h, int32_t on)
{
  if (!h) return GORILLA_ERR_ARG;
  h->force_full = on;
  return GORILLA_OK;
}

// the orbit_kernel<K,PHI> / orbit_kernel_g<K,PHI> instantiations live in gb_orbit_k{1..4}.cu and gb_orbit_rk.cu
#define GB_EXTERN_ORBIT(K) \
  extern template int launch_orbit_t<K, 0>(gorilla_b200_handle *, const Batch &, cudaStream_t); \
  extern template int launch_orbit_t<K, 1>(gorilla_b200_handle *, const Batch &, cudaStream_t); \
  extern template int launch_orbit_t<K, 2>(gorilla_b200_handle *, const Batch &, cudaStream_t);
GB_EXTERN_ORBIT(0)
GB_EXTERN_ORBIT(1)
GB_EXTERN_ORBIT(2)
GB_EXTERN_ORBIT(3)
GB_EXTERN_ORBIT(4)
// EXT variants: Hamiltonian time tracing (EXT = 1, gb_orbit_k{1..4}t.cu), + optional quantities / events (EXT = 2,
// gb_orbit_k{1..4}x.cu), adaptive sub-stepping (EXT = 3, gb_orbit_k{1..4}a.cu)
#define GB_EXTERN_ORBIT_X(K, E) \
  extern template int launch_orbit_t<K, 0, E>(gorilla_b200_handle *, const Batch &, cudaStream_t); \
  extern template int launch_orbit_t<K, 1, E>(gorilla_b200_handle *, const Batch &, cudaStream_t); \
  extern template int launch_orbit_t<K, 2, E>(gorilla_b200_handle *, const Batch &, cudaStream_t);
GB_EXTERN_ORBIT_X(1, 1)
GB_EXTERN_ORBIT_X(2, 1)
GB_EXTERN_ORBIT_X(3, 1)
GB_EXTERN_ORBIT_X(4, 1)
GB_EXTERN_ORBIT_X(1, 2)
GB_EXTERN_ORBIT_X(2, 2)
GB_EXTERN_ORBIT_X(3, 2)
GB_EXTERN_ORBIT_X(4, 2)
GB_EXTERN_ORBIT_X(0, 2)   // RK4 with handover_processing_kind = 2 (gb_orbit_rkx.cu)
GB_EXTERN_ORBIT_X(1, 3)
GB_EXTERN_ORBIT_X(2, 3)
GB_EXTERN_ORBIT_X(3, 3)
GB_EXTERN_ORBIT_X(4, 3)
GB_EXTERN_ORBIT_X(1, 5)   // adaptive sub-stepping + list consumers (gb_orbit_k{1..4}ax.cu)
GB_EXTERN_ORBIT_X(2, 5)
GB_EXTERN_ORBIT_X(3, 5)
GB_EXTERN_ORBIT_X(4, 5)
// precomputed-coefficient modes i_precomp = 1, 2 (EXT = 4, gb_orbit_k{2..4}p.cu; no strong-electric-field variant)
#define GB_EXTERN_ORBIT_P(K) \
  extern template int launch_orbit_t<K, 0, 4>(gorilla_b200_handle *, const Batch &, cudaStream_t); \
  extern template int launch_orbit_t<K, 1, 4>(gorilla_b200_handle *, const Batch &, cudaStream_t);
GB_EXTERN_ORBIT_P(2)
GB_EXTERN_ORBIT_P(3)
GB_EXTERN_ORBIT_P(4)

template <int PHI>
static int launch_orbit_k(gorilla_b200_handle *h, const Batch &bt, cudaStream_t s)
{
  // RK4: the kernel with the run-time options carries hand-over kind 2, boole_newton_precalc, ODE45 and the orbit events
  if (h->settings.ipusher == 1)
    return (h->mesh.skew || h->mesh.newton_precalc || h->mesh.ode45 || bt.ev_flags) ? launch_orbit_t<0, PHI, 2>(h, bt, s)
                                                                      : launch_orbit_t<0, PHI>(h, bt, s);
  if (h->settings.boole_adaptive_time_steps && ((bt.optq && bt.oq_mask) || bt.ev_flags || h->mesh.time_tracing == 2)) {
    // the list consumers with the long step lists of the adaptive scheme (hand-over kind 2 and i_precomp stay separate)
    if (h->mesh.skew) return fail(GORILLA_ERR_UNSUPPORTED, "boole_adaptive_time_steps with handover_processing_kind = 2 is not combined with Hamiltonian time / optional quantities / events");
    switch (h->settings.poly_order) {
      case 1: return launch_orbit_t<1, PHI, 5>(h, bt, s);
      case 2: return launch_orbit_t<2, PHI, 5>(h, bt, s);
      case 3: return launch_orbit_t<3, PHI, 5>(h, bt, s);
      default: return launch_orbit_t<4, PHI, 5>(h, bt, s);
    }
  }
  if ((bt.optq && bt.oq_mask) || bt.ev_flags || h->mesh.skew) {   // handover kind 2 lives in the EXT = 2 kernels
    switch (h->settings.poly_order) {
      case 1: return launch_orbit_t<1, PHI, 2>(h, bt, s);
      case 2: return launch_orbit_t<2, PHI, 2>(h, bt, s);
      case 3: return launch_orbit_t<3, PHI, 2>(h, bt, s);
      default: return launch_orbit_t<4, PHI, 2>(h, bt, s);
    }
  }
  if (h->mesh.i_precomp != 0) {   // precomputed coefficients (gb_orbit_k{2..4}p.cu)
    if constexpr (PHI == 2) {
      return fail(GORILLA_ERR_UNSUPPORTED, "i_precomp = 1, 2 is not combined with boole_strong_electric_field");
    } else {
      if (bt.optq || bt.ev_flags) return fail(GORILLA_ERR_UNSUPPORTED, "i_precomp = 1, 2 is not combined with optional quantities / events");
      switch (h->settings.poly_order) {
        case 2: return launch_orbit_t<2, PHI, 4>(h, bt, s);
        case 3: return launch_orbit_t<3, PHI, 4>(h, bt, s);
        default: return launch_orbit_t<4, PHI, 4>(h, bt, s);
      }
    }
  }
  if (h->settings.boole_adaptive_time_steps) {   // adaptive sub-stepping (gb_orbit_k{1..4}a.cu)
    switch (h->settings.poly_order) {
      case 1: return launch_orbit_t<1, PHI, 3>(h, bt, s);
      case 2: return launch_orbit_t<2, PHI, 3>(h, bt, s);
      case 3: return launch_orbit_t<3, PHI, 3>(h, bt, s);
      default: return launch_orbit_t<4, PHI, 3>(h, bt, s);
    }
  }
  if (h->mesh.time_tracing == 2) {
    switch (h->settings.poly_order) {
      case 1: return launch_orbit_t<1, PHI, 1>(h, bt, s);
      case 2: return launch_orbit_t<2, PHI, 1>(h, bt, s);
      case 3: return launch_orbit_t<3, PHI, 1>(h, bt, s);
      default: return launch_orbit_t<4, PHI, 1>(h, bt, s);
    }
  }
  switch (h->settings.poly_order) {
    case 1: return launch_orbit_t<1, PHI>(h, bt, s);
    case 2: return launch_orbit_t<2, PHI>(h, bt, s);
    case 3: return launch_orbit_t<3, PHI>(h, bt, s);
    default: return launch_orbit_t<4, PHI>(h, bt, s);
  }
}

static int launch_find(gorilla_b200_handle *h, const Batch &bt, cudaStream_t s)
{
  int64_t grid = (bt.n + 127) / 128;
  if (grid > (int64_t)h->num_sms * 16) grid = (int64_t)h->num_sms * 16;
  if (h->mesh.se) find_kernel<2><<<(unsigned)grid, 128, 0, s>>>(h->mesh, bt);
  else if (h->mesh.phi) find_kernel<1><<<(unsigned)grid, 128, 0, s>>>(h->mesh, bt);
  else find_kernel<0><<<(unsigned)grid, 128, 0, s>>>(h->mesh, bt);
  g_launch_count++;
  GB_CUDA(cudaGetLastError());
  return GORILLA_OK;
}

// One batched call on stream s: [find_tetra of the particles that are not localised yet] -> [optional re-sort by
// tetrahedron: resort_perm != nullptr gathers the six state arrays through a fresh permutation before the push and
// scatters them (and t_remain_out / n_pushes) back afterwards] -> push kernel -> counters into the accumulators.
static int run_device(gorilla_b200_handle *h, Batch bt, bool do_find, cudaStream_t s, bool resort = false)
{
  CallSlot *slot = nullptr;
  int rc = acquire_slot(h, &slot);
  if (rc) return rc;
  bt.ctr = slot->d_ctr;
  bt.boole_periodic_relocation = h->settings.boole_periodic_relocation;
  bt.sign_t_step = signbit(bt.t_step) ? -1 : 1;
  bt.force_full = h->force_full;
  bt.rebin = h->use_group == 2 ? 1 : 0;
  bt.oq_mask = bt.optq ? h->oq_mask : 0u;
  if (bt.optq && !bt.oq_mask) {  // nothing switched on: all zero, the plain kernel runs
    GB_CUDA(cudaMemsetAsync(bt.optq, 0, (size_t)bt.n * 4 * sizeof(double), s));
    bt.optq = nullptr;
  }
  GB_CUDA(cudaMemsetAsync(slot->d_ctr, 0, CTR_N * sizeof(unsigned long long), s));
  slot->n = bt.n;
  if (bt.n == 0) {
    GB_CUDA(cudaEventRecord(slot->done, s));
    return GORILLA_OK;
  }
  GB_CUDA(cudaEventRecord(slot->ev0, s));
  if (do_find) {
    rc = launch_find(h, bt, s);
    if (rc) return rc;
    slot->have_find_time = true;
  }
  if (resort) {
    rc = gbint::sort_permutation(h, bt.n, bt.ind_tetr, nullptr, s);
    if (!rc) rc = gbint::permute_state_inplace(h, bt.n, h->sort_perm, false, bt.x, bt.vpar, bt.vperp, bt.init, bt.ind_tetr, bt.iface, s);
    if (rc) return rc;
  }
  GB_CUDA(cudaEventRecord(slot->ev1, s));
  rc = h->mesh.se ? launch_orbit_k<2>(h, bt, s) : h->mesh.phi ? launch_orbit_k<1>(h, bt, s) : launch_orbit_k<0>(h, bt, s);
  if (rc) return rc;
  GB_CUDA(cudaEventRecord(slot->ev2, s));
  slot->have_push_time = true;
  if (resort) {
    rc = gbint::permute_state_inplace(h, bt.n, h->sort_perm, true, bt.x, bt.vpar, bt.vperp, bt.init, bt.ind_tetr, bt.iface, s);
    if (!rc && bt.t_remain_out) rc = gbint::permute_one_inplace<double>(h, bt.n, h->sort_perm, true, bt.t_remain_out, s);
    if (!rc && bt.n_pushes) rc = gbint::permute_one_inplace<int64_t>(h, bt.n, h->sort_perm, true, bt.n_pushes, s);
    if (rc) return rc;
  }
  accumulate_counters_kernel<<<1, 32, 0, s>>>(slot->d_ctr, h->d_acc);
  g_launch_count++;
  GB_CUDA(cudaGetLastError());
  GB_CUDA(cudaEventRecord(slot->done, s));
  return GORILLA_OK;
}

extern "C" int gorilla_b200_orbit_timestep_dev(gorilla_b200_handle *h, int64_t n, double *x, double *vpar, double *vperp,
                                               double t_step, int32_t *boole_initialized, int32_t *ind_tetr,
                                               int32_t *iface, double *t_remain_out, int64_t *n_pushes, void *stream)
{
  if (!h || n < 0 || (n > 0 && (!x || !vpar || !vperp || !boole_initialized || !ind_tetr || !iface)))
    return fail(GORILLA_ERR_ARG, "gorilla_b200_orbit_timestep_dev: null argument");
  GB_ENTER(h);
  Batch bt{};
  bt.n = n; bt.x = x; bt.vpar = vpar; bt.vperp = vperp; bt.t_step = t_step; bt.init = boole_initialized;
  bt.ind_tetr = ind_tetr; bt.iface = iface; bt.t_remain_out = t_remain_out; bt.n_pushes = n_pushes;
  return run_device(h, bt, true, (cudaStream_t)stream);
}

extern "C" int gorilla_b200_orbit_timestep_optional_dev(gorilla_b200_handle *h, int64_t n, double *x, double *vpar,
                                                        double *vperp, double t_step, int32_t *boole_initialized,
                                                        int32_t *ind_tetr, int32_t *iface, double *t_remain_out,
                                                        int64_t *n_pushes, double *optional_quantities, void *stream)
{
  if (!h || n < 0 || (n > 0 && (!x || !vpar || !vperp || !boole_initialized || !ind_tetr || !iface || !optional_quantities)))
    return fail(GORILLA_ERR_ARG, "gorilla_b200_orbit_timestep_optional_dev: null argument");
  if (h->settings.ipusher != 2)
    return fail(GORILLA_ERR_UNSUPPORTED, "optional quantities exist for the polynomial pusher only (ipusher = 2)");
  GB_ENTER(h);
  Batch bt{};
  bt.n = n; bt.x = x; bt.vpar = vpar; bt.vperp = vperp; bt.t_step = t_step; bt.init = boole_initialized;
  bt.ind_tetr = ind_tetr; bt.iface = iface; bt.t_remain_out = t_remain_out; bt.n_pushes = n_pushes;
  bt.optq = optional_quantities;
  return run_device(h, bt, true, (cudaStream_t)stream);
}

static int ensure_scratch(gorilla_b200_handle *h, int64_t n);
namespace gbint {
int ensure_host_scratch(gorilla_b200_handle *h, int64_t n) { return ensure_scratch(h, n); }
}
static int check_event_args(gorilla_b200_handle *h, const gorilla_event_settings *cfg, int &flags)
{
  if (h->settings.ipusher == 2 && h->settings.poly_order < 2)
    return fail(GORILLA_ERR_UNSUPPORTED, "orbit events need the polynomial pusher of order 2..4 (par_adiab_inv_poly_mod) or the RK pusher (par_adiab_inv_rk_mod)");
  flags = (cfg->boole_poincare_phi_0 ? 1 : 0) | (cfg->boole_poincare_vpar_0 ? 2 : 0) | (cfg->boole_J_par ? 4 : 0) |
          (cfg->boole_full_orbit ? 8 : 0);
  if ((flags & 8) && cfg->n_skip_full_orbit < 1) return fail(GORILLA_ERR_ARG, "n_skip_full_orbit must be >= 1");
  if ((flags & 1) && cfg->n_skip_phi_0 < 1) return fail(GORILLA_ERR_ARG, "n_skip_phi_0 must be >= 1");
  if ((flags & 6) && cfg->n_skip_vpar_0 < 1) return fail(GORILLA_ERR_ARG, "n_skip_vpar_0 must be >= 1");
  if (!flags) return fail(GORILLA_ERR_ARG, "no event kind switched on");
  return GORILLA_OK;
}
extern "C" int gorilla_b200_orbit_timestep_events_dev(gorilla_b200_handle *h, int64_t n, double *x, double *vpar,
                                                      double *vperp, double t_step, int32_t *boole_initialized,
                                                      int32_t *ind_tetr, int32_t *iface, double *t_remain_out,
                                                      int64_t *n_pushes, const gorilla_event_settings *cfg,
                                                      double *par_adiab_inv, int32_t *counter_vpar_0,
                                                      int32_t *counter_phi_0, gorilla_event *events, int64_t event_cap,
                                                      uint64_t *n_events, void *stream)
{
  if (!h || !cfg || n < 0 || event_cap < 0 || !n_events || (event_cap > 0 && !events) ||
      (n > 0 && (!x || !vpar || !vperp || !boole_initialized || !ind_tetr || !iface || !par_adiab_inv || !counter_vpar_0 ||
                 !counter_phi_0)))
    return fail(GORILLA_ERR_ARG, "gorilla_b200_orbit_timestep_events_dev: null argument");
  int flags = 0;
  int rc = check_event_args(h, cfg, flags);
  if (rc) return rc;
  GB_ENTER(h);
  Batch bt{};
  bt.n = n; bt.x = x; bt.vpar = vpar; bt.vperp = vperp; bt.t_step = t_step; bt.init = boole_initialized;
  bt.ind_tetr = ind_tetr; bt.iface = iface; bt.t_remain_out = t_remain_out; bt.n_pushes = n_pushes;
  bt.ev_flags = flags; bt.n_skip_phi_0 = cfg->n_skip_phi_0 > 0 ? cfg->n_skip_phi_0 : 1;
  bt.n_skip_vpar_0 = cfg->n_skip_vpar_0 > 0 ? cfg->n_skip_vpar_0 : 1;
  bt.n_skip_full_orbit = cfg->n_skip_full_orbit > 0 ? cfg->n_skip_full_orbit : 1;
  bt.par_adiab_inv = par_adiab_inv; bt.counter_vpar_0 = counter_vpar_0; bt.counter_phi_0 = counter_phi_0;
  bt.events = events; bt.ev_cap = event_cap; bt.ev_count = (unsigned long long *)n_events;
  return run_device(h, bt, true, (cudaStream_t)stream);
}

extern "C" int gorilla_b200_orbit_timestep_events(gorilla_b200_handle *h, int64_t n, double *x, double *vpar, double *vperp,
                                                  double t_step, int32_t *boole_initialized, int32_t *ind_tetr,
                                                  int32_t *iface, double *t_remain_out, int64_t *n_pushes,
                                                  const gorilla_event_settings *cfg, double *par_adiab_inv,
                                                  int32_t *counter_vpar_0, int32_t *counter_phi_0, gorilla_event *events,
                                                  int64_t event_cap, int64_t *n_events)
{
  if (!h || !cfg || n < 0 || event_cap < 0 || !n_events || (event_cap > 0 && !events) ||
      (n > 0 && (!x || !vpar || !vperp || !boole_initialized || !ind_tetr || !iface || !par_adiab_inv || !counter_vpar_0 ||
                 !counter_phi_0)))
    return fail(GORILLA_ERR_ARG, "gorilla_b200_orbit_timestep_events: null argument");
  int flags = 0;
  int rc = check_event_args(h, cfg, flags);
  if (rc) return rc;
  *n_events = 0;
  if (n == 0) return GORILLA_OK;
  GB_ENTER(h);
  // persistent scratch on the handle (grown on demand), no allocation per call
  rc = ensure_scratch(h, n);
  if (rc) return rc;
  if (n > h->ev_state_cap) {
    cudaFree(h->s_J); cudaFree(h->s_cv); cudaFree(h->s_cp);
    h->s_J = nullptr; h->s_cv = h->s_cp = nullptr; h->ev_state_cap = 0;
    GB_CUDA(cudaMalloc((void **)&h->s_J, (size_t)n * sizeof(double)));
    GB_CUDA(cudaMalloc((void **)&h->s_cv, (size_t)n * sizeof(int32_t)));
    GB_CUDA(cudaMalloc((void **)&h->s_cp, (size_t)n * sizeof(int32_t)));
    h->ev_state_cap = n;
  }
  if (event_cap > h->ev_cap || !h->s_ev) {
    cudaFree(h->s_ev);
    h->s_ev = nullptr; h->ev_cap = 0;
    GB_CUDA(cudaMalloc((void **)&h->s_ev, (size_t)(event_cap > 0 ? event_cap : 1) * sizeof(gorilla_event)));
    h->ev_cap = event_cap > 0 ? event_cap : 1;
  }
  if (!h->s_nev) GB_CUDA(cudaMalloc((void **)&h->s_nev, sizeof(uint64_t)));
  cudaStream_t s = nullptr;
  const size_t nd = (size_t)n * sizeof(double), ni = (size_t)n * sizeof(int32_t);
  GB_CUDA(cudaMemcpyAsync(h->s_x, x, 3 * nd, cudaMemcpyHostToDevice, s));
  GB_CUDA(cudaMemcpyAsync(h->s_vpar, vpar, nd, cudaMemcpyHostToDevice, s));
  GB_CUDA(cudaMemcpyAsync(h->s_vperp, vperp, nd, cudaMemcpyHostToDevice, s));
  GB_CUDA(cudaMemcpyAsync(h->s_J, par_adiab_inv, nd, cudaMemcpyHostToDevice, s));
  GB_CUDA(cudaMemcpyAsync(h->s_init, boole_initialized, ni, cudaMemcpyHostToDevice, s));
  GB_CUDA(cudaMemcpyAsync(h->s_ind, ind_tetr, ni, cudaMemcpyHostToDevice, s));
  GB_CUDA(cudaMemcpyAsync(h->s_iface, iface, ni, cudaMemcpyHostToDevice, s));
  GB_CUDA(cudaMemcpyAsync(h->s_cv, counter_vpar_0, ni, cudaMemcpyHostToDevice, s));
  GB_CUDA(cudaMemcpyAsync(h->s_cp, counter_phi_0, ni, cudaMemcpyHostToDevice, s));
  GB_CUDA(cudaMemsetAsync(h->s_nev, 0, sizeof(uint64_t), s));
  rc = gorilla_b200_orbit_timestep_events_dev(h, n, h->s_x, h->s_vpar, h->s_vperp, t_step, h->s_init, h->s_ind, h->s_iface,
                                              h->s_tro, h->s_np, cfg, h->s_J, h->s_cv, h->s_cp, h->s_ev, event_cap, h->s_nev, s);
  if (rc) return rc;
  uint64_t nev = 0;
  GB_CUDA(cudaMemcpyAsync(&nev, h->s_nev, sizeof(nev), cudaMemcpyDeviceToHost, s));
  GB_CUDA(cudaMemcpyAsync(x, h->s_x, 3 * nd, cudaMemcpyDeviceToHost, s));
  GB_CUDA(cudaMemcpyAsync(vpar, h->s_vpar, nd, cudaMemcpyDeviceToHost, s));
  GB_CUDA(cudaMemcpyAsync(vperp, h->s_vperp, nd, cudaMemcpyDeviceToHost, s));
  GB_CUDA(cudaMemcpyAsync(par_adiab_inv, h->s_J, nd, cudaMemcpyDeviceToHost, s));
  GB_CUDA(cudaMemcpyAsync(boole_initialized, h->s_init, ni, cudaMemcpyDeviceToHost, s));
  GB_CUDA(cudaMemcpyAsync(ind_tetr, h->s_ind, ni, cudaMemcpyDeviceToHost, s));
  GB_CUDA(cudaMemcpyAsync(iface, h->s_iface, ni, cudaMemcpyDeviceToHost, s));
  GB_CUDA(cudaMemcpyAsync(counter_vpar_0, h->s_cv, ni, cudaMemcpyDeviceToHost, s));
  GB_CUDA(cudaMemcpyAsync(counter_phi_0, h->s_cp, ni, cudaMemcpyDeviceToHost, s));
  if (t_remain_out) GB_CUDA(cudaMemcpyAsync(t_remain_out, h->s_tro, nd, cudaMemcpyDeviceToHost, s));
  if (n_pushes) GB_CUDA(cudaMemcpyAsync(n_pushes, h->s_np, (size_t)n * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  GB_CUDA(cudaStreamSynchronize(s));
  const int64_t stored = (int64_t)nev < event_cap ? (int64_t)nev : event_cap;
  if (stored > 0) GB_CUDA(cudaMemcpy(events, h->s_ev, (size_t)stored * sizeof(gorilla_event), cudaMemcpyDeviceToHost));
  *n_events = (int64_t)nev;
  unsigned long long dom = 0;
  GB_CUDA(cudaMemcpy(&dom, h->slots[h->cur_slot].d_ctr + CTR_DOMAIN, sizeof(dom), cudaMemcpyDeviceToHost));
  if (dom) return fail(GORILLA_ERR_DOMAIN, "particle start position outside the computation domain");
  return GORILLA_OK;
}

static int ensure_scratch(gorilla_b200_handle *h, int64_t n)
{
  if (n <= h->cap) return GORILLA_OK;
  cudaFree(h->s_oq);
  h->s_oq = nullptr;
  cudaFree(h->s_x); cudaFree(h->s_vpar); cudaFree(h->s_vperp); cudaFree(h->s_tro); cudaFree(h->s_e);
  cudaFree(h->s_p); cudaFree(h->s_mu); cudaFree(h->s_init); cudaFree(h->s_ind); cudaFree(h->s_iface); cudaFree(h->s_np);
  h->s_x = h->s_vpar = h->s_vperp = h->s_tro = h->s_e = h->s_p = h->s_mu = nullptr;
  h->s_init = h->s_ind = h->s_iface = nullptr; h->s_np = nullptr;
  h->cap = 0;
  GB_CUDA(cudaMalloc((void **)&h->s_x, (size_t)n * 3 * sizeof(double)));
  GB_CUDA(cudaMalloc((void **)&h->s_vpar, (size_t)n * sizeof(double)));
  GB_CUDA(cudaMalloc((void **)&h->s_vperp, (size_t)n * sizeof(double)));
  GB_CUDA(cudaMalloc((void **)&h->s_tro, (size_t)n * sizeof(double)));
  GB_CUDA(cudaMalloc((void **)&h->s_e, (size_t)n * sizeof(double)));
  GB_CUDA(cudaMalloc((void **)&h->s_p, (size_t)n * sizeof(double)));
  GB_CUDA(cudaMalloc((void **)&h->s_mu, (size_t)n * sizeof(double)));
  GB_CUDA(cudaMalloc((void **)&h->s_init, (size_t)n * sizeof(int32_t)));
  GB_CUDA(cudaMalloc((void **)&h->s_ind, (size_t)n * sizeof(int32_t)));
  GB_CUDA(cudaMalloc((void **)&h->s_iface, (size_t)n * sizeof(int32_t)));
  GB_CUDA(cudaMalloc((void **)&h->s_np, (size_t)n * sizeof(int64_t)));
  GB_CUDA(cudaMalloc((void **)&h->s_oq, (size_t)n * 4 * sizeof(double)));
  h->cap = n;
  return GORILLA_OK;
}

static int orbit_host(gorilla_b200_handle *h, int64_t n, double *x, double *vpar, double *vperp, double t_step,
                      int32_t *binit, int32_t *ind_tetr, int32_t *iface, double *tro, int64_t *np, int32_t trace_cap,
                      int32_t *tr_t, int32_t *tr_f, double *optq = nullptr)
{
  if (!h || n < 0 || (n > 0 && (!x || !vpar || !vperp || !binit || !ind_tetr || !iface)))
    return fail(GORILLA_ERR_ARG, "gorilla_b200_orbit_timestep: null argument");
  if (n == 0) return GORILLA_OK;
  GB_ENTER(h);
  int rc = ensure_scratch(h, n);
  if (rc) return rc;
  cudaStream_t s = nullptr;
  GB_CUDA(cudaMemcpyAsync(h->s_x, x, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, s));
  GB_CUDA(cudaMemcpyAsync(h->s_vpar, vpar, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
  GB_CUDA(cudaMemcpyAsync(h->s_vperp, vperp, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
  GB_CUDA(cudaMemcpyAsync(h->s_init, binit, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  GB_CUDA(cudaMemcpyAsync(h->s_ind, ind_tetr, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  GB_CUDA(cudaMemcpyAsync(h->s_iface, iface, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  Batch bt{};
  bt.n = n; bt.x = h->s_x; bt.vpar = h->s_vpar; bt.vperp = h->s_vperp; bt.t_step = t_step; bt.init = h->s_init;
  bt.ind_tetr = h->s_ind; bt.iface = h->s_iface; bt.t_remain_out = h->s_tro; bt.n_pushes = h->s_np;
  if (trace_cap > 0) {
    if (!tr_t || !tr_f) return fail(GORILLA_ERR_ARG, "trace buffers are null");
    const int64_t elems = n * (int64_t)trace_cap;
    if (elems > h->trace_cap_elems) {
      cudaFree(h->s_tr_t); cudaFree(h->s_tr_f);
      h->s_tr_t = h->s_tr_f = nullptr; h->trace_cap_elems = 0;
      GB_CUDA(cudaMalloc((void **)&h->s_tr_t, (size_t)elems * sizeof(int32_t)));
      GB_CUDA(cudaMalloc((void **)&h->s_tr_f, (size_t)elems * sizeof(int32_t)));
      h->trace_cap_elems = elems;
    }
    GB_CUDA(cudaMemsetAsync(h->s_tr_t, 0, (size_t)elems * sizeof(int32_t), s));
    GB_CUDA(cudaMemsetAsync(h->s_tr_f, 0, (size_t)elems * sizeof(int32_t), s));
    bt.trace_cap = trace_cap; bt.trace_tetr = h->s_tr_t; bt.trace_face = h->s_tr_f;
  }
  if (optq) bt.optq = h->s_oq;
  // in-library re-sort (gorilla_b200_set_host_resort): gather locality for callers that only have host arrays
  const bool resort = h->host_resort && trace_cap <= 0 && !optq && t_step != 0.0 && n >= 4096;
  rc = run_device(h, bt, true, s, resort);
  if (rc) return rc;
  if (optq) GB_CUDA(cudaMemcpyAsync(optq, h->s_oq, (size_t)n * 4 * sizeof(double), cudaMemcpyDeviceToHost, s));
  GB_CUDA(cudaMemcpyAsync(x, h->s_x, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
  GB_CUDA(cudaMemcpyAsync(vpar, h->s_vpar, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s));
  GB_CUDA(cudaMemcpyAsync(vperp, h->s_vperp, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s));
  GB_CUDA(cudaMemcpyAsync(binit, h->s_init, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  GB_CUDA(cudaMemcpyAsync(ind_tetr, h->s_ind, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  GB_CUDA(cudaMemcpyAsync(iface, h->s_iface, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  if (tro) GB_CUDA(cudaMemcpyAsync(tro, h->s_tro, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (np) GB_CUDA(cudaMemcpyAsync(np, h->s_np, (size_t)n * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  if (trace_cap > 0) {
    GB_CUDA(cudaMemcpyAsync(tr_t, h->s_tr_t, (size_t)n * trace_cap * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    GB_CUDA(cudaMemcpyAsync(tr_f, h->s_tr_f, (size_t)n * trace_cap * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  }
  GB_CUDA(cudaStreamSynchronize(s));
  unsigned long long dom = 0;
  GB_CUDA(cudaMemcpy(&dom, h->slots[h->cur_slot].d_ctr + CTR_DOMAIN, sizeof(dom), cudaMemcpyDeviceToHost));
  if (dom) return fail(GORILLA_ERR_DOMAIN, "particle start position outside the computation domain");
  return GORILLA_OK;
}

extern "C" int gorilla_b200_orbit_timestep(gorilla_b200_handle *h, int64_t n, double *x, double *vpar, double *vperp,
                                           double t_step, int32_t *boole_initialized, int32_t *ind_tetr, int32_t *iface,
                                           double *t_remain_out, int64_t *n_pushes)
{
  return orbit_host(h, n, x, vpar, vperp, t_step, boole_initialized, ind_tetr, iface, t_remain_out, n_pushes, 0, nullptr,
                    nullptr);
}
extern "C" int gorilla_b200_orbit_timestep_trace(gorilla_b200_handle *h, int64_t n, double *x, double *vpar,
                                                 double *vperp, double t_step, int32_t *boole_initialized,
                                                 int32_t *ind_tetr, int32_t *iface, double *t_remain_out,
                                                 int64_t *n_pushes, int32_t trace_cap, int32_t *trace_ind_tetr,
                                                 int32_t *trace_iface)
{
  return orbit_host(h, n, x, vpar, vperp, t_step, boole_initialized, ind_tetr, iface, t_remain_out, n_pushes, trace_cap,
                    trace_ind_tetr, trace_iface);
}

extern "C" int gorilla_b200_orbit_timestep_optional(gorilla_b200_handle *h, int64_t n, double *x, double *vpar,
                                                    double *vperp, double t_step, int32_t *boole_initialized,
                                                    int32_t *ind_tetr, int32_t *iface, double *t_remain_out,
                                                    int64_t *n_pushes, double *optional_quantities)
{
  if (!optional_quantities) return fail(GORILLA_ERR_ARG, "gorilla_b200_orbit_timestep_optional: null argument");
  if (h && h->settings.ipusher != 2)
    return fail(GORILLA_ERR_UNSUPPORTED, "optional quantities exist for the polynomial pusher only (ipusher = 2)");
  return orbit_host(h, n, x, vpar, vperp, t_step, boole_initialized, ind_tetr, iface, t_remain_out, n_pushes, 0, nullptr,
                    nullptr, optional_quantities);
}
// test hook (not in the public header): trace + optional quantities in one call
extern "C" int gorilla_b200_debug_orbit_timestep_trace_optional(gorilla_b200_handle *h, int64_t n, double *x, double *vpar,
                                                                double *vperp, double t_step, int32_t *boole_initialized,
                                                                int32_t *ind_tetr, int32_t *iface, double *t_remain_out,
                                                                int64_t *n_pushes, int32_t trace_cap, int32_t *trace_ind_tetr,
                                                                int32_t *trace_iface, double *optional_quantities)
{
  return orbit_host(h, n, x, vpar, vperp, t_step, boole_initialized, ind_tetr, iface, t_remain_out, n_pushes, trace_cap,
                    trace_ind_tetr, trace_iface, optional_quantities);
}

extern "C" int gorilla_b200_find_tetra(gorilla_b200_handle *h, int64_t n, double *x, const double *vpar,
                                       const double *vperp, int32_t *ind_tetr, int32_t *iface, int32_t sign_t_step)
{
  if (!h || n < 0 || (n > 0 && (!x || !vpar || !vperp || !ind_tetr || !iface)))
    return fail(GORILLA_ERR_ARG, "gorilla_b200_find_tetra: null argument");
  if (n == 0) return GORILLA_OK;
  GB_ENTER(h);
  int rc = ensure_scratch(h, n);
  if (rc) return rc;
  cudaStream_t s = nullptr;
  CallSlot *slot = nullptr;
  rc = acquire_slot(h, &slot);
  if (rc) return rc;
  slot->n = n;
  GB_CUDA(cudaMemcpyAsync(h->s_x, x, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, s));
  GB_CUDA(cudaMemcpyAsync(h->s_vpar, vpar, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
  GB_CUDA(cudaMemcpyAsync(h->s_vperp, vperp, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
  GB_CUDA(cudaMemsetAsync(slot->d_ctr, 0, CTR_N * sizeof(unsigned long long), s));
  Batch bt{};
  bt.n = n; bt.x = h->s_x; bt.vpar = h->s_vpar; bt.vperp = h->s_vperp; bt.init = nullptr; bt.ind_tetr = h->s_ind;
  bt.iface = h->s_iface; bt.ctr = slot->d_ctr; bt.boole_periodic_relocation = h->settings.boole_periodic_relocation;
  bt.sign_t_step = sign_t_step < 0 ? -1 : 1;
  rc = launch_find(h, bt, s);
  if (rc) return rc;
  GB_CUDA(cudaEventRecord(slot->done, s));
  GB_CUDA(cudaMemcpyAsync(x, h->s_x, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
  GB_CUDA(cudaMemcpyAsync(ind_tetr, h->s_ind, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  GB_CUDA(cudaMemcpyAsync(iface, h->s_iface, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  GB_CUDA(cudaStreamSynchronize(s));
  unsigned long long dom = 0;
  GB_CUDA(cudaMemcpy(&dom, slot->d_ctr + CTR_DOMAIN, sizeof(dom), cudaMemcpyDeviceToHost));
  if (dom) return fail(GORILLA_ERR_DOMAIN, "particle start position outside the computation domain");
  return GORILLA_OK;
}

extern "C" int gorilla_b200_invariants_dev(gorilla_b200_handle *h, int64_t n, const double *x, const double *vpar,
                                           const double *vperp, const int32_t *ind_tetr, double *energy, double *p_phi,
                                           double *perpinv, void *stream)
{
  if (!h || n < 0 || (n > 0 && (!x || !vpar || !vperp || !ind_tetr))) return fail(GORILLA_ERR_ARG, "invariants: null argument");
  if (n == 0) return GORILLA_OK;
  GB_ENTER(h);
  int64_t grid = (n + 255) / 256;
  if (grid > (int64_t)h->num_sms * 8) grid = (int64_t)h->num_sms * 8;
  invariants_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(h->mesh, n, x, vpar, vperp, ind_tetr, energy, p_phi,
                                                                       perpinv);
  g_launch_count++;
  GB_CUDA(cudaGetLastError());
  return GORILLA_OK;
}
extern "C" int gorilla_b200_invariants(gorilla_b200_handle *h, int64_t n, const double *x, const double *vpar,
                                       const double *vperp, const int32_t *ind_tetr, double *energy, double *p_phi,
                                       double *perpinv)
{
  if (!h || n < 0 || (n > 0 && (!x || !vpar || !vperp || !ind_tetr))) return fail(GORILLA_ERR_ARG, "invariants: null argument");
  if (n == 0) return GORILLA_OK;
  GB_ENTER(h);
  int rc = ensure_scratch(h, n);
  if (rc) return rc;
  cudaStream_t s = nullptr;
  GB_CUDA(cudaMemcpyAsync(h->s_x, x, (size_t)n * 3 * sizeof(double), cudaMemcpyHostToDevice, s));
  GB_CUDA(cudaMemcpyAsync(h->s_vpar, vpar, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
  GB_CUDA(cudaMemcpyAsync(h->s_vperp, vperp, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s));
  GB_CUDA(cudaMemcpyAsync(h->s_ind, ind_tetr, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
  rc = gorilla_b200_invariants_dev(h, n, h->s_x, h->s_vpar, h->s_vperp, h->s_ind, h->s_e, h->s_p, h->s_mu, s);
  if (rc) return rc;
  if (energy) GB_CUDA(cudaMemcpyAsync(energy, h->s_e, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (p_phi) GB_CUDA(cudaMemcpyAsync(p_phi, h->s_p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (perpinv) GB_CUDA(cudaMemcpyAsync(perpinv, h->s_mu, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s));
  GB_CUDA(cudaStreamSynchronize(s));
  return GORILLA_OK;
}

extern "C" int gorilla_b200_get_counters(gorilla_b200_handle *h, gorilla_counters *out)
{
  if (!h || !out) return fail(GORILLA_ERR_ARG, "get_counters: null argument");
  memset(out, 0, sizeof(*out));
  if (h->cur_slot < 0) return GORILLA_OK;
  GB_ENTER(h);
  CallSlot &slot = h->slots[h->cur_slot];
  GB_CUDA(cudaEventSynchronize(slot.done));
  unsigned long long c[CTR_N];
  GB_CUDA(cudaMemcpy(c, slot.d_ctr, sizeof(c), cudaMemcpyDeviceToHost));
  out->n_particles = slot.n;
  out->n_pushes = (int64_t)c[CTR_PUSHES];
  out->n_lost = (int64_t)(c[CTR_LOST] + c[CTR_LOST_PREV]);   // ind_tetr == -1 after the call: lost in it or before it
  out->n_finished = (int64_t)c[CTR_FINISHED];
  for (int i = 0; i < 4; i++) out->n_fallback[i] = (int64_t)c[CTR_FB0 + i];
  out->n_domain_errors = (int64_t)c[CTR_DOMAIN];
  out->n_adaptive = (int64_t)c[CTR_ADAPT];
  out->n_lost_inner = (int64_t)c[CTR_LOST_INNER];
  out->n_failed = (int64_t)c[CTR_FAILED];
  float ms = 0.f;
  if (slot.have_push_time) {
    GB_CUDA(cudaEventElapsedTime(&ms, slot.ev1, slot.ev2));
    out->kernel_ms = ms;
  }
  if (slot.have_find_time) {
    GB_CUDA(cudaEventElapsedTime(&ms, slot.ev0, slot.ev1));
    out->find_ms = ms;
  }
  return GORILLA_OK;
}
