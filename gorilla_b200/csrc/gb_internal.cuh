// gb_internal.cuh -- pieces shared between gorilla_b200.cu and the per-order kernel translation units
// (gb_orbit_k{1..4}.cu exist only to compile the four polynomial orders in parallel).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <atomic>
#include "../../include/gorilla_b200.h"
#include "gb_find.cuh"
#include "gb_rk.cuh"

namespace gbint {
void set_error(const char *msg);
void count_launch(int n);
}

#define GB_CUDA(call)                                                                              \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      char buf__[512];                                                                             \
      snprintf(buf__, sizeof(buf__), "%s:%d: %s failed: %s", __FILE__, __LINE__, #call,            \
               cudaGetErrorString(e__));                                                           \
      gbint::set_error(buf__);                                                                     \
      return GORILLA_ERR_CUDA;                                                                     \
    }                                                                                              \
  } while (0)

using namespace gb;

enum { CTR_PUSHES = 0, CTR_LOST, CTR_FINISHED, CTR_FB0, CTR_FB1, CTR_FB2, CTR_FB3, CTR_DOMAIN, CTR_QUEUE, CTR_N };

struct Batch {
  int64_t n;
  double *x, *vpar, *vperp;
  double t_step;
  int32_t *init, *ind_tetr, *iface;
  double *t_remain_out;
  int64_t *n_pushes;
  int32_t trace_cap;
  int32_t *trace_tetr, *trace_face;
  unsigned long long *ctr;
  int32_t boole_periodic_relocation;
  int32_t sign_t_step;
  int32_t force_full; // debugging/parity: route every push through the complete ladder
};

// ----------------------------------------------------------------------------------------------------
// minimum resident CTAs per SM the compiler must allow for (caps registers per thread); tunable per order
#ifndef GB_MINB_K1
#define GB_MINB_K1 4
#endif
#ifndef GB_MINB_K2
#define GB_MINB_K2 4
#endif
#ifndef GB_MINB_K3
#define GB_MINB_K3 4
#endif
#ifndef GB_MINB_K4
#define GB_MINB_K4 4
#endif
#ifndef GB_MINB_RK
#define GB_MINB_RK 3
#endif
constexpr int gb_min_blocks(int K) { return K == 0 ? GB_MINB_RK : K == 1 ? GB_MINB_K1 : K == 2 ? GB_MINB_K2 : K == 3 ? GB_MINB_K3 : GB_MINB_K4; }

template <int K, int PHI>
__global__ void __launch_bounds__(128, gb_min_blocks(K)) orbit_kernel(const __grid_constant__ MeshDev m, const Batch bt)
{
  const unsigned lane = threadIdx.x & 31u;
  bool active = false, exhausted = false;
  int64_t idx = -1;
  double x[3] = {0, 0, 0}, vpar = 0, perpinv = 0, t_remain = 0, z_save[3] = {0, 0, 0};
  int32_t ind_tetr = -1, iface = -1, ind_save = -1;
  int64_t npush = 0;
  unsigned long long c_push = 0, c_lost = 0, c_fin = 0, c_fb0 = 0, c_fb1 = 0, c_fb2 = 0, c_fb3 = 0;

  for (;;) {
    if (!active && !exhausted) {
      // warp-aggregated pull from the particle queue
      const unsigned need = __activemask();
      const int leader = __ffs(need) - 1;
      unsigned long long base = 0;
      if ((int)lane == leader) base = atomicAdd(bt.ctr + CTR_QUEUE, (unsigned long long)__popc(need));
      base = __shfl_sync(need, base, leader);
      idx = (int64_t)(base + (unsigned long long)__popc(need & ((1u << lane) - 1u)));
      if (idx >= bt.n) {
        exhausted = true;
      } else {
        ind_tetr = bt.ind_tetr[idx];
        iface = bt.iface[idx];
        const bool inited = bt.init ? (bt.init[idx] != 0) : true;
        if (!inited || ind_tetr < 1) {
          // not localised (find_tetra failed) or already lost: orbit_timestep_gorilla returns at :59-61,
          // resp. leaves the loop at :103-109 without touching the particle
          if (bt.t_remain_out) bt.t_remain_out[idx] = bt.t_step;
          if (bt.n_pushes) bt.n_pushes[idx] = 0;
          if (inited && ind_tetr < 1) c_lost++;
        } else if (bt.t_step == 0.0) {
          if (bt.t_remain_out) bt.t_remain_out[idx] = 0.0;
          if (bt.n_pushes) bt.n_pushes[idx] = 0;
        } else {
          x[0] = bt.x[3 * idx];
          x[1] = bt.x[3 * idx + 1];
          x[2] = bt.x[3 * idx + 2];
          vpar = bt.vpar[idx];
          const double vperp = bt.vperp[idx];
          // :71-78  z_save = x - x1 ; perpinv = -0.5*vperp**2/bmod_func(z_save, ind_tetr)
          const double *pg = m.geom + ((int64_t)ind_tetr - 1) * GEOM_ND;
          z_save[0] = x[0] - ldg(pg);
          z_save[1] = x[1] - ldg(pg + 1);
          z_save[2] = x[2] - ldg(pg + 2);
          perpinv = -0.5 * (vperp * vperp) / bmod_at<PHI>(m, ind_tetr, z_save);
          t_remain = bt.t_step;
          ind_save = ind_tetr;
          npush = 0;
          active = true;
        }
      }
    }
    if (__all_sync(0xffffffffu, exhausted && !active)) break;
    if (active) {
      ind_save = ind_tetr;
      PushOut o;
      bool done = false;
      if constexpr (K == 0) {  // ipusher = 1: RK4 pusher
        if (!bt.force_full) {
          RkPusher<PHI> R;
          R.init(&m, perpinv, ind_tetr, x, iface, vpar, t_remain);
          done = R.template push<true>(o);
        }
        if (!done) o = push_rk_full_call<PHI>(&m, perpinv, ind_tetr, iface, x[0], x[1], x[2], vpar, t_remain);
      } else {
        if (!bt.force_full) {
          PolyPusher<K, PHI> P;
          P.mp = &m;
          P.perpinv = perpinv;
          done = P.push_fast(ind_tetr, iface, x, vpar, t_remain, o);
        }
        if (!done) o = push_full_call<K, PHI>(&m, perpinv, ind_tetr, iface, x[0], x[1], x[2], vpar, t_remain);
      }
      x[0] = o.x[0];
      x[1] = o.x[1];
      x[2] = o.x[2];
      vpar = o.vpar;
      if (o.z_save_set) {
        z_save[0] = o.z_save[0];
        z_save[1] = o.z_save[1];
        z_save[2] = o.z_save[2];
      }
      ind_tetr = o.ind_tetr;
      iface = o.iface;
      if (bt.trace_cap > 0 && npush < bt.trace_cap) {
        bt.trace_tetr[idx * bt.trace_cap + npush] = ind_tetr;
        bt.trace_face[idx * bt.trace_cap + npush] = iface;
      }
      npush++;
      c_push++;
      if (o.fallback) {
        if (o.fallback & 1) c_fb0++;
        if (o.fallback & 2) c_fb1++;
        if (o.fallback & 4) c_fb2++;
        if (o.fallback & 8) c_fb3++;
      }
      t_remain = t_remain - o.t_pass;
      if (o.finished || ind_tetr == -1) {
        // :142  vperp = vperp_func(z_save, perpinv, ind_tetr_save)
        double vperp_new = 0.0;
        if (perpinv != 0.0) vperp_new = sqrt(2.0 * fabs(perpinv) * bmod_at<PHI>(m, ind_save, z_save));
        bt.x[3 * idx] = x[0];
        bt.x[3 * idx + 1] = x[1];
        bt.x[3 * idx + 2] = x[2];
        bt.vpar[idx] = vpar;
        bt.vperp[idx] = vperp_new;
        bt.ind_tetr[idx] = ind_tetr;
        bt.iface[idx] = iface;
        if (bt.t_remain_out) bt.t_remain_out[idx] = t_remain;
        if (bt.n_pushes) bt.n_pushes[idx] = npush;
        if (o.finished) c_fin++;
        else c_lost++;
        active = false;
      }
    }
  }
  // counters: warp reduce, one atomic per warp and counter
  unsigned long long v[7] = {c_push, c_lost, c_fin, c_fb0, c_fb1, c_fb2, c_fb3};
#pragma unroll
  for (int k = 0; k < 7; k++) {
    unsigned long long s = v[k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
    if (lane == 0 && s) atomicAdd(bt.ctr + k, s);
  }
}

// ----------------------------------------------------------------------------------------------------
struct gorilla_b200_handle {
  int device = 0;
  int num_sms = 0;
  MeshDev mesh{};
  gorilla_settings settings{};
  double *d_geom = nullptr, *d_bpart = nullptr, *d_phi = nullptr, *d_cold = nullptr, *d_se = nullptr;
  unsigned long long *d_ctr = nullptr;
  // scratch for the host-pointer entry points
  int64_t cap = 0;
  double *s_x = nullptr, *s_vpar = nullptr, *s_vperp = nullptr, *s_tro = nullptr, *s_e = nullptr, *s_p = nullptr, *s_mu = nullptr;
  int32_t *s_init = nullptr, *s_ind = nullptr, *s_iface = nullptr;
  int64_t *s_np = nullptr;
  int64_t trace_cap_elems = 0;
  int32_t *s_tr_t = nullptr, *s_tr_f = nullptr;
  // sort scratch
  size_t sort_tmp_bytes = 0;
  void *sort_tmp = nullptr;
  int64_t sort_cap = 0;
  uint32_t *sort_keys_in = nullptr, *sort_keys_out = nullptr;
  int64_t *sort_vals_in = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
  bool have_find_time = false, have_push_time = false;
  int64_t last_n = 0;
  int ctas_per_sm = 0, threads_per_cta = 128;
  int force_full = 0;
  cudaStream_t last_stream = nullptr;
};

template <int K, int PHI>
int launch_orbit_t(gorilla_b200_handle *h, const Batch &bt, cudaStream_t s)
{
  int per_sm = h->ctas_per_sm;
  if (per_sm <= 0) {
    GB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, orbit_kernel<K, PHI>, h->threads_per_cta, 0));
    if (per_sm < 1) per_sm = 1;
  }
  int64_t grid = (int64_t)h->num_sms * per_sm;
  const int64_t need = (bt.n + h->threads_per_cta - 1) / h->threads_per_cta;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  orbit_kernel<K, PHI><<<(unsigned)grid, h->threads_per_cta, 0, s>>>(h->mesh, bt);
  gbint::count_launch(1);
  GB_CUDA(cudaGetLastError());
  return GORILLA_OK;
}
