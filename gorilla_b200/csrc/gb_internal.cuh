// gb_internal.cuh -- pieces shared between gorilla_b200.cu and the per-order kernel translation units
// (gb_orbit_k{1..4}.cu and gb_orbit_rk.cu exist only to compile the pusher variants in parallel).
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

enum { CTR_PUSHES = 0, CTR_LOST, CTR_FINISHED, CTR_FB0, CTR_FB1, CTR_FB2, CTR_FB3, CTR_ADAPT, CTR_DOMAIN, CTR_QUEUE, CTR_N };

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
  // EXT kernels: optional quantities summed over the pushes of the time step, [n][4] = t_hamiltonian, gyrophase,
  // vpar_int, vpar2_int (nullable); oq_mask bit q set = quantity q requested (boole_array_optional_quantities)
  double *optq;
  uint32_t oq_mask;
  // EXT = 2 kernels: orbit events (gorilla_plot_mod.f90:585-638).  ev_flags bit0 boole_poincare_phi_0, bit1
  // boole_poincare_vpar_0, bit2 boole_J_par; per-particle state in/out; events appended to a global buffer
  int32_t ev_flags, n_skip_phi_0, n_skip_vpar_0;
  double *par_adiab_inv;
  int32_t *counter_vpar_0, *counter_phi_0;
  gorilla_event *events;
  long long ev_cap;
  unsigned long long *ev_count;
};

// append the events of one push to the global buffer (order between particles is not defined; a record carries the
// particle and push index); the counter keeps counting past the capacity so that the caller sees the overflow
__device__ __forceinline__ void emit_events(const Batch &bt, long long particle, long long push, const EvState &es)
{
#pragma unroll
  for (int k = 0; k < 2; k++) {
    if (k < es.n) {
      const unsigned long long slot = atomicAdd(bt.ev_count, 1ull);
      if ((long long)slot < bt.ev_cap) {
        gorilla_event *e = bt.events + slot;
        e->particle = particle;
        e->kind = es.e[k].kind;
        e->counter = es.e[k].counter;
        e->push = push;
        e->x[0] = es.e[k].x[0]; e->x[1] = es.e[k].x[1]; e->x[2] = es.e[k].x[2];
        e->value[0] = es.e[k].v[0]; e->value[1] = es.e[k].v[1];
      }
    }
  }
}

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
#ifndef GB_MINB_X2
#define GB_MINB_X2 3   // EXT = 2 kernels of the polynomial orders (0 = as the plain kernel of the order): 168 registers measured +25-29 % over 128
#endif
constexpr int gb_min_blocks(int K, int EXT = 0)
{
  return (EXT == 2 && K != 0 && GB_MINB_X2 > 0) ? GB_MINB_X2
         : K == 0 ? GB_MINB_RK : K == 1 ? GB_MINB_K1 : K == 2 ? GB_MINB_K2 : K == 3 ? GB_MINB_K3 : GB_MINB_K4;
}

// Loop state of one particle between pushes.  It is only touched at the start and at the end of a push, while the
// push itself needs every register it can get; left to the compiler it is spilled to local memory, whose reloads
// miss the L1 (thrashed by the record gathers) and cost an L2 round trip each (ncu: ~10 such waits per push, half of
// all stall samples).  Shared memory is explicit, conflict free ([field][thread]) and ~30 cycles away.
#define GB_THREADS 128
__device__ __forceinline__ unsigned tid_now()
{
  unsigned t;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t));
  return t;
}
enum { LS_X0 = 0, LS_X1, LS_X2, LS_VPAR, LS_PERPINV, LS_TREM, LS_ZS0, LS_ZS1, LS_ZS2, LS_ND };
enum { LC_LOST = 0, LC_FIN, LC_FB0, LC_FB1, LC_FB2, LC_FB3, LC_ADAPT, LC_N };

// per-lane accumulators of the optional quantities (EXT kernels only)
template <bool EXT, int NT>
struct OqSlots {
  double v[5][NT];   // 0..3 optional quantities, 4 par_adiab_inv
  int c[2][NT];      // counter_banana_mappings, counter_phi_0_mappings
};
template <int NT>
struct OqSlots<false, NT> {
  double v[1][1];
  int c[1][1];
};

// EXT = 1: Hamiltonian time tracing (i_time_tracing_option = 2); EXT = 2: time tracing option read at run time plus the
// optional quantities of pusher_tetra_poly; the plain variant (EXT = 0) is the hot path of the default settings and
// carries none of that code (the optional-quantity code alone costs the order-2 kernel ~400 bytes of spills).
template <int K, int PHI, int EXT = 0>
__global__ void __launch_bounds__(GB_THREADS, gb_min_blocks(K, EXT)) orbit_kernel(const __grid_constant__ MeshDev m, const Batch bt)
{
  __shared__ double s_d[LS_ND][GB_THREADS], s_stash[6][GB_THREADS];
  __shared__ OqSlots<EXT == 2, GB_THREADS> s_oq;
#define LOQ(q) (((volatile double *)s_oq.v[q])[tid_now()])
#define LEC(q) (((volatile int *)s_oq.c[q])[tid_now()])
  __shared__ long long s_idx[GB_THREADS], s_npush[GB_THREADS];
  __shared__ unsigned long long s_cpush[GB_THREADS];
  __shared__ unsigned int s_cnt[LC_N][GB_THREADS];
  __shared__ int s_ind_save[GB_THREADS];
  const unsigned lane = threadIdx.x & 31u;
  // every accessor re-reads %tid.x through a volatile asm: otherwise the compiler forms the slot addresses once,
  // keeps them live across the push and spills THEM
#define LS(f) (((volatile double *)s_d[f])[tid_now()])
#define LCNT(f) (((volatile unsigned int *)s_cnt[f])[tid_now()])
#define p_idx (((volatile long long *)s_idx) + tid_now())
#define p_npush (((volatile long long *)s_npush) + tid_now())
#define p_cpush (((volatile unsigned long long *)s_cpush) + tid_now())
#define p_ind_save (((volatile int *)s_ind_save) + tid_now())
  int32_t ind_tetr = -1, iface = -1;
  *p_cpush = 0;
#pragma unroll
  for (int k = 0; k < LC_N; k++) LCNT(k) = 0;

  // Pull the next particle that actually has to be pushed into this lane's slot; false when the queue is empty.
  // Warp-aggregated: the lanes that arrive here together take consecutive queue entries with one atomic.
  auto refill = [&]() -> bool {
    for (;;) {
      const unsigned need = __activemask();
      const int leader = __ffs(need) - 1;
      unsigned long long base = 0;
      if ((int)lane == leader) base = atomicAdd(bt.ctr + CTR_QUEUE, (unsigned long long)__popc(need));
      base = __shfl_sync(need, base, leader);
      const int64_t idx = (int64_t)(base + (unsigned long long)__popc(need & ((1u << lane) - 1u)));
      if (idx >= bt.n) return false;
      ind_tetr = bt.ind_tetr[idx];
      iface = bt.iface[idx];
      const bool inited = bt.init ? (bt.init[idx] != 0) : true;
      if (!inited || ind_tetr < 1) {
        // not localised (find_tetra failed) or already lost: orbit_timestep_gorilla returns at :59-61,
        // resp. leaves the loop at :103-109 without touching the particle
        if (bt.t_remain_out) bt.t_remain_out[idx] = bt.t_step;
        if (bt.n_pushes) bt.n_pushes[idx] = 0;
        if constexpr (EXT == 2) {
          if (bt.optq) { bt.optq[4 * idx] = 0.0; bt.optq[4 * idx + 1] = 0.0; bt.optq[4 * idx + 2] = 0.0; bt.optq[4 * idx + 3] = 0.0; }
        }
        if (inited && ind_tetr < 1) LCNT(LC_LOST) = LCNT(LC_LOST) + 1;
        continue;
      }
      if (bt.t_step == 0.0) {
        if (bt.t_remain_out) bt.t_remain_out[idx] = 0.0;
        if (bt.n_pushes) bt.n_pushes[idx] = 0;
        if constexpr (EXT == 2) {
          if (bt.optq) { bt.optq[4 * idx] = 0.0; bt.optq[4 * idx + 1] = 0.0; bt.optq[4 * idx + 2] = 0.0; bt.optq[4 * idx + 3] = 0.0; }
        }
        continue;
      }
      const double x0 = bt.x[3 * idx], x1 = bt.x[3 * idx + 1], x2 = bt.x[3 * idx + 2];
      const double vperp = bt.vperp[idx];
      // :71-78  z_save = x - x1 ; perpinv = -0.5*vperp**2/bmod_func(z_save, ind_tetr)
      const double *pg = m.geom + ((int64_t)ind_tetr - 1) * GEOM_ND;
      const double zs[3] = {x0 - ldg(pg), x1 - ldg(pg + 1), x2 - ldg(pg + 2)};
      LS(LS_X0) = x0; LS(LS_X1) = x1; LS(LS_X2) = x2;
      LS(LS_VPAR) = bt.vpar[idx];
      LS(LS_ZS0) = zs[0]; LS(LS_ZS1) = zs[1]; LS(LS_ZS2) = zs[2];
      LS(LS_PERPINV) = -0.5 * (vperp * vperp) / bmod_at<PHI>(m, ind_tetr, zs);
      LS(LS_TREM) = bt.t_step;
      *p_idx = idx;
      *p_npush = 0;
      if constexpr (EXT == 2) {
        LOQ(0) = 0.0; LOQ(1) = 0.0; LOQ(2) = 0.0; LOQ(3) = 0.0;
        if (bt.ev_flags) { LOQ(4) = bt.par_adiab_inv[idx]; LEC(0) = bt.counter_vpar_0[idx]; LEC(1) = bt.counter_phi_0[idx]; }
      }
      return true;
    }
  };

  // One lane = one particle at a time.  A lane whose particle is done refills itself at the end of the same loop
  // body and leaves the loop for good when the queue is empty, so the body has no "is this lane active" region (whose
  // convergence-barrier register was live, and spilled, across every push).
  bool active = refill();
  while (active) {
    {
      *p_ind_save = ind_tetr;
      PushOut o;
      bool done = false;
      const double perpinv = LS(LS_PERPINV);
      if constexpr (K == 0) {  // ipusher = 1: RK4 pusher
        if (!bt.force_full) {
          const double x[3] = {LS(LS_X0), LS(LS_X1), LS(LS_X2)};
          RkPusher<PHI, (EXT == 2 ? 2 : 0)> R;
          R.P.r.set_stash(&s_stash[0][tid_now()], GB_THREADS);
          R.init(&m, perpinv, ind_tetr, x, iface, LS(LS_VPAR), LS(LS_TREM));
          done = R.template push<true>(o);
        }
        if (!done)
          o = push_rk_full_call<PHI, (EXT == 2 ? 2 : 0)>(&m, perpinv, ind_tetr, iface, LS(LS_X0), LS(LS_X1), LS(LS_X2), LS(LS_VPAR), LS(LS_TREM));
      } else {
        if (!bt.force_full) {
          const double x[3] = {LS(LS_X0), LS(LS_X1), LS(LS_X2)};
          PolyPusher<K, PHI, EXT> P;
          P.mp = &m;
          P.perpinv = perpinv;
          if constexpr (EXT == 2) P.oq_mask = bt.oq_mask;
          P.r.set_stash(&s_stash[0][tid_now()], GB_THREADS);
          done = P.push_fast(ind_tetr, iface, x, LS(LS_VPAR), LS(LS_TREM), o, &LS(LS_TREM));
          if constexpr (EXT == 2) {
            if (done && bt.oq_mask) {
#pragma unroll
              for (int q = 0; q < 4; q++) LOQ(q) = LOQ(q) + P.oq[q];
            }
            if constexpr (K >= 2) {
              if (done && bt.ev_flags && !o.finished) {
                EvState es;
                es.flags = bt.ev_flags; es.nskip_p = bt.n_skip_phi_0; es.nskip_v = bt.n_skip_vpar_0;
                es.J = LOQ(4); es.cnt_v = LEC(0); es.cnt_p = LEC(1);
                P.events_after_push(LS(LS_VPAR), o, es);
                LOQ(4) = es.J; LEC(0) = es.cnt_v; LEC(1) = es.cnt_p;
                if (es.n) emit_events(bt, *p_idx, *p_npush, es);
              }
            }
          }
        }
        if (!done) {
          if constexpr (EXT == 2) {
            const PushOutX ox = push_full_call_x<K, PHI>(&m, perpinv, ind_tetr, iface, LS(LS_X0), LS(LS_X1), LS(LS_X2),
                                                         LS(LS_VPAR), LS(LS_TREM), bt.oq_mask, bt.ev_flags, bt.n_skip_phi_0,
                                                         bt.n_skip_vpar_0, LOQ(4), LEC(0), LEC(1));
            o = ox.o;
#pragma unroll
            for (int q = 0; q < 4; q++) LOQ(q) = LOQ(q) + ox.oq[q];
            if (bt.ev_flags) {
              LOQ(4) = ox.es.J; LEC(0) = ox.es.cnt_v; LEC(1) = ox.es.cnt_p;
              if (ox.es.n) emit_events(bt, *p_idx, *p_npush, ox.es);
            }
          } else {
            o = push_full_call<K, PHI, EXT>(&m, perpinv, ind_tetr, iface, LS(LS_X0), LS(LS_X1), LS(LS_X2), LS(LS_VPAR), LS(LS_TREM));
          }
        }
      }
      LS(LS_X0) = o.x[0]; LS(LS_X1) = o.x[1]; LS(LS_X2) = o.x[2];
      LS(LS_VPAR) = o.vpar;
      if (o.z_save_set) { LS(LS_ZS0) = o.z_save[0]; LS(LS_ZS1) = o.z_save[1]; LS(LS_ZS2) = o.z_save[2]; }
      ind_tetr = o.ind_tetr;
      iface = o.iface;
      const long long npush = *p_npush;
      if (bt.trace_cap > 0 && npush < bt.trace_cap) {
        const long long idx = *p_idx;
        bt.trace_tetr[idx * bt.trace_cap + npush] = ind_tetr;
        bt.trace_face[idx * bt.trace_cap + npush] = iface;
      }
      *p_npush = npush + 1;
      if (o.fallback) {
        if (o.fallback & 1) LCNT(LC_FB0) = LCNT(LC_FB0) + 1;
        if (o.fallback & 2) LCNT(LC_FB1) = LCNT(LC_FB1) + 1;
        if (o.fallback & 4) LCNT(LC_FB2) = LCNT(LC_FB2) + 1;
        if (o.fallback & 8) LCNT(LC_FB3) = LCNT(LC_FB3) + 1;
        if constexpr (EXT == 3) {
          if (o.fallback & 16) LCNT(LC_ADAPT) = LCNT(LC_ADAPT) + 1;
        }
      }
      const double t_remain = LS(LS_TREM) - o.t_pass;
      LS(LS_TREM) = t_remain;
      if (o.finished || ind_tetr == -1) {
        // :142  vperp = vperp_func(z_save, perpinv, ind_tetr_save)
        const long long idx = *p_idx;
        const double pinv = LS(LS_PERPINV);
        const double zs[3] = {LS(LS_ZS0), LS(LS_ZS1), LS(LS_ZS2)};
        double vperp_new = 0.0;
        if (pinv != 0.0) vperp_new = sqrt(2.0 * fabs(pinv) * bmod_at<PHI>(m, *p_ind_save, zs));
        bt.x[3 * idx] = o.x[0];
        bt.x[3 * idx + 1] = o.x[1];
        bt.x[3 * idx + 2] = o.x[2];
        bt.vpar[idx] = o.vpar;
        bt.vperp[idx] = vperp_new;
        bt.ind_tetr[idx] = ind_tetr;
        bt.iface[idx] = iface;
        if (bt.t_remain_out) bt.t_remain_out[idx] = t_remain;
        if (bt.n_pushes) bt.n_pushes[idx] = npush + 1;
        if constexpr (EXT == 2) {
          if (bt.optq) {
#pragma unroll
            for (int q = 0; q < 4; q++) bt.optq[4 * idx + q] = LOQ(q);
          }
          if (bt.ev_flags) { bt.par_adiab_inv[idx] = LOQ(4); bt.counter_vpar_0[idx] = LEC(0); bt.counter_phi_0[idx] = LEC(1); }
        }
        *p_cpush = *p_cpush + (unsigned long long)(npush + 1);
        if (o.finished) LCNT(LC_FIN) = LCNT(LC_FIN) + 1;
        else LCNT(LC_LOST) = LCNT(LC_LOST) + 1;
        active = refill();
      }
    }
  }
  __syncwarp();
  // counters: warp reduce, one atomic per warp and counter
  unsigned long long v[8] = {*p_cpush, LCNT(LC_LOST), LCNT(LC_FIN), LCNT(LC_FB0), LCNT(LC_FB1), LCNT(LC_FB2), LCNT(LC_FB3),
                             LCNT(LC_ADAPT)};
#pragma unroll
  for (int k = 0; k < 8; k++) {
    unsigned long long s = v[k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
    if (lane == 0 && s) atomicAdd(bt.ctr + k, s);
  }
#undef LS
#undef LCNT
#undef LOQ
#undef LEC
#undef p_idx
#undef p_npush
#undef p_cpush
#undef p_ind_save
}

// ----------------------------------------------------------------------------------------------------
// orbit_kernel_g<K,PHI> -- orders 3 and 4 with the iterative exit-time solve run in LOCK STEP by the warps that share an
// SM sub-partition.
//
// ncu: the order-3/4 kernels are instruction-fetch bound (no_instruction = 57 % of all stall samples).  One iteration of
// the root solver is ~1600 instructions (26 KB) against a 6 KB L0 instruction cache per sub-partition, and with CTAs of
// four warps the four warps of a sub-partition belong to four CTAs and are at unrelated places in that loop, so every
// warp streams the whole loop through the L0 on its own.  Here a CTA has 16 warps; warps w, w+4, w+8, w+12 sit on
// sub-partition w and form a group with its own named barrier.  The group walks the push loop together and executes
// every solver iteration behind the barrier, so its four warps fetch the same lines at the same time and one L0 fill
// serves all of them.  Per-particle arithmetic is that of orbit_kernel<K,PHI> (same functions): results are identical.
#define GBG_THREADS 512
#define GBG_GROUP 128   // threads per group = 4 warps

// group-wide OR of a per-thread predicate + barrier; must be reached by whole warps of the group in converged state
__device__ __forceinline__ bool group_any(bool pred, int bar_id)
{
  int r;
  asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.s32 p, %1, 0;\n\tbarrier.red.or.pred q, %2, %3, p;\n\tselp.s32 %0, 1, 0, q;\n\t}"
               : "=r"(r) : "r"((int)pred), "r"(bar_id), "r"(GBG_GROUP) : "memory");
  return r != 0;
}

// monic polynomial solve of every lane that has one (busy), iteration by iteration behind the group barrier
static __device__ __noinline__ double solve_group(bool busy, int deg, double q0, double q1, double q2, double q3, double lambda,
                                                  double tau_ready, int bar_id)
{
  SgSolver S;
  cd poly[5];
  poly[0] = mk(q0, 0.0);
  poly[1] = mk(deg == 1 ? 1.0 : q1, 0.0);
  poly[2] = mk(deg == 2 ? 1.0 : q2, 0.0);
  poly[3] = mk(deg == 3 ? 1.0 : q3, 0.0);
  poly[4] = mk(1.0, 0.0);
  S.start(busy ? deg : 2, poly);
  double tau = tau_ready;
  for (;;) {
    if (!group_any(busy, bar_id)) break;
    if (busy && S.step()) {
      tau = min_positive_real_root(S.deg, S.roots, lambda);
      busy = false;
    }
  }
  return tau;
}

template <int K, int PHI, int EXT = 0>
__global__ void __launch_bounds__(GBG_THREADS, 1) orbit_kernel_g(const __grid_constant__ MeshDev m, const Batch bt)
{
  extern __shared__ __align__(16) unsigned char g_smem[];
  double (*s_d)[GBG_THREADS] = reinterpret_cast<double (*)[GBG_THREADS]>(g_smem);                       // [LS_ND]
  double (*s_stash)[GBG_THREADS] = s_d + LS_ND;                                                         // [6]
  long long *s_idx = reinterpret_cast<long long *>(s_stash + 6), *s_npush = s_idx + GBG_THREADS;
  unsigned long long *s_cpush = reinterpret_cast<unsigned long long *>(s_npush + GBG_THREADS);
  unsigned int (*s_cnt)[GBG_THREADS] = reinterpret_cast<unsigned int (*)[GBG_THREADS]>(s_cpush + GBG_THREADS);  // [LC_N]
  int *s_ind_save = reinterpret_cast<int *>(s_cnt + LC_N);
  double (*s_oq)[GBG_THREADS] = reinterpret_cast<double (*)[GBG_THREADS]>(s_ind_save + GBG_THREADS);   // [5], EXT = 2 only
  int (*s_ec)[GBG_THREADS] = reinterpret_cast<int (*)[GBG_THREADS]>(s_oq + 5);                          // [2], EXT = 2 only
#define LOQ(q) (((volatile double *)s_oq[q])[tid_now()])
#define LEC(q) (((volatile int *)s_ec[q])[tid_now()])
  const unsigned lane = threadIdx.x & 31u;
  const int bar_id = 1 + (int)((threadIdx.x >> 5) & 3u);   // warps w, w+4, w+8, w+12 share sub-partition w
#define LS(f) (((volatile double *)s_d[f])[tid_now()])
#define LCNT(f) (((volatile unsigned int *)s_cnt[f])[tid_now()])
#define p_idx (((volatile long long *)s_idx) + tid_now())
#define p_npush (((volatile long long *)s_npush) + tid_now())
#define p_cpush (((volatile unsigned long long *)s_cpush) + tid_now())
#define p_ind_save (((volatile int *)s_ind_save) + tid_now())
  int32_t ind_tetr = -1, iface = -1;
  *p_cpush = 0;
#pragma unroll
  for (int k = 0; k < LC_N; k++) LCNT(k) = 0;

  auto refill = [&]() -> bool {
    for (;;) {
      const unsigned need = __activemask();
      const int leader = __ffs(need) - 1;
      unsigned long long base = 0;
      if ((int)lane == leader) base = atomicAdd(bt.ctr + CTR_QUEUE, (unsigned long long)__popc(need));
      base = __shfl_sync(need, base, leader);
      const int64_t idx = (int64_t)(base + (unsigned long long)__popc(need & ((1u << lane) - 1u)));
      if (idx >= bt.n) return false;
      ind_tetr = bt.ind_tetr[idx];
      iface = bt.iface[idx];
      const bool inited = bt.init ? (bt.init[idx] != 0) : true;
      if (!inited || ind_tetr < 1) {
        if (bt.t_remain_out) bt.t_remain_out[idx] = bt.t_step;
        if (bt.n_pushes) bt.n_pushes[idx] = 0;
        if constexpr (EXT == 2) {
          if (bt.optq) { bt.optq[4 * idx] = 0.0; bt.optq[4 * idx + 1] = 0.0; bt.optq[4 * idx + 2] = 0.0; bt.optq[4 * idx + 3] = 0.0; }
        }
        if (inited && ind_tetr < 1) LCNT(LC_LOST) = LCNT(LC_LOST) + 1;
        continue;
      }
      if (bt.t_step == 0.0) {
        if (bt.t_remain_out) bt.t_remain_out[idx] = 0.0;
        if (bt.n_pushes) bt.n_pushes[idx] = 0;
        if constexpr (EXT == 2) {
          if (bt.optq) { bt.optq[4 * idx] = 0.0; bt.optq[4 * idx + 1] = 0.0; bt.optq[4 * idx + 2] = 0.0; bt.optq[4 * idx + 3] = 0.0; }
        }
        continue;
      }
      const double x0 = bt.x[3 * idx], x1 = bt.x[3 * idx + 1], x2 = bt.x[3 * idx + 2];
      const double vperp = bt.vperp[idx];
      const double *pg = m.geom + ((int64_t)ind_tetr - 1) * GEOM_ND;
      const double zs[3] = {x0 - ldg(pg), x1 - ldg(pg + 1), x2 - ldg(pg + 2)};
      LS(LS_X0) = x0; LS(LS_X1) = x1; LS(LS_X2) = x2;
      LS(LS_VPAR) = bt.vpar[idx];
      LS(LS_ZS0) = zs[0]; LS(LS_ZS1) = zs[1]; LS(LS_ZS2) = zs[2];
      LS(LS_PERPINV) = -0.5 * (vperp * vperp) / bmod_at<PHI>(m, ind_tetr, zs);
      LS(LS_TREM) = bt.t_step;
      *p_idx = idx;
      *p_npush = 0;
      if constexpr (EXT == 2) {
        LOQ(0) = 0.0; LOQ(1) = 0.0; LOQ(2) = 0.0; LOQ(3) = 0.0;
        if (bt.ev_flags) { LOQ(4) = bt.par_adiab_inv[idx]; LEC(0) = bt.counter_vpar_0[idx]; LEC(1) = bt.counter_phi_0[idx]; }
      }
      return true;
    }
  };

  bool active = refill();
  for (;;) {
    if (!group_any(active, bar_id)) break;   // the group leaves together
    PushOut o;
    bool begun = false, done = false;
    SolveTask t;
    t.kind = 0; t.deg = 2; t.tau = 0.0; t.lambda = 1.0;
    t.q[0] = t.q[1] = t.q[2] = t.q[3] = 0.0;
    int iface_new = 0;
    double tau_max = 0.0;
    PolyPusher<K, PHI, EXT> P;
    P.mp = &m;
    if constexpr (EXT == 2) P.oq_mask = bt.oq_mask;
    P.r.set_stash(&s_stash[0][tid_now()], GBG_THREADS);
    if (active && !bt.force_full) {
      const double x[3] = {LS(LS_X0), LS(LS_X1), LS(LS_X2)};
      P.perpinv = LS(LS_PERPINV);
      begun = P.fast_begin(ind_tetr, iface, x, LS(LS_VPAR), LS(LS_TREM), t, iface_new, tau_max) && t.kind != 0;
    }
    const double tau = solve_group(begun && t.kind == 2, t.deg, t.q[0], t.q[1], t.q[2], t.q[3], t.lambda, t.tau, bar_id);
    if (begun) {
      P.t_remain = LS(LS_TREM);
      done = P.fast_end(tau, iface_new, tau_max, true, o);
    }
    if (active) {
      *p_ind_save = ind_tetr;
      if constexpr (EXT == 2) {
        if (done) {
          if (bt.oq_mask) {
#pragma unroll
            for (int q = 0; q < 4; q++) LOQ(q) = LOQ(q) + P.oq[q];
          }
          if (bt.ev_flags && !o.finished) {
            EvState es;
            es.flags = bt.ev_flags; es.nskip_p = bt.n_skip_phi_0; es.nskip_v = bt.n_skip_vpar_0;
            es.J = LOQ(4); es.cnt_v = LEC(0); es.cnt_p = LEC(1);
            P.events_after_push(LS(LS_VPAR), o, es);
            LOQ(4) = es.J; LEC(0) = es.cnt_v; LEC(1) = es.cnt_p;
            if (es.n) emit_events(bt, *p_idx, *p_npush, es);
          }
        } else {
          const PushOutX ox = push_full_call_x<K, PHI>(&m, LS(LS_PERPINV), ind_tetr, iface, LS(LS_X0), LS(LS_X1), LS(LS_X2),
                                                       LS(LS_VPAR), LS(LS_TREM), bt.oq_mask, bt.ev_flags, bt.n_skip_phi_0,
                                                       bt.n_skip_vpar_0, LOQ(4), LEC(0), LEC(1));
          o = ox.o;
#pragma unroll
          for (int q = 0; q < 4; q++) LOQ(q) = LOQ(q) + ox.oq[q];
          if (bt.ev_flags) {
            LOQ(4) = ox.es.J; LEC(0) = ox.es.cnt_v; LEC(1) = ox.es.cnt_p;
            if (ox.es.n) emit_events(bt, *p_idx, *p_npush, ox.es);
          }
        }
      } else {
        if (!done)
          o = push_full_call<K, PHI, EXT>(&m, LS(LS_PERPINV), ind_tetr, iface, LS(LS_X0), LS(LS_X1), LS(LS_X2), LS(LS_VPAR), LS(LS_TREM));
      }
      LS(LS_X0) = o.x[0]; LS(LS_X1) = o.x[1]; LS(LS_X2) = o.x[2];
      LS(LS_VPAR) = o.vpar;
      if (o.z_save_set) { LS(LS_ZS0) = o.z_save[0]; LS(LS_ZS1) = o.z_save[1]; LS(LS_ZS2) = o.z_save[2]; }
      const int ind_prev = ind_tetr;
      ind_tetr = o.ind_tetr;
      iface = o.iface;
      const long long npush = *p_npush;
      if (bt.trace_cap > 0 && npush < bt.trace_cap) {
        const long long idx = *p_idx;
        bt.trace_tetr[idx * bt.trace_cap + npush] = ind_tetr;
        bt.trace_face[idx * bt.trace_cap + npush] = iface;
      }
      *p_npush = npush + 1;
      if (o.fallback) {
        if (o.fallback & 1) LCNT(LC_FB0) = LCNT(LC_FB0) + 1;
        if (o.fallback & 2) LCNT(LC_FB1) = LCNT(LC_FB1) + 1;
        if (o.fallback & 4) LCNT(LC_FB2) = LCNT(LC_FB2) + 1;
        if (o.fallback & 8) LCNT(LC_FB3) = LCNT(LC_FB3) + 1;
        if constexpr (EXT == 3) {
          if (o.fallback & 16) LCNT(LC_ADAPT) = LCNT(LC_ADAPT) + 1;
        }
      }
      const double t_remain = LS(LS_TREM) - o.t_pass;
      LS(LS_TREM) = t_remain;
      if (o.finished || ind_tetr == -1) {
        const long long idx = *p_idx;
        const double pinv = LS(LS_PERPINV);
        const double zs[3] = {LS(LS_ZS0), LS(LS_ZS1), LS(LS_ZS2)};
        double vperp_new = 0.0;
        if (pinv != 0.0) vperp_new = sqrt(2.0 * fabs(pinv) * bmod_at<PHI>(m, ind_prev, zs));
        bt.x[3 * idx] = o.x[0];
        bt.x[3 * idx + 1] = o.x[1];
        bt.x[3 * idx + 2] = o.x[2];
        bt.vpar[idx] = o.vpar;
        bt.vperp[idx] = vperp_new;
        bt.ind_tetr[idx] = ind_tetr;
        bt.iface[idx] = iface;
        if (bt.t_remain_out) bt.t_remain_out[idx] = t_remain;
        if (bt.n_pushes) bt.n_pushes[idx] = npush + 1;
        if constexpr (EXT == 2) {
          if (bt.optq) {
#pragma unroll
            for (int q = 0; q < 4; q++) bt.optq[4 * idx + q] = LOQ(q);
          }
          if (bt.ev_flags) { bt.par_adiab_inv[idx] = LOQ(4); bt.counter_vpar_0[idx] = LEC(0); bt.counter_phi_0[idx] = LEC(1); }
        }
        *p_cpush = *p_cpush + (unsigned long long)(npush + 1);
        if (o.finished) LCNT(LC_FIN) = LCNT(LC_FIN) + 1;
        else LCNT(LC_LOST) = LCNT(LC_LOST) + 1;
        active = refill();
      }
    }
  }
  __syncwarp();
  unsigned long long v[8] = {*p_cpush, LCNT(LC_LOST), LCNT(LC_FIN), LCNT(LC_FB0), LCNT(LC_FB1), LCNT(LC_FB2), LCNT(LC_FB3),
                             LCNT(LC_ADAPT)};
#pragma unroll
  for (int k = 0; k < 8; k++) {
    unsigned long long sacc = v[k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sacc += __shfl_down_sync(0xffffffffu, sacc, off);
    if (lane == 0 && sacc) atomicAdd(bt.ctr + k, sacc);
  }
#undef LS
#undef LCNT
#undef LOQ
#undef LEC
#undef p_idx
#undef p_npush
#undef p_cpush
#undef p_ind_save
}
constexpr size_t GBG_SMEM = (size_t)GBG_THREADS * ((LS_ND + 6) * 8 + 3 * 8 + LC_N * 4 + 4);
constexpr size_t GBG_SMEM_EXT = GBG_SMEM + (size_t)GBG_THREADS * (5 * 8 + 2 * 4);

// ----------------------------------------------------------------------------------------------------
struct gorilla_b200_handle {
  int device = 0;
  int num_sms = 0;
  MeshDev mesh{};
  gorilla_settings settings{};
  double *d_geom = nullptr, *d_bpart = nullptr, *d_phi = nullptr, *d_cold = nullptr, *d_se = nullptr, *d_ham = nullptr, *d_skew = nullptr;
  double *s_oq = nullptr;   // [cap][4] scratch for the optional quantities (host-pointer entry point)
  uint32_t oq_mask = 0;
  int32_t *d_bin_start = nullptr, *d_bin_items = nullptr;
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
  int use_group = 1;  // orders 3/4: lock-step solver kernel (orbit_kernel_g)
  cudaStream_t last_stream = nullptr;
};

template <int K, int PHI, int EXT = 0>
int launch_orbit_t(gorilla_b200_handle *h, const Batch &bt, cudaStream_t s)
{
  if constexpr (K >= 3) {
    if (h->use_group) {
      constexpr size_t smem_g = EXT == 2 ? GBG_SMEM_EXT : GBG_SMEM;
      GB_CUDA(cudaFuncSetAttribute(orbit_kernel_g<K, PHI, EXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
      int64_t grid_g = h->num_sms;
      const int64_t need_g = (bt.n + GBG_THREADS - 1) / GBG_THREADS;
      if (grid_g > need_g) grid_g = need_g;
      if (grid_g < 1) grid_g = 1;
      orbit_kernel_g<K, PHI, EXT><<<(unsigned)grid_g, GBG_THREADS, smem_g, s>>>(h->mesh, bt);
      gbint::count_launch(1);
      GB_CUDA(cudaGetLastError());
      return GORILLA_OK;
    }
  }
  int per_sm = h->ctas_per_sm;
  if (per_sm <= 0) {
    GB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, orbit_kernel<K, PHI, EXT>, h->threads_per_cta, 0));
    if (per_sm < 1) per_sm = 1;
  }
  int64_t grid = (int64_t)h->num_sms * per_sm;
  const int64_t need = (bt.n + h->threads_per_cta - 1) / h->threads_per_cta;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  orbit_kernel<K, PHI, EXT><<<(unsigned)grid, h->threads_per_cta, 0, s>>>(h->mesh, bt);
  gbint::count_launch(1);
  GB_CUDA(cudaGetLastError());
  return GORILLA_OK;
}
