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

struct gorilla_b200_handle;
namespace gbint {
void set_error(const char *msg);
void count_launch(int n);
int fail(int code, const char *msg);
// gb_diag.cu: particle re-sorting through the handle's scratch
int sort_permutation(gorilla_b200_handle *h, int64_t n, const int32_t *ind_tetr, int64_t *perm, cudaStream_t s);
int permute_state_inplace(gorilla_b200_handle *h, int64_t n, const int64_t *perm, bool inverse, double *x, double *vpar,
                          double *vperp, int32_t *init, int32_t *ind, int32_t *iface, cudaStream_t s);
template <typename T>
int permute_one_inplace(gorilla_b200_handle *h, int64_t n, const int64_t *perm, bool inverse, T *a, cudaStream_t s);
int ensure_host_scratch(gorilla_b200_handle *h, int64_t n);   // the s_* arrays of the host-pointer entry points
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

enum { CTR_PUSHES = 0, CTR_LOST, CTR_FINISHED, CTR_FB0, CTR_FB1, CTR_FB2, CTR_FB3, CTR_ADAPT, CTR_DOMAIN, CTR_QUEUE, CTR_LOST_INNER,
       CTR_FAILED, CTR_LOST_PREV, CTR_N };

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
  int32_t rebin;      // orders 3/4 (orbit_kernel_g): re-bin the root solves of a group by solver mode between iterations
  // EXT kernels: optional quantities summed over the pushes of the time step, [n][4] = t_hamiltonian, gyrophase,
  // vpar_int, vpar2_int (nullable); oq_mask bit q set = quantity q requested (boole_array_optional_quantities)
  double *optq;
  uint32_t oq_mask;
  // EXT = 2 kernels: orbit events (gorilla_plot_mod.f90:585-638).  ev_flags bit0 boole_poincare_phi_0, bit1
  // boole_poincare_vpar_0, bit2 boole_J_par; per-particle state in/out; events appended to a global buffer
  int32_t ev_flags, n_skip_phi_0, n_skip_vpar_0;   // ev_flags bit3: boole_full_orbit (:553-579)
  int32_t n_skip_full_orbit;
  double *par_adiab_inv;
  int32_t *counter_vpar_0, *counter_phi_0;
  gorilla_event *events;
  long long ev_cap;
  unsigned long long *ev_count;
  // EXT = 5 kernels (adaptive sub-stepping + list consumers): per-thread step lists, [thread][lst_cap][5]
  double *lst;
  int32_t lst_cap;
};

// append the events of one push to the global buffer (order between particles is not defined; a record carries the
// particle and push index); the counter keeps counting past the capacity so that the caller sees the overflow
__device__ __forceinline__ void emit_events(const Batch &bt, long long particle, long long push, const EvState &es, double t)
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
        e->t = t;
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
  return ((EXT == 2 || EXT == 5) && K != 0 && GB_MINB_X2 > 0) ? GB_MINB_X2
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
enum { LC_LOST = 0, LC_FIN, LC_FB0, LC_FB1, LC_FB2, LC_FB3, LC_ADAPT, LC_LOST_INNER, LC_FAILED, LC_LOST_PREV, LC_N };

// The per-lane slots of a CTA of NT threads ([field][thread], conflict free), shared by both kernel shapes: the 4-warp
// kernel places them in static shared memory, the 16-warp lock-step kernel in dynamic shared memory.  Every accessor
// re-reads %tid.x through a volatile asm: otherwise the compiler forms the slot addresses once, keeps them live across the
// push and spills THEM.
// STASH = false (cooperative-gather kernels that stage the whole record, PHI >= 1): no stash in shared memory, the kernel
// keeps those six doubles in local memory.
template <int NT, bool STASH = true>
struct LaneSlots {
  static constexpr int NSTASH = STASH ? 6 : 0;
  double (*d)[NT];        // [LS_ND] loop state
  double (*stash)[NT];    // [6] first vertex + topology words of the current record (Rec::load)
  long long *idx, *npush;
  unsigned long long *cpush;
  unsigned int (*cnt)[NT];  // [LC_N]
  int *ind_save;
  double (*oq)[NT];       // [5] EXT = 2: optional quantities 0..3, par_adiab_inv
  int (*ec)[NT];          // [2] EXT = 2: counter_banana_mappings, counter_phi_0_mappings
  static constexpr size_t BYTES = (size_t)NT * ((LS_ND + NSTASH) * 8 + 3 * 8 + LC_N * 4 + 4 + 4 /*pad to 8*/);
  static constexpr size_t BYTES_EXT2 = BYTES + (size_t)NT * (5 * 8 + 2 * 4);
  __device__ __forceinline__ void carve(unsigned char *base, bool ext2)
  {
    d = reinterpret_cast<double (*)[NT]>(base);
    stash = d + LS_ND;
    idx = reinterpret_cast<long long *>(stash + NSTASH);
    npush = idx + NT;
    cpush = reinterpret_cast<unsigned long long *>(npush + NT);
    cnt = reinterpret_cast<unsigned int (*)[NT]>(cpush + NT);
    ind_save = reinterpret_cast<int *>(cnt + LC_N);
    oq = reinterpret_cast<double (*)[NT]>(ind_save + 2 * NT);
    ec = reinterpret_cast<int (*)[NT]>(oq + 5);
    (void)ext2;
  }
  __device__ __forceinline__ volatile double &D(int f) const { return ((volatile double *)d[f])[tid_now()]; }
  __device__ __forceinline__ volatile unsigned int &C(int f) const { return ((volatile unsigned int *)cnt[f])[tid_now()]; }
  __device__ __forceinline__ volatile double &OQ(int q) const { return ((volatile double *)oq[q])[tid_now()]; }
  __device__ __forceinline__ volatile int &EC(int q) const { return ((volatile int *)ec[q])[tid_now()]; }
  __device__ __forceinline__ volatile long long &Idx() const { return ((volatile long long *)idx)[tid_now()]; }
  __device__ __forceinline__ volatile long long &Npush() const { return ((volatile long long *)npush)[tid_now()]; }
  __device__ __forceinline__ volatile unsigned long long &Cpush() const { return ((volatile unsigned long long *)cpush)[tid_now()]; }
  __device__ __forceinline__ volatile int &IndSave() const { return ((volatile int *)ind_save)[tid_now()]; }
  __device__ __forceinline__ volatile double *Stash() const { return &stash[0][tid_now()]; }
  __device__ __forceinline__ void zero_counters() const
  {
    Cpush() = 0;
#pragma unroll
    for (int k = 0; k < LC_N; k++) C(k) = 0;
  }
};

// Pull the next particle that actually has to be pushed into this lane's slot; false when the queue is empty.
// Warp-aggregated: the lanes that arrive here together take consecutive queue entries with one atomic.
template <int PHI, int EXT, int NT, bool ST>
__device__ __forceinline__ bool lane_refill(const MeshDev &m, const Batch &bt, const LaneSlots<NT, ST> &S, unsigned lane,
                                            int32_t &ind_tetr, int32_t &iface)
{
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
      if constexpr (EXT == 2 || EXT == 5) {
        if (bt.optq) { bt.optq[4 * idx] = 0.0; bt.optq[4 * idx + 1] = 0.0; bt.optq[4 * idx + 2] = 0.0; bt.optq[4 * idx + 3] = 0.0; }
      }
      if (inited && ind_tetr < 1) S.C(LC_LOST_PREV) = S.C(LC_LOST_PREV) + 1;   // lost in an earlier call
      continue;
    }
    if (bt.t_step == 0.0) {
      if (bt.t_remain_out) bt.t_remain_out[idx] = 0.0;
      if (bt.n_pushes) bt.n_pushes[idx] = 0;
      if constexpr (EXT == 2 || EXT == 5) {
        if (bt.optq) { bt.optq[4 * idx] = 0.0; bt.optq[4 * idx + 1] = 0.0; bt.optq[4 * idx + 2] = 0.0; bt.optq[4 * idx + 3] = 0.0; }
      }
      continue;
    }
    const double x0 = bt.x[3 * idx], x1 = bt.x[3 * idx + 1], x2 = bt.x[3 * idx + 2];
    const double vperp = bt.vperp[idx];
    // :71-78  z_save = x - x1 ; perpinv = -0.5*vperp**2/bmod_func(z_save, ind_tetr)
    const double *pg = m.geom + ((int64_t)ind_tetr - 1) * GEOM_ND;
    const double zs[3] = {x0 - ldg(pg), x1 - ldg(pg + 1), x2 - ldg(pg + 2)};
    S.D(LS_X0) = x0; S.D(LS_X1) = x1; S.D(LS_X2) = x2;
    S.D(LS_VPAR) = bt.vpar[idx];
    S.D(LS_ZS0) = zs[0]; S.D(LS_ZS1) = zs[1]; S.D(LS_ZS2) = zs[2];
    S.D(LS_PERPINV) = -0.5 * (vperp * vperp) / bmod_at<PHI>(m, ind_tetr, zs);
    S.D(LS_TREM) = bt.t_step;
    S.Idx() = idx;
    S.Npush() = 0;
    if constexpr (EXT == 2 || EXT == 5) {
      S.OQ(0) = 0.0; S.OQ(1) = 0.0; S.OQ(2) = 0.0; S.OQ(3) = 0.0;
      if (bt.ev_flags) { S.OQ(4) = bt.par_adiab_inv[idx]; S.EC(0) = bt.counter_vpar_0[idx]; S.EC(1) = bt.counter_phi_0[idx]; }
    }
    return true;
  }
}

// EXT = 2: optional quantities and orbit events of a push that the fast path completed
template <int K, int PHI, int NT>
__device__ __forceinline__ void lane_ext2_after_fast(const Batch &bt, const LaneSlots<NT> &S, PolyPusher<K, PHI, 2> &P,
                                                     const PushOut &o)
{
  if (bt.oq_mask) {
#pragma unroll
    for (int q = 0; q < 4; q++) S.OQ(q) = S.OQ(q) + P.oq[q];
  }
  if constexpr (K >= 2) {
    if (bt.ev_flags && !o.finished) {
      EvState es;
      es.flags = bt.ev_flags; es.nskip_p = bt.n_skip_phi_0; es.nskip_v = bt.n_skip_vpar_0;
      es.J = S.OQ(4); es.cnt_v = S.EC(0); es.cnt_p = S.EC(1);
      P.events_after_push(S.D(LS_VPAR), o, es);
      S.OQ(4) = es.J; S.EC(0) = es.cnt_v; S.EC(1) = es.cnt_p;
      if (es.n) emit_events(bt, S.Idx(), S.Npush(), es, bt.t_step - (S.D(LS_TREM) - o.t_pass));
    }
  }
}
// EXT = 2: the complete ladder with optional quantities and events
template <int K, int PHI, int NT, int EXT = 2>
__device__ __forceinline__ PushOut lane_ext2_full(const MeshDev &m, const Batch &bt, const LaneSlots<NT> &S, int32_t ind_tetr,
                                                  int32_t iface)
{
  double *lst = nullptr;
  if constexpr (EXT == 5) lst = bt.lst + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * (size_t)bt.lst_cap * 5;
  const PushOutX ox = push_full_call_x<K, PHI, EXT>(&m, S.D(LS_PERPINV), ind_tetr, iface, S.D(LS_X0), S.D(LS_X1), S.D(LS_X2),
                                                    S.D(LS_VPAR), S.D(LS_TREM), bt.oq_mask, bt.ev_flags, bt.n_skip_phi_0,
                                                    bt.n_skip_vpar_0, S.OQ(4), S.EC(0), S.EC(1), lst, bt.lst_cap);
#pragma unroll
  for (int q = 0; q < 4; q++) S.OQ(q) = S.OQ(q) + ox.oq[q];
  if (bt.ev_flags) {
    S.OQ(4) = ox.es.J; S.EC(0) = ox.es.cnt_v; S.EC(1) = ox.es.cnt_p;
    if (ox.es.n) emit_events(bt, S.Idx(), S.Npush(), ox.es, bt.t_step - (S.D(LS_TREM) - ox.o.t_pass));
  }
  return ox.o;
}

// Book-keeping after a push (orbit_timestep_gorilla.f90:129-142): loop state, trace, counters; when the particle has
// finished its time step or is lost, its state goes back to the caller's arrays.  Returns true in that case (the lane
// needs a new particle).  ind_prev = the tetrahedron the push started in (ind_tetr_save of the reference).
template <int PHI, int EXT, int NT, bool ST>
__device__ __forceinline__ bool lane_after_push(const MeshDev &m, const Batch &bt, const LaneSlots<NT, ST> &S, const PushOut &o,
                                                int ind_prev, int32_t &ind_tetr, int32_t &iface)
{
  S.D(LS_X0) = o.x[0]; S.D(LS_X1) = o.x[1]; S.D(LS_X2) = o.x[2];
  S.D(LS_VPAR) = o.vpar;
  if (o.z_save_set) { S.D(LS_ZS0) = o.z_save[0]; S.D(LS_ZS1) = o.z_save[1]; S.D(LS_ZS2) = o.z_save[2]; }
  ind_tetr = o.ind_tetr;
  iface = o.iface;
  const long long npush = S.Npush();
  if (bt.trace_cap > 0 && npush < bt.trace_cap) {
    const long long idx = S.Idx();
    bt.trace_tetr[idx * bt.trace_cap + npush] = ind_tetr;
    bt.trace_face[idx * bt.trace_cap + npush] = iface;
  }
  S.Npush() = npush + 1;
  if (o.fallback) {
    if (o.fallback & 1) S.C(LC_FB0) = S.C(LC_FB0) + 1;
    if (o.fallback & 2) S.C(LC_FB1) = S.C(LC_FB1) + 1;
    if (o.fallback & 4) S.C(LC_FB2) = S.C(LC_FB2) + 1;
    if (o.fallback & 8) S.C(LC_FB3) = S.C(LC_FB3) + 1;
    if constexpr (EXT == 3 || EXT == 5) {
      if (o.fallback & 16) S.C(LC_ADAPT) = S.C(LC_ADAPT) + 1;
    }
  }
  const double t_remain = S.D(LS_TREM) - o.t_pass;
  S.D(LS_TREM) = t_remain;
  if constexpr (EXT == 2 || EXT == 5) {
    if (bt.ev_flags & 8) {   // boole_full_orbit: the orbit point after every n_skip_full_orbit-th push (gorilla_plot_mod.f90:553-579)
      const long long cnt = npush + 1;   // counter_tetrahedron_passes
      if (cnt / bt.n_skip_full_orbit * bt.n_skip_full_orbit == cnt) {
        const unsigned long long slot = atomicAdd(bt.ev_count, 1ull);
        if ((long long)slot < bt.ev_cap) {
          const double zs[3] = {S.D(LS_ZS0), S.D(LS_ZS1), S.D(LS_ZS2)};
          gorilla_event *e = bt.events + slot;
          e->particle = S.Idx();
          e->kind = GORILLA_EVENT_FULL_ORBIT;
          e->counter = (int32_t)cnt;
          e->push = npush;
          e->x[0] = o.x[0]; e->x[1] = o.x[1]; e->x[2] = o.x[2];
          orbit_point_invariants(m, ind_prev, zs, o.vpar, S.D(LS_PERPINV), e->value[0], e->value[1]);
          e->t = bt.t_step - t_remain;
        }
      }
    }
  }
  if (!(o.finished || ind_tetr == -1)) return false;
  // :142  vperp = vperp_func(z_save, perpinv, ind_tetr_save)
  const long long idx = S.Idx();
  const double pinv = S.D(LS_PERPINV);
  const double zs[3] = {S.D(LS_ZS0), S.D(LS_ZS1), S.D(LS_ZS2)};
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
  if constexpr (EXT == 2 || EXT == 5) {
    if (bt.optq) {
#pragma unroll
      for (int q = 0; q < 4; q++) bt.optq[4 * idx + q] = S.OQ(q);
    }
    if (bt.ev_flags) { bt.par_adiab_inv[idx] = S.OQ(4); bt.counter_vpar_0[idx] = S.EC(0); bt.counter_phi_0[idx] = S.EC(1); }
  }
  S.Cpush() = S.Cpush() + (unsigned long long)(npush + 1);
  if (o.finished) {
    S.C(LC_FIN) = S.C(LC_FIN) + 1;
  } else {
    // lost: through a boundary face (hand-over to neighbour -1; the position is the exit point) or removed by the pusher
    // (no valid exit time / trouble shooting failed: z_save not set).  Flux-coordinate grids have two boundaries: the inner
    // annulus edge s = sfc_s_min and the outer surface s = 1 (SURVEY 8d config 3 reports them separately).
    S.C(LC_LOST) = S.C(LC_LOST) + 1;
    if (!o.z_save_set) S.C(LC_FAILED) = S.C(LC_FAILED) + 1;
    else if (m.coord_system == 2 && o.x[0] < 0.5 * (m.sfc_s_min + 1.0)) S.C(LC_LOST_INNER) = S.C(LC_LOST_INNER) + 1;
  }
  return true;
}

// counters: warp reduce, one atomic per warp and counter
template <int NT, bool ST>
__device__ __forceinline__ void lane_reduce_counters(const Batch &bt, const LaneSlots<NT, ST> &S, unsigned lane)
{
  __syncwarp();
  unsigned long long v[11] = {S.Cpush(), S.C(LC_LOST), S.C(LC_FIN), S.C(LC_FB0), S.C(LC_FB1), S.C(LC_FB2), S.C(LC_FB3),
                              S.C(LC_ADAPT), S.C(LC_LOST_INNER), S.C(LC_FAILED), S.C(LC_LOST_PREV)};
  const int slot[11] = {CTR_PUSHES, CTR_LOST, CTR_FINISHED, CTR_FB0, CTR_FB1, CTR_FB2, CTR_FB3, CTR_ADAPT, CTR_LOST_INNER,
                        CTR_FAILED, CTR_LOST_PREV};
#pragma unroll
  for (int k = 0; k < 11; k++) {
    unsigned long long s = v[k];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
    if (lane == 0 && s) atomicAdd(bt.ctr + slot[k], s);
  }
}

// EXT = 1: Hamiltonian time tracing (i_time_tracing_option = 2); EXT = 2: time tracing option read at run time plus the
// optional quantities of pusher_tetra_poly; the plain variant (EXT = 0) is the hot path of the default settings and
// carries none of that code (the optional-quantity code alone costs the order-2 kernel ~400 bytes of spills).
// GATHER = 1: the geom / bpart sub-records reach the lane through the bulk-copy engine and a shared-memory slot, prefetched
// one push ahead (gb_mesh.cuh); 48 KB of dynamic shared memory per CTA on top of the lane slots => three CTAs per SM.
// GATHER = 2: the same slots filled by the warp-cooperative cp.async gather (gb_mesh.cuh).
#define GB_BULK_SMEM ((size_t)GB_THREADS * (GB_BULK_STRIDE + 16))
constexpr size_t gb_gather_smem(int gather, int phi)
{
  return gather == 1 ? GB_BULK_SMEM : gather == 2 ? (size_t)GB_THREADS * coop_smem_per_thread(phi) : 0;
}
// CTAs per SM: what the shared memory holds (GATHER = 2 with PHI = 1 / 2 stages 532 / 724 bytes per lane)
constexpr int gb_gather_min_blocks(int gather, int phi) { return gather == 2 && phi >= 1 ? 2 : 3; }
template <int K, int PHI, int EXT = 0, int GATHER = 0>
__global__ void __launch_bounds__(GB_THREADS, GATHER ? gb_gather_min_blocks(GATHER, PHI) : gb_min_blocks(K, EXT)) orbit_kernel(const __grid_constant__ MeshDev m, const Batch bt)
{
  constexpr bool ALL_STAGED = (GATHER == 2 && PHI >= 1);   // no shared-memory stash (see LaneSlots)
  using Slots = LaneSlots<GB_THREADS, !ALL_STAGED>;
  __shared__ __align__(16) unsigned char s_raw[(EXT == 2 || EXT == 5) ? Slots::BYTES_EXT2 : Slots::BYTES];
  Slots S;
  S.carve(s_raw, EXT == 2 || EXT == 5);
  double local_stash[6];
  const unsigned lane = threadIdx.x & 31u;
  int32_t ind_tetr = -1, iface = -1;
  S.zero_counters();
  if constexpr (GATHER == 1) bulk_init();
  if constexpr (GATHER == 2) coop_init<PHI>();
  unsigned wmask = 0xffffffffu;   // GATHER = 2: the lanes of this warp that are still in the push loop

  // One lane = one particle at a time.  A lane whose particle is done refills itself at the end of the same loop
  // body and leaves the loop for good when the queue is empty, so the body has no "is this lane active" region (whose
  // convergence-barrier register was live, and spilled, across every push).
  bool active = lane_refill<PHI, EXT>(m, bt, S, lane, ind_tetr, iface);
  if constexpr (GATHER == 2) wmask = __ballot_sync(0xffffffffu, active);
  while (active) {
    if constexpr (GATHER == 2) coop_wait(wmask);   // the records requested during the previous push are in the slots
    S.IndSave() = ind_tetr;
    PushOut o;
    bool done = false;
    const double perpinv = S.D(LS_PERPINV);
    if constexpr (K == 0) {  // ipusher = 1: RK4 pusher
      if (!bt.force_full) {
        const double x[3] = {S.D(LS_X0), S.D(LS_X1), S.D(LS_X2)};
        RkPusher<PHI, (EXT == 2 ? 2 : 0)> R;
        R.P.r.set_stash(ALL_STAGED ? local_stash : S.Stash(), ALL_STAGED ? 1 : GB_THREADS, GATHER, wmask);
        R.init(&m, perpinv, ind_tetr, x, iface, S.D(LS_VPAR), S.D(LS_TREM));
        done = R.template push<true>(o);
      }
      if (!done)
        o = push_rk_full_call<PHI, (EXT == 2 ? 2 : 0)>(&m, perpinv, ind_tetr, iface, S.D(LS_X0), S.D(LS_X1), S.D(LS_X2),
                                                        S.D(LS_VPAR), S.D(LS_TREM));
      if constexpr (EXT == 2) {
        if (bt.ev_flags && !o.finished) {   // J_par / banana tips / toroidal mappings (par_adiab_inv_rk_mod)
          EvState es;
          es.flags = bt.ev_flags; es.nskip_p = bt.n_skip_phi_0; es.nskip_v = bt.n_skip_vpar_0;
          es.J = S.OQ(4); es.cnt_v = S.EC(0); es.cnt_p = S.EC(1); es.n = 0;
          es = rk_events_call<PHI>(&m, perpinv, ind_tetr, iface, S.D(LS_X0), S.D(LS_X1), S.D(LS_X2), S.D(LS_VPAR), S.D(LS_TREM),
                                   o, es);
          S.OQ(4) = es.J; S.EC(0) = es.cnt_v; S.EC(1) = es.cnt_p;
          if (es.n) emit_events(bt, S.Idx(), S.Npush(), es, bt.t_step - (S.D(LS_TREM) - o.t_pass));
        }
      }
    } else if constexpr (EXT == 5) {
      // adaptive sub-stepping with the list consumers: every push takes the complete path (the step lists live in global memory)
      o = lane_ext2_full<K, PHI, GB_THREADS, 5>(m, bt, S, ind_tetr, iface);
    } else {
      if (!bt.force_full) {
        const double x[3] = {S.D(LS_X0), S.D(LS_X1), S.D(LS_X2)};
        PolyPusher<K, PHI, EXT> P;
        P.mp = &m;
        P.perpinv = perpinv;
        if constexpr (EXT == 2) P.oq_mask = bt.oq_mask;
        P.r.set_stash(ALL_STAGED ? local_stash : S.Stash(), ALL_STAGED ? 1 : GB_THREADS, GATHER, wmask);
        done = P.push_fast(ind_tetr, iface, x, S.D(LS_VPAR), S.D(LS_TREM), o, &S.D(LS_TREM));
        if constexpr (EXT == 2) {
          if (done) lane_ext2_after_fast<K, PHI>(bt, S, P, o);
        }
      }
      if (!done) {
        if constexpr (EXT == 2)
          o = lane_ext2_full<K, PHI>(m, bt, S, ind_tetr, iface);
        else
          o = push_full_call<K, PHI, EXT>(&m, perpinv, ind_tetr, iface, S.D(LS_X0), S.D(LS_X1), S.D(LS_X2), S.D(LS_VPAR),
                                          S.D(LS_TREM));
      }
    }
    if (lane_after_push<PHI, EXT>(m, bt, S, o, S.IndSave(), ind_tetr, iface))
      active = lane_refill<PHI, EXT>(m, bt, S, lane, ind_tetr, iface);
    if constexpr (GATHER == 2) wmask = __ballot_sync(wmask, active);   // lanes that leave the loop drop out of the warp's gather
  }
  // no copy may still be on its way to this CTA's shared memory when it exits
  if constexpr (GATHER == 1) bulk_wait();
  if constexpr (GATHER == 2) asm volatile("cp.async.wait_all;" ::: "memory");
  lane_reduce_counters(bt, S, lane);
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

// group-wide OR of a per-thread predicate + barrier.  Every lane of the group's four warps reaches it (lanes never leave the
// push loop on their own); __syncwarp() re-converges a warp whose lanes come out of divergent code first.
// Two forms.  Shipped: one barrier.red.or with a thread count (one instruction).  compute-sanitizer's synccheck reports that
// form as "divergent thread(s) in block" -- it is the partial-count barrier.red it objects to: with GB_GROUP_BARRIER_SMEM=1
// the OR goes through shared memory (one ballot word per warp, double-buffered by a per-warp phase so that one barrier per
// call is enough: a warp can only overwrite its word of phase p two calls later, i.e. after a barrier that every reader of
// the old value has passed) behind a plain named barrier.sync, the same tests pass under synccheck, results are identical,
// and orders 3/4 run 3-11 % slower (profiles/r02_compute_sanitizer.txt).
#ifndef GB_GROUP_BARRIER_SMEM
#define GB_GROUP_BARRIER_SMEM 0
#endif
#if GB_GROUP_BARRIER_SMEM
__device__ __forceinline__ volatile unsigned *group_flags()
{
  __shared__ unsigned f[48];   // [phase][warp] ballot words, then the phase of every warp
  return f;
}
__device__ __forceinline__ void group_init()
{
  if ((threadIdx.x & 31u) == 0) group_flags()[32 + (threadIdx.x >> 5)] = 0u;
  __syncwarp();
}
__device__ __forceinline__ bool group_any(bool pred, int bar_id)
{
  volatile unsigned *f = group_flags();
  const unsigned w = threadIdx.x >> 5, grp = w & 3u;
  __syncwarp();
  const unsigned b = __ballot_sync(0xffffffffu, pred);
  const unsigned ph = f[32 + w];
  __syncwarp();   // every lane has read the phase before lane 0 flips it
  if ((threadIdx.x & 31u) == 0) {
    f[ph * 16 + w] = b;
    f[32 + w] = ph ^ 1u;
  }
  asm volatile("barrier.sync %0, %1;" ::"r"(bar_id), "r"(GBG_GROUP) : "memory");
  const unsigned r = f[ph * 16 + grp] | f[ph * 16 + grp + 4] | f[ph * 16 + grp + 8] | f[ph * 16 + grp + 12];
  return r != 0u;
}
#else
__device__ __forceinline__ void group_init() {}
__device__ __forceinline__ bool group_any(bool pred, int bar_id)
{
  __syncwarp();
  int r;
  asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.s32 p, %1, 0;\n\tbarrier.red.or.pred q, %2, %3, p;\n\tselp.s32 %0, 1, 0, q;\n\t}"
               : "=r"(r) : "r"((int)pred), "r"(bar_id), "r"(GBG_GROUP) : "memory");
  return r != 0;
}
#endif

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

// ---- the same solves, RE-BINNED between iterations ---------------------------------------------------------------------
// ncu on the lock-step kernel: 10 of 32 lanes are active per instruction inside the solver.  Two causes: (1) the lanes of a
// warp are in different solver modes (a Laguerre iteration with its square root and Adams bound costs ~3x a Newton iteration,
// and the warp pays the most expensive mode present), (2) the 128 solves of a group finish after 15..40 iterations and the
// group waits for the last one with ever emptier warps.  Here the state of every solve lives in shared memory (24 doubles + 3
// words), and before every iteration the solves that are still running are counting-sorted by mode and handed out densely:
// warp 0 of the group takes the first 32, warp 1 the next 32, ...; a warp without work skips the iteration, a warp with work
// runs (almost) one mode.  The arithmetic of an iteration is SgSolver::step() as before, so results are identical.
#define GBR_ND 24   // q0..q3 | lambda | work[0..3] | roots[0..3] | root | stopping_crit2 ; field 0 holds tau once done
struct RebinSlots {
  double (*d)[GBG_THREADS];   // [GBR_ND][slot]
  int (*w)[GBG_THREADS];      // [3][slot]: packed state, i | j << 16, iter
  int *order;                 // [slot]
  int *wcnt;                  // [16 warps][4]
  static constexpr size_t BYTES = (size_t)GBG_THREADS * (GBR_ND * 8 + 3 * 4 + 4) + 16 * 4 * 4;
  __device__ __forceinline__ void carve(unsigned char *base)
  {
    d = reinterpret_cast<double (*)[GBG_THREADS]>(base);
    w = reinterpret_cast<int (*)[GBG_THREADS]>(d + GBR_ND);
    order = reinterpret_cast<int *>(w + 3);
    wcnt = order + GBG_THREADS;
  }
};
enum { GBR_DONE_BIT = 1 << 14 };
__device__ __forceinline__ void rebin_pack(const RebinSlots &R, int s, const SgSolver &S, bool all)
{
  R.w[0][s] = S.deg | (S.n << 3) | (S.phase << 6) | (S.pol << 8) | (S.mode << 11) | ((int)S.good_to_go << 13) |
              (S.done ? GBR_DONE_BIT : 0);
  R.w[1][s] = S.i | (S.j << 16);
  R.w[2][s] = S.iter;
  R.d[21][s] = S.root.re; R.d[22][s] = S.root.im; R.d[23][s] = S.stopping_crit2;
  if (all) {   // work / roots only change when a search or a polish ends
#pragma unroll
    for (int k = 0; k < 4; k++) {
      R.d[5 + 2 * k][s] = S.work[k].re; R.d[6 + 2 * k][s] = S.work[k].im;
      R.d[13 + 2 * k][s] = S.roots[k].re; R.d[14 + 2 * k][s] = S.roots[k].im;
    }
  }
}
__device__ __forceinline__ void rebin_unpack(const RebinSlots &R, int s, SgSolver &S)
{
  const int w0 = R.w[0][s], w1 = R.w[1][s];
  S.deg = w0 & 7; S.n = (w0 >> 3) & 7; S.phase = (w0 >> 6) & 3; S.pol = (w0 >> 8) & 7; S.mode = (w0 >> 11) & 3;
  S.good_to_go = (w0 >> 13) & 1; S.done = false;
  S.i = w1 & 0xffff; S.j = (w1 >> 16) & 0xffff; S.iter = R.w[2][s];
  const cd zero = mk(0.0, 0.0), one = mk(1.0, 0.0);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    // the polynomial is real and monic: poly[k] = (q_k, 0) below the degree, (1, 0) at it, 0 above (SgSolver::start)
    S.poly[k] = (k < S.deg) ? mk(R.d[k][s], 0.0) : (k == S.deg) ? one : zero;
    S.work[k] = mk(R.d[5 + 2 * k][s], R.d[6 + 2 * k][s]);
    S.roots[k] = mk(R.d[13 + 2 * k][s], R.d[14 + 2 * k][s]);
  }
  S.poly[4] = (S.deg == 4) ? one : zero;
  S.work[4] = (S.deg == 4) ? one : zero;   // only read while n = 4, where it is the leading (1, 0)
  S.root = mk(R.d[21][s], R.d[22][s]);
  S.stopping_crit2 = R.d[23][s];
}
// mode class of a running solve: 0 Laguerre-type iteration (search in mode 2, fall-back search, polish), 1 SG, 2 Newton
__device__ __forceinline__ int rebin_class(int w0)
{
  if (w0 & GBR_DONE_BIT) return 3;
  const int phase = (w0 >> 6) & 3, mode = (w0 >> 11) & 3;
  return (phase != 0 || mode == 2) ? 0 : (mode == 1 ? 1 : 2);
}
__device__ __forceinline__ void group_sync(int bar_id)
{
  __syncwarp();
  asm volatile("barrier.sync %0, %1;" ::"r"(bar_id), "r"(GBG_GROUP) : "memory");
}
static __device__ __noinline__ double solve_group_rebin(bool busy, int deg, double q0, double q1, double q2, double q3,
                                                        double lambda, double tau_ready, int bar_id, RebinSlots R)
{
  const unsigned lane = threadIdx.x & 31u;
  const int w = (int)(threadIdx.x >> 5), grp = w & 3, wig = w >> 2;   // warps grp, grp+4, grp+8, grp+12 form the group
  const int gl = wig * 32 + (int)lane, gbase = grp * GBG_GROUP, own = gbase + gl;
  if (busy) {
    SgSolver S;
    cd poly[5];
    poly[0] = mk(q0, 0.0);
    poly[1] = mk(deg == 1 ? 1.0 : q1, 0.0);
    poly[2] = mk(deg == 2 ? 1.0 : q2, 0.0);
    poly[3] = mk(deg == 3 ? 1.0 : q3, 0.0);
    poly[4] = mk(1.0, 0.0);
    S.start(deg, poly);
    R.d[0][own] = q0; R.d[1][own] = q1; R.d[2][own] = q2; R.d[3][own] = q3; R.d[4][own] = lambda;
    rebin_pack(R, own, S, true);
  } else {
    R.w[0][own] = GBR_DONE_BIT;
    R.d[0][own] = tau_ready;
  }
  group_sync(bar_id);
  for (;;) {
    const int c = rebin_class(R.w[0][own]);
    const unsigned b0 = __ballot_sync(0xffffffffu, c == 0), b1 = __ballot_sync(0xffffffffu, c == 1),
                   b2 = __ballot_sync(0xffffffffu, c == 2);
    if (lane == 0) {
      R.wcnt[w * 4 + 0] = __popc(b0); R.wcnt[w * 4 + 1] = __popc(b1); R.wcnt[w * 4 + 2] = __popc(b2);
    }
    group_sync(bar_id);
    int tot[3] = {0, 0, 0}, before[3] = {0, 0, 0};
#pragma unroll
    for (int q = 0; q < 4; q++) {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const int v = R.wcnt[(grp + 4 * q) * 4 + k];
        tot[k] += v;
        if (q < wig) before[k] += v;
      }
    }
    const int total = tot[0] + tot[1] + tot[2];
    if (total == 0) break;   // the same for every thread of the group
    if (c < 3) {
      const unsigned mine = c == 0 ? b0 : c == 1 ? b1 : b2;
      const int off = (c == 0 ? 0 : c == 1 ? tot[0] : tot[0] + tot[1]) + (c == 0 ? before[0] : c == 1 ? before[1] : before[2]) +
                      __popc(mine & ((1u << lane) - 1u));
      R.order[gbase + off] = own;
    }
    group_sync(bar_id);
    if (gl < total) {
      const int s = R.order[gbase + gl];
      SgSolver S;
      rebin_unpack(R, s, S);
      const int n0 = S.n, ph0 = S.phase, pol0 = S.pol;
      const bool fin = S.step();
      rebin_pack(R, s, S, fin || S.n != n0 || S.phase != ph0 || S.pol != pol0);
      if (fin) R.d[0][s] = min_positive_real_root(S.deg, S.roots, R.d[4][s]);
    }
    group_sync(bar_id);
  }
  return R.d[0][own];
}

template <int K, int PHI, int EXT = 0>
__global__ void __launch_bounds__(GBG_THREADS, 1) orbit_kernel_g(const __grid_constant__ MeshDev m, const Batch bt)
{
  extern __shared__ __align__(16) unsigned char g_smem[];
  LaneSlots<GBG_THREADS> S;
  S.carve(g_smem, EXT == 2);
  RebinSlots RB;
  RB.carve(g_smem + (EXT == 2 ? LaneSlots<GBG_THREADS>::BYTES_EXT2 : LaneSlots<GBG_THREADS>::BYTES));
  const unsigned lane = threadIdx.x & 31u;
  const int bar_id = 1 + (int)((threadIdx.x >> 5) & 3u);   // warps w, w+4, w+8, w+12 share sub-partition w
  int32_t ind_tetr = -1, iface = -1;
  S.zero_counters();
  group_init();

  bool active = lane_refill<PHI, EXT>(m, bt, S, lane, ind_tetr, iface);
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
    P.r.set_stash(S.Stash(), GBG_THREADS);
    if (active && !bt.force_full) {
      const double x[3] = {S.D(LS_X0), S.D(LS_X1), S.D(LS_X2)};
      P.perpinv = S.D(LS_PERPINV);
      begun = P.fast_begin(ind_tetr, iface, x, S.D(LS_VPAR), S.D(LS_TREM), t, iface_new, tau_max) && t.kind != 0;
    }
    double tau;
    if (EXT != 2 && bt.rebin)
      tau = solve_group_rebin(begun && t.kind == 2, t.deg, t.q[0], t.q[1], t.q[2], t.q[3], t.lambda, t.tau, bar_id, RB);
    else
      tau = solve_group(begun && t.kind == 2, t.deg, t.q[0], t.q[1], t.q[2], t.q[3], t.lambda, t.tau, bar_id);
    if (begun) {
      P.t_remain = S.D(LS_TREM);
      done = P.fast_end(tau, iface_new, tau_max, true, o);
    }
    if (active) {
      if constexpr (EXT == 2) {
        if (done) lane_ext2_after_fast<K, PHI>(bt, S, P, o);
        else o = lane_ext2_full<K, PHI>(m, bt, S, ind_tetr, iface);
      } else {
        if (!done)
          o = push_full_call<K, PHI, EXT>(&m, S.D(LS_PERPINV), ind_tetr, iface, S.D(LS_X0), S.D(LS_X1), S.D(LS_X2),
                                          S.D(LS_VPAR), S.D(LS_TREM));
      }
      if (lane_after_push<PHI, EXT>(m, bt, S, o, ind_tetr, ind_tetr, iface))
        active = lane_refill<PHI, EXT>(m, bt, S, lane, ind_tetr, iface);
    }
  }
  lane_reduce_counters(bt, S, lane);
}
constexpr size_t GBG_SMEM = LaneSlots<GBG_THREADS>::BYTES;
constexpr size_t GBG_SMEM_EXT = LaneSlots<GBG_THREADS>::BYTES_EXT2;

// ----------------------------------------------------------------------------------------------------
// energy_tot_func, p_phi_func (SRC/supporting_functions_mod.f90:279-301, 377-408) and perpinv = -vperp^2/(2|B|)
// (SRC/orbit_timestep_gorilla.f90:77) of one particle in tetrahedron it (1-based)
__device__ __forceinline__ void particle_invariants(const MeshDev &m, int32_t it, const double *x, double vl, double vp,
                                                    double &e, double &p, double &mu)
{
  const double *pg = m.geom + ((int64_t)it - 1) * GEOM_ND;
  const double *pb = m.bpart + ((int64_t)it - 1) * BPART_ND;
  const double *pc = m.cold + ((int64_t)it - 1) * COLD_ND;
  const double z[3] = {x[0] - pg[0], x[1] - pg[1], x[2] - pg[2]};
  const double gB[3] = {pb[B_GB], pb[B_GB + 1], pb[B_GB + 2]};
  const double bmod = pb[B_BMOD1] + dot3(gB, z);
  mu = -0.5 * (vp * vp) / bmod;
  const double vperp_e = sqrt(2.0 * fabs(mu) * bmod);
  double phi = 0.0;
  if (m.phi) {
    const double *pp = m.phi + ((int64_t)it - 1) * PHI_ND;
    const double gP[3] = {pp[P_GPHI], pp[P_GPHI + 1], pp[P_GPHI + 2]};
    phi = pp[P_PHI1] + dot3(gP, z);
  }
  e = m.particle_mass / 2.0 * (vperp_e * vperp_e + vl * vl) + m.particle_charge * phi;
  const double *ps = m.se ? m.se + ((int64_t)it - 1) * SE_ND : nullptr;
  if (ps) {  // :299
    const double g2[3] = {ps[S_GV2EMOD], ps[S_GV2EMOD + 1], ps[S_GV2EMOD + 2]};
    e = e + 0.5 * m.particle_mass * (ps[S_V2EMOD1] + dot3(z, g2));
  }
  const double gh[3] = {pc[C_GHPHI], pc[C_GHPHI + 1], pc[C_GHPHI + 2]};
  const double gA[3] = {pc[C_GAPHI], pc[C_GAPHI + 1], pc[C_GAPHI + 2]};
  p = m.particle_mass * vl * (pc[C_HPHI1] + dot3(gh, z)) + m.particle_mass / m.cm_over_e * (pc[C_APHI1] + dot3(gA, z));
  if (ps) {  // :402-406 (cylindrical coordinates: phi is the second covariant component)
    const double gv[3] = {ps[S_GVE2], ps[S_GVE2 + 1], ps[S_GVE2 + 2]};
    p = p + m.particle_mass * (ps[S_VE2_1] + dot3(z, gv));
  }
}

// device partials of one diagnostics reduction: [0..2] max |E/E0-1|, |mu/mu0-1|, |p_phi/p_phi0-1| ; [3..5] the sums of their
// squares ; then as int64: [6] particles sampled, [7] particles in the batch, [8..] the accumulated counters (CTR_N)
enum { DG_MAX = 0, DG_SUM = 3, DG_NSAMP = 6, DG_NPART = 7, DG_CTR = 8, GB_DIAG_ND = DG_CTR + CTR_N };

// One orbit_timestep* call in flight: its own device counter block (incl. the work-queue cursor) and timing events, so that
// calls issued on different streams of one handle do not share a cursor.  The handle keeps a small ring; a slot is reused
// only after the call that used it last has completed on the device.
#define GB_NSLOTS 8
struct CallSlot {
  unsigned long long *d_ctr = nullptr;   // [CTR_N]
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, done = nullptr;
  bool used = false, have_find_time = false, have_push_time = false;
  int64_t n = 0;
};

struct gorilla_b200_handle {
  int device = 0;
  int num_sms = 0;
  MeshDev mesh{};
  gorilla_settings settings{};
  double *d_geom = nullptr, *d_bpart = nullptr, *d_phi = nullptr, *d_cold = nullptr, *d_se = nullptr, *d_ham = nullptr, *d_skew = nullptr;
  double *d_poly4 = nullptr;   // tetra_physics_poly4 records (i_precomp = 1, 2)
  double *d_rec44 = nullptr;   // geom + bpart as one contiguous record per tetrahedron (bulk-copy / cooperative gather only)
  int rec_nd = 0;              // doubles per record in d_rec44: 44 (bulk copy) or coop_nd(PHI) = 48 / 96 (cooperative gather)
  double *d_lst = nullptr;     // EXT = 5 kernels: per-thread step lists
  size_t lst_bytes = 0;
  cudaEvent_t lst_done = nullptr;
  bool lst_used = false;
  double *s_oq = nullptr;   // [cap][4] scratch for the optional quantities (host-pointer entry point)
  uint32_t oq_mask = 0;
  int32_t *d_bin_start = nullptr, *d_bin_items = nullptr;
  CallSlot slots[GB_NSLOTS];
  int cur_slot = -1;                     // slot of the most recent orbit_timestep* call
  unsigned long long *d_acc = nullptr;   // [CTR_N] counters accumulated over all calls since gorilla_b200_diag_reset
  // scratch for the host-pointer entry points
  int64_t cap = 0;
  double *s_x = nullptr, *s_vpar = nullptr, *s_vperp = nullptr, *s_tro = nullptr, *s_e = nullptr, *s_p = nullptr, *s_mu = nullptr;
  int32_t *s_init = nullptr, *s_ind = nullptr, *s_iface = nullptr;
  int64_t *s_np = nullptr;
  int64_t trace_cap_elems = 0;
  int32_t *s_tr_t = nullptr, *s_tr_f = nullptr;
  // host-pointer event capture: per-particle state, event buffer, event counter
  int64_t ev_state_cap = 0, ev_cap = 0;
  double *s_J = nullptr;
  int32_t *s_cv = nullptr, *s_cp = nullptr;
  gorilla_event *s_ev = nullptr;
  uint64_t *s_nev = nullptr;
  // sort scratch (one sort at a time per handle: a sort on another stream waits for the previous one, sort_done)
  size_t sort_tmp_bytes = 0;
  void *sort_tmp = nullptr;
  int64_t sort_cap = 0;
  uint32_t *sort_keys_in = nullptr, *sort_keys_out = nullptr;
  int64_t *sort_vals_in = nullptr, *sort_perm = nullptr;
  cudaEvent_t sort_done = nullptr;
  bool sort_used = false;
  // gorilla_b200_resort_dev / in-library re-sort of the host-pointer path: gather scratch
  int64_t gather_cap = 0;
  double *g_d = nullptr;     // [gather_cap][3]
  int32_t *g_i = nullptr;    // [gather_cap]
  int32_t host_resort = 0;   // gorilla_b200_set_host_resort
  // diagnostics reduction (gorilla_b200_diag_reduce*): device partials + pinned host copy
  double *d_diag = nullptr;  // [GB_DIAG_ND]
  void *h_diag = nullptr;    // pinned
  // multi-GPU (gorilla_b200_comm_*): NCCL communicator of this rank, loaded at run time
  void *comm = nullptr;
  int32_t rank = 0, nranks = 1;
  int64_t l2_bytes = 0, hot_bytes = 0;   // L2 capacity of the device, bytes of hot records of the mesh
  int32_t bulk_gather = 0;               // gorilla_b200_set_gather: records through the bulk-copy engine (orders 1, 2 / RK4, EXT = 0)
  int ctas_per_sm = 0, threads_per_cta = 128;
  int force_full = 0;
  int use_group = 1;  // orders 3/4: lock-step solver kernel (orbit_kernel_g); 2 = with the solves re-binned by mode
};

// makes the handle's device current for the duration of an entry point (a handle belongs to the device it was created on)
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev)
  {
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess) { ok = false; return; }
    if (cur != dev) {
      ok = cudaSetDevice(dev) == cudaSuccess;
      if (ok) prev = cur;
    }
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define GB_ENTER(h)                                                                                          \
  DeviceGuard dg__((h)->device);                                                                             \
  if (!dg__.ok) { gbint::set_error("cannot make the handle's CUDA device current"); return GORILLA_ERR_CUDA; }

template <int K, int PHI, int EXT = 0>
int launch_orbit_t(gorilla_b200_handle *h, const Batch &bt, cudaStream_t s)
{
  if constexpr (EXT == 5) {
    // one CTA per SM: each thread owns a step list of 3 * max_n_intermediate_steps entries (manage_intermediate_steps_arrays,
    // pusher_tetra_poly.f90:98-101) of 40 bytes in global scratch
    int64_t grid5 = h->num_sms;
    const int64_t need5 = (bt.n + GB_THREADS - 1) / GB_THREADS;
    if (grid5 > need5) grid5 = need5;
    if (grid5 < 1) grid5 = 1;
    const int32_t cap = 3 * h->settings.max_n_intermediate_steps;
    const size_t bytes = (size_t)grid5 * GB_THREADS * (size_t)cap * 5 * sizeof(double);
    if (bytes > ((size_t)64 << 30))
      return gbint::fail(GORILLA_ERR_UNSUPPORTED, "adaptive sub-stepping with Hamiltonian time / optional quantities / events: the step lists "
                                                  "(threads x 3 max_n_intermediate_steps x 40 B) exceed 64 GB; lower max_n_intermediate_steps");
    if (!h->lst_done) GB_CUDA(cudaEventCreateWithFlags(&h->lst_done, cudaEventDisableTiming));
    if (bytes > h->lst_bytes) {
      GB_CUDA(cudaDeviceSynchronize());   // a launch on another stream may still be using the old region
      if (h->d_lst) GB_CUDA(cudaFree(h->d_lst));
      h->d_lst = nullptr; h->lst_bytes = 0;
      GB_CUDA(cudaMalloc((void **)&h->d_lst, bytes));
      h->lst_bytes = bytes;
    }
    Batch b5 = bt;
    b5.lst = h->d_lst;
    b5.lst_cap = cap;
    // one scratch region per handle: a launch on another stream waits for the previous one to finish with it
    if (h->lst_used) GB_CUDA(cudaStreamWaitEvent(s, h->lst_done, 0));
    orbit_kernel<K, PHI, 5><<<(unsigned)grid5, GB_THREADS, 0, s>>>(h->mesh, b5);
    gbint::count_launch(1);
    GB_CUDA(cudaGetLastError());
    GB_CUDA(cudaEventRecord(h->lst_done, s));
    h->lst_used = true;
    return GORILLA_OK;
  } else {
  if constexpr (K >= 3) {
    if (h->use_group) {
      const size_t smem_g = EXT == 2 ? GBG_SMEM_EXT : (bt.rebin ? GBG_SMEM + RebinSlots::BYTES : GBG_SMEM);
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
  if constexpr (EXT == 0 && (K == 0 || K == 2)) {
    // force_full (test hook) never reaches the warp-level gather call of the fast path: those launches use the plain kernel
    if (h->bulk_gather && h->threads_per_cta == GB_THREADS && !(h->bulk_gather == 2 && bt.force_full)) {
      auto launch_gather = [&](auto kernel, size_t smem) -> int {
        GB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        int per_sm_b = h->ctas_per_sm;
        if (per_sm_b <= 0) {
          GB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_b, kernel, GB_THREADS, smem));
          if (per_sm_b < 1) per_sm_b = 1;
        }
        int64_t grid_b = (int64_t)h->num_sms * per_sm_b;
        const int64_t need_b = (bt.n + GB_THREADS - 1) / GB_THREADS;
        if (grid_b > need_b) grid_b = need_b;
        if (grid_b < 1) grid_b = 1;
        kernel<<<(unsigned)grid_b, GB_THREADS, smem, s>>>(h->mesh, bt);
        gbint::count_launch(1);
        GB_CUDA(cudaGetLastError());
        return GORILLA_OK;
      };
      if (h->bulk_gather == 2) return launch_gather(orbit_kernel<K, PHI, EXT, 2>, gb_gather_smem(2, PHI));
      return launch_gather(orbit_kernel<K, PHI, EXT, 1>, GB_BULK_SMEM);
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
}
