// gb_diag.cu -- what sits around the push kernels behind the C ABI: particle re-sorting by tetrahedron, the diagnostics
// reduction (counters + conservation of energy / magnetic moment / toroidal momentum) and the multi-GPU communicator.
//
// Reference (paths relative to the GORILLA tree):
//   counters     counter_tetrahedron_passes  SRC/gorilla_plot_mod.f90:550 ; lost particles :290-294,488-491
//   invariants   energy_tot_func / p_phi_func SRC/supporting_functions_mod.f90:279-301,377-408 ; perpinv
//                SRC/orbit_timestep_gorilla.f90:77 ; written per time step by gorilla_plot (:603) for the user to compare
// The reference is a single OpenMP process: it has no reduction over processes.  Here particles shard over the GPUs of one
// box with the mesh replicated; the only exchange of the path is the reduction of these few hundred bytes,
// done with NCCL (loaded at run time: the library has no link-time dependency on it).
#include <dlfcn.h>
#include <string.h>
#include <math.h>
#include <cub/device/device_radix_sort.cuh>
#include <nccl.h>
#include "gb_internal.cuh"

using gbint::fail;

// ---------------------------------------------------------------------------------------------------- re-sorting
__global__ void sort_keys_kernel(int64_t n, const int32_t *ind_tetr, uint32_t *keys, int64_t *vals)
{
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t t = ind_tetr[i];
    keys[i] = t < 1 ? 0xffffffffu : (uint32_t)t;   // lost particles last
    vals[i] = i;
  }
}

// the six state arrays of a batch, permuted in one pass: out[i] = in[perm[i]] (INVERSE: out[perm[i]] = in[i])
template <bool INVERSE>
__global__ void permute_state_kernel(int64_t n, const int64_t *perm, const double *x, const double *vpar, const double *vperp,
                                     const int32_t *init, const int32_t *ind, const int32_t *iface, double *ox, double *ovpar,
                                     double *ovperp, int32_t *oinit, int32_t *oind, int32_t *oiface)
{
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = perm[i];
    const int64_t src = INVERSE ? i : p, dst = INVERSE ? p : i;
    ox[3 * dst] = x[3 * src]; ox[3 * dst + 1] = x[3 * src + 1]; ox[3 * dst + 2] = x[3 * src + 2];
    ovpar[dst] = vpar[src];
    ovperp[dst] = vperp[src];
    if (init) oinit[dst] = init[src];
    oind[dst] = ind[src];
    oiface[dst] = iface[src];
  }
}
template <typename T, bool INVERSE>
__global__ void permute_one_kernel(int64_t n, const int64_t *perm, const T *in, T *out)
{
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = perm[i];
    if (INVERSE) out[p] = in[i];
    else out[i] = in[p];
  }
}

static int grid_for(const gorilla_b200_handle *h, int64_t n, int block)
{
  int64_t g = (n + block - 1) / block;
  if (g > (int64_t)h->num_sms * 8) g = (int64_t)h->num_sms * 8;
  return (int)(g < 1 ? 1 : g);
}

static int ensure_sort_scratch(gorilla_b200_handle *h, int64_t n, cudaStream_t s)
{
  if (n <= h->sort_cap) return GORILLA_OK;
  if (h->sort_used) GB_CUDA(cudaEventSynchronize(h->sort_done));
  cudaFree(h->sort_keys_in); cudaFree(h->sort_keys_out); cudaFree(h->sort_vals_in); cudaFree(h->sort_tmp); cudaFree(h->sort_perm);
  h->sort_keys_in = h->sort_keys_out = nullptr; h->sort_vals_in = h->sort_perm = nullptr; h->sort_tmp = nullptr; h->sort_cap = 0;
  GB_CUDA(cudaMalloc((void **)&h->sort_keys_in, (size_t)n * sizeof(uint32_t)));
  GB_CUDA(cudaMalloc((void **)&h->sort_keys_out, (size_t)n * sizeof(uint32_t)));
  GB_CUDA(cudaMalloc((void **)&h->sort_vals_in, (size_t)n * sizeof(int64_t)));
  GB_CUDA(cudaMalloc((void **)&h->sort_perm, (size_t)n * sizeof(int64_t)));
  size_t bytes = 0;
  GB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, h->sort_keys_in, h->sort_keys_out, h->sort_vals_in, h->sort_perm,
                                          (int)n, 0, 32, s));
  GB_CUDA(cudaMalloc(&h->sort_tmp, bytes));
  h->sort_tmp_bytes = bytes;
  h->sort_cap = n;
  return GORILLA_OK;
}

// permutation that orders particles by tetrahedron index into perm (device); the handle's sort scratch is shared, so a sort
// issued on another stream first waits for the previous one
namespace gbint {
int sort_permutation(gorilla_b200_handle *h, int64_t n, const int32_t *ind_tetr, int64_t *perm /* nullptr: h->sort_perm */,
                     cudaStream_t s)
{
  if (n > 0x7fffffffLL) return fail(GORILLA_ERR_ARG, "sort_permutation: n too large");
  int rc = ensure_sort_scratch(h, n, s);
  if (rc) return rc;
  if (!perm) perm = h->sort_perm;
  if (h->sort_used) GB_CUDA(cudaStreamWaitEvent(s, h->sort_done, 0));
  sort_keys_kernel<<<grid_for(h, n, 256), 256, 0, s>>>(n, ind_tetr, h->sort_keys_in, h->sort_vals_in);
  count_launch(1);
  GB_CUDA(cudaGetLastError());
  size_t bytes = h->sort_tmp_bytes;
  GB_CUDA(cub::DeviceRadixSort::SortPairs(h->sort_tmp, bytes, h->sort_keys_in, h->sort_keys_out, h->sort_vals_in, perm, (int)n,
                                          0, 32, s));
  count_launch(4);
  GB_CUDA(cudaEventRecord(h->sort_done, s));
  h->sort_used = true;
  return GORILLA_OK;
}

int ensure_gather_scratch(gorilla_b200_handle *h, int64_t n)
{
  if (n <= h->gather_cap) return GORILLA_OK;
  if (h->sort_used) GB_CUDA(cudaEventSynchronize(h->sort_done));
  cudaFree(h->g_d); cudaFree(h->g_i);
  h->g_d = nullptr; h->g_i = nullptr; h->gather_cap = 0;
  GB_CUDA(cudaMalloc((void **)&h->g_d, (size_t)n * 5 * sizeof(double)));   // x(3), vpar, vperp
  GB_CUDA(cudaMalloc((void **)&h->g_i, (size_t)n * 3 * sizeof(int32_t)));  // init, ind_tetr, iface
  h->gather_cap = n;
  return GORILLA_OK;
}

// gather (INVERSE = false) or scatter back (true) the state arrays through perm into the handle's scratch, then copy the
// scratch over the originals: an in-place permutation of caller-owned arrays
int permute_state_inplace(gorilla_b200_handle *h, int64_t n, const int64_t *perm, bool inverse, double *x, double *vpar,
                          double *vperp, int32_t *init, int32_t *ind, int32_t *iface, cudaStream_t s)
{
  int rc = ensure_gather_scratch(h, n);
  if (rc) return rc;
  if (h->sort_used) GB_CUDA(cudaStreamWaitEvent(s, h->sort_done, 0));
  double *gx = h->g_d, *gv = h->g_d + 3 * n, *gw = h->g_d + 4 * n;
  int32_t *gb = h->g_i, *gt = h->g_i + n, *gf = h->g_i + 2 * n;
  if (inverse)
    permute_state_kernel<true><<<grid_for(h, n, 256), 256, 0, s>>>(n, perm, x, vpar, vperp, init, ind, iface, gx, gv, gw, gb, gt, gf);
  else
    permute_state_kernel<false><<<grid_for(h, n, 256), 256, 0, s>>>(n, perm, x, vpar, vperp, init, ind, iface, gx, gv, gw, gb, gt, gf);
  count_launch(1);
  GB_CUDA(cudaGetLastError());
  GB_CUDA(cudaMemcpyAsync(x, gx, (size_t)n * 3 * sizeof(double), cudaMemcpyDeviceToDevice, s));
  GB_CUDA(cudaMemcpyAsync(vpar, gv, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, s));
  GB_CUDA(cudaMemcpyAsync(vperp, gw, (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, s));
  if (init) GB_CUDA(cudaMemcpyAsync(init, gb, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
  GB_CUDA(cudaMemcpyAsync(ind, gt, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
  GB_CUDA(cudaMemcpyAsync(iface, gf, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
  GB_CUDA(cudaEventRecord(h->sort_done, s));   // the gather scratch is free again once these copies are done
  h->sort_used = true;
  return GORILLA_OK;
}
template <typename T>
int permute_one_inplace(gorilla_b200_handle *h, int64_t n, const int64_t *perm, bool inverse, T *a, cudaStream_t s)
{
  static_assert(sizeof(T) <= sizeof(double), "scratch is sized for doubles");
  int rc = ensure_gather_scratch(h, n);
  if (rc) return rc;
  if (h->sort_used) GB_CUDA(cudaStreamWaitEvent(s, h->sort_done, 0));
  T *g = reinterpret_cast<T *>(h->g_d);
  if (inverse) permute_one_kernel<T, true><<<grid_for(h, n, 256), 256, 0, s>>>(n, perm, a, g);
  else permute_one_kernel<T, false><<<grid_for(h, n, 256), 256, 0, s>>>(n, perm, a, g);
  count_launch(1);
  GB_CUDA(cudaGetLastError());
  GB_CUDA(cudaMemcpyAsync(a, g, (size_t)n * sizeof(T), cudaMemcpyDeviceToDevice, s));
  GB_CUDA(cudaEventRecord(h->sort_done, s));
  h->sort_used = true;
  return GORILLA_OK;
}
template int permute_one_inplace<double>(gorilla_b200_handle *, int64_t, const int64_t *, bool, double *, cudaStream_t);
template int permute_one_inplace<int64_t>(gorilla_b200_handle *, int64_t, const int64_t *, bool, int64_t *, cudaStream_t);
} // namespace gbint

extern "C" int gorilla_b200_sort_permutation_dev(gorilla_b200_handle *h, int64_t n, const int32_t *ind_tetr, int64_t *perm,
                                                 void *stream)
{
  if (!h || n < 0 || (n > 0 && (!ind_tetr || !perm))) return fail(GORILLA_ERR_ARG, "sort_permutation: null argument");
  if (n == 0) return GORILLA_OK;
  GB_ENTER(h);
  return gbint::sort_permutation(h, n, ind_tetr, perm, (cudaStream_t)stream);
}

extern "C" int gorilla_b200_resort_dev(gorilla_b200_handle *h, int64_t n, double *x, double *vpar, double *vperp,
                                       int32_t *boole_initialized, int32_t *ind_tetr, int32_t *iface, int32_t n_extra,
                                       double *const *extra, int64_t *perm_out, void *stream)
{
  if (!h || n < 0 || n_extra < 0 || (n_extra > 0 && !extra) || (n > 0 && (!x || !vpar || !vperp || !ind_tetr || !iface)))
    return fail(GORILLA_ERR_ARG, "gorilla_b200_resort_dev: null argument");
  if (n == 0) return GORILLA_OK;
  GB_ENTER(h);
  cudaStream_t s = (cudaStream_t)stream;
  int rc = gbint::sort_permutation(h, n, ind_tetr, nullptr, s);
  if (rc) return rc;
  rc = gbint::permute_state_inplace(h, n, h->sort_perm, false, x, vpar, vperp, boole_initialized, ind_tetr, iface, s);
  if (rc) return rc;
  for (int k = 0; k < n_extra; k++) {
    if (!extra[k]) continue;
    rc = gbint::permute_one_inplace<double>(h, n, h->sort_perm, false, extra[k], s);
    if (rc) return rc;
  }
  if (perm_out) GB_CUDA(cudaMemcpyAsync(perm_out, h->sort_perm, (size_t)n * sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
  GB_CUDA(cudaEventRecord(h->sort_done, s));
  return GORILLA_OK;
}

extern "C" int gorilla_b200_set_host_resort(gorilla_b200_handle *h, int32_t on)
{
  if (!h) return fail(GORILLA_ERR_ARG, "null handle");
  h->host_resort = on ? 1 : 0;
  return GORILLA_OK;
}

// ---------------------------------------------------------------------------------------------------- diagnostics
// device partials of one reduction (layout: enum DG_* in gb_internal.cuh)

__device__ __forceinline__ void atomic_max_nonneg(double *addr, double v)
{
  // non-negative IEEE doubles order like their bit patterns
  atomicMax(reinterpret_cast<unsigned long long *>(addr), (unsigned long long)__double_as_longlong(v));
}

__global__ void __launch_bounds__(256) diag_kernel(const __grid_constant__ MeshDev m, int64_t n, const double *x,
                                                   const double *vpar, const double *vperp, const int32_t *ind_tetr,
                                                   const double *e0, const double *p0, const double *mu0, double *out)
{
  double mx[3] = {0.0, 0.0, 0.0}, sq[3] = {0.0, 0.0, 0.0};
  long long ns = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t it = ind_tetr[i];
    if (it < 1) continue;
    double e, p, mu;
    particle_invariants(m, it, &x[3 * i], vpar[i], vperp[i], e, p, mu);
    const double ref[3] = {e0 ? e0[i] : NAN, mu0 ? mu0[i] : NAN, p0 ? p0[i] : NAN}, now[3] = {e, mu, p};
    bool any = false;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      // a reference of 0 (e.g. mu of a particle with v_perp = 0) or a non-finite one carries no relative drift
      if (!(fabs(ref[k]) <= DBL_MAX) || ref[k] == 0.0 || !(fabs(now[k]) <= DBL_MAX)) continue;
      const double d = fabs(now[k] / ref[k] - 1.0);
      mx[k] = fmax(mx[k], d);
      sq[k] += d * d;
      any = true;
    }
    if (any) ns++;
  }
  __shared__ double s_mx[3][8], s_sq[3][8];
  __shared__ long long s_ns[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      mx[k] = fmax(mx[k], __shfl_down_sync(0xffffffffu, mx[k], off));
      sq[k] += __shfl_down_sync(0xffffffffu, sq[k], off);
    }
    ns += __shfl_down_sync(0xffffffffu, ns, off);
  }
  if (lane == 0) {
    for (int k = 0; k < 3; k++) { s_mx[k][w] = mx[k]; s_sq[k][w] = sq[k]; }
    s_ns[w] = ns;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 1; q < 8; q++) {
      for (int k = 0; k < 3; k++) { mx[k] = fmax(mx[k], s_mx[k][q]); sq[k] += s_sq[k][q]; }
      ns += s_ns[q];
    }
    for (int k = 0; k < 3; k++) {
      atomic_max_nonneg(out + DG_MAX + k, mx[k]);
      atomicAdd(out + DG_SUM + k, sq[k]);
    }
    atomicAdd(reinterpret_cast<unsigned long long *>(out + DG_NSAMP), (unsigned long long)ns);
  }
}

__global__ void diag_pack_kernel(double *out, const unsigned long long *acc, long long n)
{
  const int k = threadIdx.x;
  unsigned long long *o = reinterpret_cast<unsigned long long *>(out);
  if (k == 0) o[DG_NPART] = (unsigned long long)n;
  if (k < CTR_N) o[DG_CTR + k] = acc[k];
}

// ---- NCCL, loaded at run time
namespace {
struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};
NcclApi g_nccl;
std::string g_nccl_err;

bool load_nccl()
{
  if (g_nccl.ok) return true;
  if (!g_nccl.lib) {
    const char *names[] = {getenv("GORILLA_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
      if (!nm || !*nm) continue;
      g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) {
      g_nccl_err = std::string("cannot load NCCL (libnccl.so.2): ") + (dlerror() ? dlerror() : "not found");
      return false;
    }
  }
#define GB_SYM(field, name)                                                        \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(g_nccl.lib, name)); \
  if (!g_nccl.field) { g_nccl_err = std::string("NCCL symbol missing: ") + name; return false; }
  GB_SYM(GetUniqueId, "ncclGetUniqueId")
  GB_SYM(CommInitRank, "ncclCommInitRank")
  GB_SYM(CommDestroy, "ncclCommDestroy")
  GB_SYM(AllReduce, "ncclAllReduce")
  GB_SYM(GroupStart, "ncclGroupStart")
  GB_SYM(GroupEnd, "ncclGroupEnd")
  GB_SYM(GetErrorString, "ncclGetErrorString")
#undef GB_SYM
  g_nccl.ok = true;
  return true;
}
int nccl_fail(const char *what, ncclResult_t r)
{
  std::string msg = std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error");
  return fail(GORILLA_ERR_CUDA, msg.c_str());
}
#define GB_NCCL(call)                                     \
  do {                                                    \
    ncclResult_t r__ = (call);                            \
    if (r__ != ncclSuccess) return nccl_fail(#call, r__); \
  } while (0)
} // namespace

static_assert(GORILLA_COMM_ID_BYTES == sizeof(ncclUniqueId), "gorilla_b200.h: GORILLA_COMM_ID_BYTES must be sizeof(ncclUniqueId)");

extern "C" int gorilla_b200_comm_unique_id(void *id)
{
  if (!id) return fail(GORILLA_ERR_ARG, "gorilla_b200_comm_unique_id: null argument");
  if (!load_nccl()) return fail(GORILLA_ERR_UNSUPPORTED, g_nccl_err.c_str());
  ncclUniqueId u;
  GB_NCCL(g_nccl.GetUniqueId(&u));
  memcpy(id, &u, sizeof(u));
  return GORILLA_OK;
}

extern "C" int gorilla_b200_comm_init(gorilla_b200_handle *h, const void *id, int32_t rank, int32_t nranks)
{
  if (!h || !id || nranks < 1 || rank < 0 || rank >= nranks) return fail(GORILLA_ERR_ARG, "gorilla_b200_comm_init: bad argument");
  if (h->comm) return fail(GORILLA_ERR_ARG, "gorilla_b200_comm_init: the handle already has a communicator");
  if (!load_nccl()) return fail(GORILLA_ERR_UNSUPPORTED, g_nccl_err.c_str());
  GB_ENTER(h);
  ncclUniqueId u;
  memcpy(&u, id, sizeof(u));
  ncclComm_t c = nullptr;
  GB_NCCL(g_nccl.CommInitRank(&c, nranks, u, rank));
  h->comm = c;
  h->rank = rank;
  h->nranks = nranks;
  return GORILLA_OK;
}

extern "C" int gorilla_b200_comm_free(gorilla_b200_handle *h)
{
  if (!h) return fail(GORILLA_ERR_ARG, "null handle");
  if (h->comm && g_nccl.ok) {
    GB_ENTER(h);
    g_nccl.CommDestroy((ncclComm_t)h->comm);
  }
  h->comm = nullptr;
  h->rank = 0;
  h->nranks = 1;
  return GORILLA_OK;
}

extern "C" int gorilla_b200_comm_allreduce_f64(gorilla_b200_handle *h, double *buf, int64_t count, int32_t op, void *stream)
{
  if (!h || count < 0 || (count > 0 && !buf) || op < 0 || op > 2) return fail(GORILLA_ERR_ARG, "comm_allreduce_f64: bad argument");
  if (!h->comm || count == 0) return GORILLA_OK;   // one rank: nothing to exchange
  GB_ENTER(h);
  const ncclRedOp_t ops[3] = {ncclSum, ncclMax, ncclMin};
  GB_NCCL(g_nccl.AllReduce(buf, buf, (size_t)count, ncclFloat64, ops[op], (ncclComm_t)h->comm, (cudaStream_t)stream));
  return GORILLA_OK;
}

// contiguous shards [r N/G, (r+1) N/G) (SURVEY.md 8e / BASELINE config 5)
extern "C" int gorilla_b200_shard_range(int64_t n_total, int32_t rank, int32_t nranks, int64_t *first, int64_t *count)
{
  if (n_total < 0 || nranks < 1 || rank < 0 || rank >= nranks || !first || !count)
    return fail(GORILLA_ERR_ARG, "gorilla_b200_shard_range: bad argument");
  const int64_t a = (int64_t)((__int128)n_total * rank / nranks), b = (int64_t)((__int128)n_total * (rank + 1) / nranks);
  *first = a;
  *count = b - a;
  return GORILLA_OK;
}

extern "C" int gorilla_b200_diag_reset(gorilla_b200_handle *h, void *stream)
{
  if (!h) return fail(GORILLA_ERR_ARG, "null handle");
  GB_ENTER(h);
  GB_CUDA(cudaMemsetAsync(h->d_acc, 0, CTR_N * sizeof(unsigned long long), (cudaStream_t)stream));
  return GORILLA_OK;
}

extern "C" int gorilla_b200_diag_reduce_dev(gorilla_b200_handle *h, int64_t n, const double *x, const double *vpar,
                                            const double *vperp, const int32_t *ind_tetr, const double *energy_ref,
                                            const double *p_phi_ref, const double *perpinv_ref, gorilla_diag *out, void *stream)
{
  if (!h || !out || n < 0 || (n > 0 && (!x || !vpar || !vperp || !ind_tetr)))
    return fail(GORILLA_ERR_ARG, "gorilla_b200_diag_reduce_dev: null argument");
  GB_ENTER(h);
  cudaStream_t s = (cudaStream_t)stream;
  GB_CUDA(cudaMemsetAsync(h->d_diag, 0, GB_DIAG_ND * sizeof(double), s));
  if (n > 0) {
    diag_kernel<<<grid_for(h, n, 256), 256, 0, s>>>(h->mesh, n, x, vpar, vperp, ind_tetr, energy_ref, p_phi_ref, perpinv_ref,
                                                    h->d_diag);
    gbint::count_launch(1);
    GB_CUDA(cudaGetLastError());
  }
  diag_pack_kernel<<<1, 32, 0, s>>>(h->d_diag, h->d_acc, (long long)n);
  gbint::count_launch(1);
  GB_CUDA(cudaGetLastError());
  if (h->comm) {
    // the path's only exchange: ~200 bytes per rank.  max of the drifts, sums of squares, integer counters.
    GB_NCCL(g_nccl.GroupStart());
    GB_NCCL(g_nccl.AllReduce(h->d_diag + DG_MAX, h->d_diag + DG_MAX, 3, ncclFloat64, ncclMax, (ncclComm_t)h->comm, s));
    GB_NCCL(g_nccl.AllReduce(h->d_diag + DG_SUM, h->d_diag + DG_SUM, 3, ncclFloat64, ncclSum, (ncclComm_t)h->comm, s));
    GB_NCCL(g_nccl.AllReduce(h->d_diag + DG_NSAMP, h->d_diag + DG_NSAMP, GB_DIAG_ND - DG_NSAMP, ncclInt64, ncclSum,
                             (ncclComm_t)h->comm, s));
    GB_NCCL(g_nccl.GroupEnd());
  }
  GB_CUDA(cudaMemcpyAsync(h->h_diag, h->d_diag, GB_DIAG_ND * sizeof(double), cudaMemcpyDeviceToHost, s));
  GB_CUDA(cudaStreamSynchronize(s));
  const double *d = static_cast<const double *>(h->h_diag);
  const int64_t *q = static_cast<const int64_t *>(h->h_diag);
  memset(out, 0, sizeof(*out));
  out->nranks = h->nranks;
  out->n_particles = q[DG_NPART];
  out->n_sampled = q[DG_NSAMP];
  const int64_t *c = q + DG_CTR;
  out->n_pushes = c[CTR_PUSHES];
  out->n_lost = c[CTR_LOST];
  out->n_lost_inner = c[CTR_LOST_INNER];
  out->n_failed = c[CTR_FAILED];
  out->n_lost_outer = c[CTR_LOST] - c[CTR_LOST_INNER] - c[CTR_FAILED];
  out->n_finished = c[CTR_FINISHED];
  for (int k = 0; k < 4; k++) out->n_fallback[k] = c[CTR_FB0 + k];
  out->n_adaptive = c[CTR_ADAPT];
  const double ns = out->n_sampled > 0 ? (double)out->n_sampled : 1.0;
  out->max_delta_energy = d[DG_MAX];
  out->max_delta_perpinv = d[DG_MAX + 1];
  out->max_delta_p_phi = d[DG_MAX + 2];
  out->rms_delta_energy = sqrt(d[DG_SUM] / ns);
  out->rms_delta_perpinv = sqrt(d[DG_SUM + 1] / ns);
  out->rms_delta_p_phi = sqrt(d[DG_SUM + 2] / ns);
  return GORILLA_OK;
}

// HOST-pointer variant (what a Fortran caller has): uploads the batch and the reference values into the handle's scratch.
extern "C" int gorilla_b200_diag_reduce(gorilla_b200_handle *h, int64_t n, const double *x, const double *vpar,
                                        const double *vperp, const int32_t *ind_tetr, const double *energy_ref,
                                        const double *p_phi_ref, const double *perpinv_ref, gorilla_diag *out)
{
  if (!h || !out || n < 0 || (n > 0 && (!x || !vpar || !vperp || !ind_tetr)))
    return fail(GORILLA_ERR_ARG, "gorilla_b200_diag_reduce: null argument");
  GB_ENTER(h);
  cudaStream_t s = nullptr;
  if (n > 0) {
    int rc = gbint::ensure_host_scratch(h, n);
    if (rc) return rc;
    const size_t nd = (size_t)n * sizeof(double);
    GB_CUDA(cudaMemcpyAsync(h->s_x, x, 3 * nd, cudaMemcpyHostToDevice, s));
    GB_CUDA(cudaMemcpyAsync(h->s_vpar, vpar, nd, cudaMemcpyHostToDevice, s));
    GB_CUDA(cudaMemcpyAsync(h->s_vperp, vperp, nd, cudaMemcpyHostToDevice, s));
    GB_CUDA(cudaMemcpyAsync(h->s_ind, ind_tetr, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    if (energy_ref) GB_CUDA(cudaMemcpyAsync(h->s_e, energy_ref, nd, cudaMemcpyHostToDevice, s));
    if (p_phi_ref) GB_CUDA(cudaMemcpyAsync(h->s_p, p_phi_ref, nd, cudaMemcpyHostToDevice, s));
    if (perpinv_ref) GB_CUDA(cudaMemcpyAsync(h->s_mu, perpinv_ref, nd, cudaMemcpyHostToDevice, s));
  }
  return gorilla_b200_diag_reduce_dev(h, n, h->s_x, h->s_vpar, h->s_vperp, h->s_ind, energy_ref ? h->s_e : nullptr,
                                      p_phi_ref ? h->s_p : nullptr, perpinv_ref ? h->s_mu : nullptr, out, s);
}
