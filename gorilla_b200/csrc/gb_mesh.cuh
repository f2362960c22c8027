// gb_mesh.cuh -- device-resident mesh: the reference's per-tetrahedron AoS records repacked as a
// structure of arrays of 32-byte-aligned SUB-RECORDS (one array per access group).
//
// Reference types being repacked:
//   type tetrahedron_physics  (142 doubles, SRC/tetra_physics_mod.f90:9-83)
//   type tetrahedron_grid     (20 int32,   SRC/tetra_grid_mod.f90:6-15)
//   type tetrahedron_physics_precomp_poly1 (SRC/tetra_physics_poly_precomp_mod.f90:8-15) is NOT
//   stored: n.alpha, n.beta, n.curlA are re-formed on the fly in the (rare) paths that use them.
//
// Why sub-records and not 60 scalar arrays: one crossing gathers the whole record of ONE random
// tetrahedron per lane.  A sector is 32 B, so a scalar SoA would move 60 sectors (1920 B) per
// crossing; sub-records move 4 + 7 (+5) sectors = 352 B (512 B with the electrostatic part),
// against 344 / 488 algorithmic bytes.  Groups are separate arrays so that the
// eps_Phi = 0 specialisation never touches the Phi group and the cold group is only read by the
// fall-back / diagnostic paths.
//
//   geom  [ntetr][16] : x1(3) dist_ref anorm(3,4)                                         128 B
//   bpart [ntetr][28] : bmod1 gB(3) curlA(3) curlh(3) gBxh1(3) gBxcurlA alpmat(3,3) spalpmat
//                       dt_dtau_const | neighbour_tetr(4) int32 | packed face/periodic flags  224 B
//   phi   [ntetr][20] : Phi1 gPhi(3) gPhixh1(3) gPhixcurlA betmat(3,3) spbetmat (2 pad)        160 B
//   cold  [ntetr][12] : tetra_dist_ref R1 Er_mod h_phi1 gh_phi(3) Aphi1 gAphi(3) (1 pad)        96 B
//   se    [ntetr][32] : strong-electric-field group (boole_strong_electric_field, cylindrical grids only):
//                       v2Emod_1 gv2Emod(3) gv2Emodxh1(3) gBxcurlvE gPhixcurlvE gv2EmodxcurlvE gv2EmodxcurlA
//                       curlvE(3) gammat(3,3) spgammat v_E_mod_average | vE2_1 gvE2(3) (invariants) (3 pad)   256 B
//                       (hot part = first 26 doubles = 7 sectors)
//   ham   [ntetr][8]  : type hamiltonian_time_type (tetra_physics_mod.f90:105-114): h1_in_curlA h1_in_curlh
//                       vec_mismatch_der(3) vec_parcurr_der(3); read at the end of a push by the EXT kernels only       64 B
// Matrices keep the Fortran column-major order: alpmat(i,j) -> [i + 3*j].
#pragma once
#include <stdint.h>
#include "gb_math.cuh"

namespace gb {

enum { GEOM_ND = 16, BPART_ND = 28, PHI_ND = 20, COLD_ND = 12, SE_ND = 32, HAM_ND = 8, SKEW_ND = 192 };
enum { B_BMOD1 = 0, B_GB = 1, B_CURLA = 4, B_CURLH = 7, B_GBXH1 = 10, B_GBXCURLA = 13, B_ALP = 14,
       B_SPALP = 23, B_DTDTAU = 24, B_TOPO = 25 };
enum { P_PHI1 = 0, P_GPHI = 1, P_GPHIXH1 = 4, P_GPHIXCURLA = 7, P_BET = 8, P_SPBET = 17 };
enum { S_V2EMOD1 = 0, S_GV2EMOD = 1, S_GV2EMODXH1 = 4, S_GBXCURLVE = 7, S_GPHIXCURLVE = 8, S_GV2EMODXCURLVE = 9,
       S_GV2EMODXCURLA = 10, S_CURLVE = 11, S_GAMMAT = 14, S_SPGAMMAT = 23, S_VE_MOD_AVG = 24, S_HOT_ND = 26,
       S_VE2_1 = 26, S_GVE2 = 27 };
// type tetrahedron_physics_precomp_poly4 (SRC/tetra_physics_poly_precomp_mod.f90:21-45), 544 doubles, kept in the reference's
// own order (4x4 matrices column-major: M(i,j) at [i-1 + 4*(j-1)]; anorm_in_amat*(:,n) is column n); offsets in doubles
enum { P4_AMAT = 0,        // amat1_0 amat1_1 | amat2_0..2 | amat3_0..3 | amat4_0..4 : 14 matrices of 16
       P4_AN_AMAT = 224,   // anorm_in_amat*, same 14
       P4_B0 = 448, P4_B1 = 452, P4_B2 = 456, P4_B3 = 460, P4_A10_B0 = 464, P4_A11_B0 = 480, P4_AN_B0 = 496,
       P4_AN_A10_B0 = 512, P4_AN_A11_B0 = 528, P4_ND = 544 };
enum { C_TETRA_DIST_REF = 0, C_R1 = 1, C_ER_MOD = 2, C_HPHI1 = 3, C_GHPHI = 4, C_APHI1 = 7, C_GAPHI = 8 };

// packed per-face topology: 7 bits per face f (0..3) at bit 7*f:
//   bits 0-2 neighbour_face+1 (0..5), bits 3-4 perbou_phi+1 (0..2), bits 5-6 perbou_theta+1 (0..2)
GB_HD int topo_face(uint32_t flags, int f) { return (int)((flags >> (7 * f)) & 7u) - 1; }
GB_HD int topo_perphi(uint32_t flags, int f) { return (int)((flags >> (7 * f + 3)) & 3u) - 1; }
GB_HD int topo_pertheta(uint32_t flags, int f) { return (int)((flags >> (7 * f + 5)) & 3u) - 1; }

struct MeshDev {
  int64_t ntetr;
  const double *geom;
  const double *bpart;
  const double *phi;  // nullptr when the whole Phi group is exactly zero
  const double *cold;
  const double *se;   // nullptr unless boole_strong_electric_field
  // handover_processing_kind = 2 (EXT = 2 kernels): per tetrahedron and face 48 doubles = two 192-byte blocks,
  //   leave: skew_ref_x1x2x3(3) inv_skew_coord_x1x2x3(3,3) skew_coord_xyz(3,3) skew_ref_xyz(3)
  //   enter: skew_ref_xyz(3) inv_skew_coord_xyz(3,3) skew_coord_x1x2x3(3,3) skew_ref_x1x2x3(3)      (matrices column-major)
  const double *skew;
  // bulk-copy gather: geom(16) + bpart(28) of a tetrahedron as ONE contiguous 352-byte record (a second copy of those two
  // arrays, made when the bulk gather is switched on), so that a lane needs one bulk copy per push instead of two
  const double *rec44;
  // precomputed-coefficient modes (EXT = 4 kernels): tetra_physics_poly4 records as the reference stores them, built by the
  // library at init from the tetra_physics records (make_precomp_poly4); i_precomp = 1 or 2
  const double *poly4;
  int32_t i_precomp;
  int32_t newton_precalc;   // RK pusher: normal velocity / acceleration / quadratic start guess from poly4 (EXT = 2 kernels)
  int32_t ode45, pad_ode45; // RK pusher: boole_pusher_ode45 (EXT = 2 kernels)
  double rel_err_ode45;
  const double *ham;  // hamiltonian_time records (EXT kernels only): h1_in_curlA h1_in_curlh vec_mismatch_der(3) vec_parcurr_der(3)
  int32_t prefetch;    // 1: once the exit face of a push is known, prefetch the neighbour's records into the L2 (pays when
                       // the records a batch touches do not fit the L2; costs 12-20 % when they do -- host decides)
  int32_t pad_prefetch;
  double desired_delta_energy;      // adaptive sub-stepping (EXT = 3 kernels): gorilla_settings_mod.f90:76-77
  int32_t max_n_intermediate_steps;
  int32_t time_tracing; // i_time_tracing_option: 1 = dt/dtau constant per cell, 2 = Hamiltonian time (EXT kernels)
  double cm_over_e, particle_mass, particle_charge;
  double period_phi;   // 2*pi/n_field_periods, formed exactly as the reference does (2.d0*pi/n_field_periods)
  double period_theta; // 2.d0*pi
  int32_t sign_sqg, coord_system, grid_size2, boole_guess;
  // find_tetra
  int32_t grid_kind, grid_size1, grid_size3, n_field_periods;
  double Rmin, Rmax, Zmin, Zmax, sfc_s_min;
  // find_tetra acceleration for the slice-wise grids (kinds 2, 3, 4): the tetrahedra of one phi slice binned on a uniform
  // 2-D grid over the two non-toroidal coordinates (bin_c0, bin_c1 = coordinate indices); nullptr = scan the whole slice
  const int32_t *bin_start, *bin_items;
  int32_t bin_nu, bin_nv, bin_c0, bin_c1;
  double bin_u0, bin_v0, bin_du_inv, bin_dv_inv;
};

GB_HD double ldg(const double *p)
{
#if defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}
#if defined(__CUDA_ARCH__)
// Gathers allocate in L1 on purpose: particles are sorted by tetrahedron, so neighbouring lanes and warps re-read the
// same records (measured: ld.global.nc.L1::no_allocate costs 25 % at order 2).
GB_HD void ld2(const double *p, double &a, double &b)
{
  double2 v = __ldg(reinterpret_cast<const double2 *>(p));
  a = v.x;
  b = v.y;
}
// 256-bit gather (sm_100): one 32-byte sector per lane and instruction.  The order-2 kernel saturates the L1 data pipe
// (l1tex__data_pipe_lsu_wavefronts 92 %): lanes hold different records, so every load instruction costs one wavefront
// per distinct cache line whatever its width -- twice the width, half the wavefronts.  Sub-records are 32-byte aligned.
GB_HD void ld4(const double *p, double &a, double &b, double &c, double &d)
{
  asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
#else
GB_HD void ld2(const double *p, double &a, double &b)
{
  a = p[0];
  b = p[1];
}
GB_HD void ld4(const double *p, double &a, double &b, double &c, double &d)
{
  a = p[0]; b = p[1]; c = p[2]; d = p[3];
}
#endif

// An explicit prefetch of the neighbour's record once the exit face is known was measured three times in round 1 on the
// L2-resident VMEC workload and cost 12-20 % each time.  Where the gather misses the L2 (3.8 M / 4.2 M-tetrahedron meshes) the
// push is latency bound instead (ncu: long_scoreboard 9.5 stalled warps per issue, DRAM at 9 % of peak), so the prefetch is
// a RUN-TIME option of the mesh handle (MeshDev::prefetch).
GB_HD void prefetch_l2(const void *p)
{
#if defined(__CUDA_ARCH__)
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}
// the 128-byte lines of the hot sub-records of tetrahedron t (1-based)
// staged: geom / bpart of that tetrahedron are already on their way to shared memory (GATHER kernels); only the sub-records
// that stay per-lane loads (Phi, strong E) are requested
template <int PHI>
GB_HD void prefetch_record(const MeshDev &m, int ind_tetr, bool staged = false)
{
  if (ind_tetr < 1) return;
  const int64_t t = (int64_t)ind_tetr - 1;
  if (!staged) {
    const char *pg = reinterpret_cast<const char *>(m.geom + t * GEOM_ND);
    const char *pb = reinterpret_cast<const char *>(m.bpart + t * BPART_ND);
    prefetch_l2(pg);                    // 128 bytes, line aligned
    prefetch_l2(pb);
    prefetch_l2(pb + 128);
    prefetch_l2(pb + 8 * BPART_ND - 32);  // last sector: a third line when the record straddles two boundaries
  }
  if (PHI) {
    const char *pp = reinterpret_cast<const char *>(m.phi + t * PHI_ND);
    prefetch_l2(pp);
    prefetch_l2(pp + 8 * PHI_ND - 32);
  }
  if (PHI == 2) {
    const char *ps = reinterpret_cast<const char *>(m.se + t * SE_ND);
    prefetch_l2(ps);
    prefetch_l2(ps + 128);
  }
}

// ---- bulk-copy gather (kernels launched with GATHER = 1) --------------------------------------------------------------
// ncu on the meshes whose records do not fit the L2 (3.8 M / 4.2 M tetrahedra): the gather is latency bound with only ~180
// sectors in flight per SM (long_scoreboard 9.5 stalled warps per issue, DRAM at 9 % of peak), and neither more resident warps
// nor an L2 prefetch change that -- the per-lane LDGs queue in the L1's miss path.  The bulk-copy engine (TMA,
// cp.async.bulk global -> shared, completion on an mbarrier) does not go through the L1: every lane owns a 368-byte slot of
// dynamic shared memory and an mbarrier; the geom (128 B) and bpart (224 B) sub-records of a tetrahedron are fetched with two
// bulk copies, issued for the NEIGHBOUR as soon as the exit face of the running push is known, and unpacked from shared
// memory (LDS.128, conflict free with the 16-byte pad) at the start of the next push.  (Filling the same slot with 22 per-lane
// 16-byte cp.async copies instead -- LDGSTS, through the L1 again -- was measured at a third of the bulk-copy rate.)
#define GB_BULK_STRIDE 368
#if defined(__CUDACC__)
__device__ __forceinline__ unsigned gb_tid_now()
{
  unsigned t;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t));
  return t;
}
// dynamic shared memory of a GATHER = 1 kernel of NT threads: [NT][368] slots | [NT] u64 mbarrier | [NT] i32 tetrahedron in the slot
// | [NT] i32 flags (bit 0 copy pending, bit 1 mbarrier phase parity)
__device__ __forceinline__ unsigned bulk_base()
{
  extern __shared__ __align__(16) unsigned char gb_dyn_smem[];
  return (unsigned)__cvta_generic_to_shared(gb_dyn_smem);
}
__device__ __forceinline__ unsigned bulk_slot() { return bulk_base() + gb_tid_now() * GB_BULK_STRIDE; }
__device__ __forceinline__ unsigned bulk_bar() { return bulk_base() + blockDim.x * GB_BULK_STRIDE + gb_tid_now() * 8u; }
__device__ __forceinline__ unsigned bulk_tet() { return bulk_base() + blockDim.x * (GB_BULK_STRIDE + 8u) + gb_tid_now() * 4u; }
__device__ __forceinline__ unsigned bulk_flg() { return bulk_base() + blockDim.x * (GB_BULK_STRIDE + 12u) + gb_tid_now() * 4u; }
__device__ __forceinline__ int lds_i32(unsigned a) { int v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_i32(unsigned a, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void bulk_init()
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bulk_bar()) : "memory");
  sts_i32(bulk_tet(), 0);
  sts_i32(bulk_flg(), 0);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait()
{
  const int f = lds_i32(bulk_flg());
  if (f & 1) {
    const unsigned bar = bulk_bar(), parity = (unsigned)(f >> 1) & 1u;
    asm volatile("{\n\t.reg .pred p;\n\tGB_BULK_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra GB_BULK_WAIT;\n\t}"
                 ::"r"(bar), "r"(parity) : "memory");
    sts_i32(bulk_flg(), (f ^ 2) & ~1);   // phase flips, nothing pending
  }
}
__device__ __forceinline__ void bulk_issue(const double *rec44, int ind_tetr)
{
  const unsigned bar = bulk_bar(), slot = bulk_slot();
  const double *pr = rec44 + ((int64_t)ind_tetr - 1) * 44;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the slot's last reads (generic proxy) come first
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 352;" ::"r"(bar) : "memory");
  // (the instruction takes its addresses from uniform registers: the compiler serialises the lanes of the warp around it)
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 352, [%2];"
               ::"r"(slot), "l"(pr), "r"(bar) : "memory");
  sts_i32(bulk_tet(), ind_tetr);
  sts_i32(bulk_flg(), lds_i32(bulk_flg()) | 1);
}
// make the slot hold the record of ind_tetr
__device__ __forceinline__ void bulk_acquire(const double *rec44, int ind_tetr)
{
  bulk_wait();
  if (lds_i32(bulk_tet()) != ind_tetr) {
    bulk_issue(rec44, ind_tetr);
    bulk_wait();
  }
}
__device__ __forceinline__ void lds2(unsigned a, double &x, double &y)
{
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a));
}

// ---- warp-cooperative gather (kernels launched with GATHER = 2) -----------------------------------------------------------
// The per-lane gathers cost the L1 one wavefront per lane and instruction: every lane holds a different record, so the 11
// 256-bit loads of a push are 352 wavefronts per warp (ncu: the first uses of the record carry ~30 % of all stall samples of
// the order-2 kernel although 94 % of the sectors hit the L2).  Here the WARP copies the 32 records its lanes need next:
// once the exit faces are known every lane hands the tetrahedron behind its face to the warp (shuffle), and the records are
// fetched with 16-byte cp.async copies (LDGSTS, no registers), eight consecutive lanes taking one 128-byte line of one record:
// an instruction touches 4 lines instead of 32, 24 instructions move the 32 records (stored with a stride of 384 bytes =
// three lines for this purpose, MeshDev::rec44 with GB_COOP_ND doubles per tetrahedron) straight into the lanes'
// shared-memory slots, without the lane-by-lane serialisation the bulk-copy instruction needs (its operands are
// warp-uniform).  The copies overlap the rest of the push; the lanes wait (cp.async.wait_all + __syncwarp) at the top of the
// next push and read their slot with LDS.128.  A lane whose slot does not hold the record it needs (new particle, push
// redone by the complete path, warp no longer complete at the end of the queue) falls back to the per-lane loads.
// dynamic shared memory of such a kernel of NT threads: [NT][368] slots | [NT] i32 tetrahedron in the slot
// What is staged depends on the field content of the mesh (kernel template PHI): the magnetic record alone (PHI = 0: geom +
// bpart = 22 16-byte pieces, stored with a stride of 48 doubles = three 128-byte lines) or everything a push reads: with the
// electrostatic group (PHI = 1) geom + bpart + phi = 32 pieces in records of 64 doubles = four lines, with the
// strong-electric-field terms (PHI = 2) geom + bpart + phi + the 26 hot doubles of se = 45 pieces in records of 96 doubles =
// six lines.  The PHI >= 1 kernels run two CTAs per SM (the 532- / 724-byte slots are what the shared memory holds) and keep
// the six end-of-push doubles of Rec in local memory instead of a shared-memory stash.  Slot stride: an odd number of 16-byte
// pieces, conflict free for LDS.128.
__host__ __device__ constexpr int coop_chunks(int phi)
{
  return (GEOM_ND + BPART_ND + (phi >= 1 ? PHI_ND : 0) + (phi == 2 ? S_HOT_ND : 0)) / 2;
}
__host__ __device__ constexpr int coop_nd(int phi) { return phi == 2 ? 96 : phi == 1 ? 64 : 48; }
__host__ __device__ constexpr unsigned coop_stride(int phi) { return 16u * (unsigned)(coop_chunks(phi) | 1); }
__host__ __device__ constexpr size_t coop_smem_per_thread(int phi) { return coop_stride(phi) + 4; }
template <int PHI>
__device__ __forceinline__ unsigned coop_slot() { return bulk_base() + gb_tid_now() * coop_stride(PHI); }
template <int PHI>
__device__ __forceinline__ unsigned coop_tag(unsigned t) { return bulk_base() + blockDim.x * coop_stride(PHI) + t * 4u; }
template <int PHI>
__device__ __forceinline__ void coop_init() { sts_i32(coop_tag<PHI>(gb_tid_now()), 0); }
// every copy into the slots of this warp has landed (called by all lanes of wmask together)
__device__ __forceinline__ void coop_wait(unsigned wmask)
{
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncwarp(wmask);
}
// called by all lanes of wmask together; ind_next < 1 = this lane has nothing to fetch
template <int PHI>
__device__ __forceinline__ void coop_issue(const double *rec, int ind_next, unsigned wmask)
{
  constexpr int NCH = coop_chunks(PHI), ND = coop_nd(PHI);
  constexpr unsigned STRIDE = coop_stride(PHI);
  const unsigned tid = gb_tid_now(), lane = tid & 31u;
  const bool full = (wmask == 0xffffffffu);
  const int want = ind_next >= 1 ? ind_next : 0;
  sts_i32(coop_tag<PHI>(tid), full ? want : 0);
  if (full) {
    // lanes 8 sub .. 8 sub + 7 copy the record of lane 4 g + sub (g = 0..7), lane q of them piece q of every 128-byte line
    const unsigned sub = lane >> 3, q = lane & 7u;
    const char *srcl = reinterpret_cast<const char *>(rec) - 8 * ND + q * 16u;
    const unsigned dstl = bulk_base() + ((tid & ~31u) + sub) * STRIDE + q * 16u;
#pragma unroll
    for (int g = 0; g < 8; g++) {
      // a lane with nothing to fetch gets the first record (its tag says "empty"): no branch in the copy sequence
      const unsigned t = (unsigned)max(__shfl_sync(0xffffffffu, want, 4 * g + (int)sub), 1);
      const char *src = srcl + (uint64_t)t * (8 * ND);
      const unsigned dst = dstl + (unsigned)g * 4u * STRIDE;
      // line j of the record: pieces 8 j .. 8 j + 7, the last line up to piece NCH - 1
#pragma unroll
      for (int j = 0; j < (NCH + 7) / 8; j++) {
        if (8 * j + 8 <= NCH)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 128u * j), "l"(src + 128 * j) : "memory");
        else
          asm volatile("{\n\t.reg .pred pq;\n\tsetp.lt.u32 pq, %2, %3;\n\t@pq cp.async.cg.shared.global [%0], [%1], 16;\n\t}"
                       ::"r"(dst + 128u * j), "l"(src + 128 * j), "r"(q), "n"(NCH % 8) : "memory");
      }
    }
  }
}
#endif

// One tetrahedron's hot record in registers.  PHI: 0 = magnetic part only (Phi group exactly zero), 1 = with the
// electrostatic group, 2 = electrostatic + strong-electric-field groups.
template <int PHI>
struct Rec {
  double x1[3], dist_ref, an[4][3]; // an[f][i] = anorm(i+1, f+1)
  double bmod1, gB[3], curlA[3], curlh[3], gBxh1[3], gBxcurlA, alp[9], spalp, dtdtau;
  double Phi1, gPhi[3], gPhixh1[3], gPhixcurlA, bet[9], spbet;
  double v2Emod1, gv2Emod[3], gv2Emodxh1[3], gBxcurlvE, gPhixcurlvE, gv2EmodxcurlvE, gv2EmodxcurlA, curlvE[3], gam[9], spgam,
      vE_mod_avg;
  // The first vertex (for x = z + x1 when the push ends) and the topology (hand-over) are only needed at the END of a
  // push.  Carried in registers they are spilled to local memory by the 128-register cap and every reload costs an L2
  // round trip (the L1 is thrashed by the record gathers).  load() therefore parks them in a caller-provided stash
  // -- shared memory in the kernel, st[k * sts], k = 0..5 -- and x1s()/nb()/flags() read them back from there.
  volatile double *st;
  int sts;
  int gmode;       // how this lane's geom / bpart sub-records arrive: 0 per-lane loads, 1 bulk copy into its shared-memory slot
                   // (GATHER = 1 kernels), 2 warp-cooperative cp.async into the slot (GATHER = 2 kernels)
  unsigned wmask;  // gmode 2: the lanes of this warp that are still in the push loop
  GB_HD void set_stash(volatile double *p, int stride, int gather = 0, unsigned warp_mask = 0u)
  {
    st = p; sts = stride; gmode = gather; wmask = warp_mask;
  }
  // ask for the record of the tetrahedron behind the exit face while the push is still running (ind_next < 1: nothing to
  // fetch).  gmode 2: a warp-level operation, every lane of wmask has to call it.
  GB_HD void prefetch_next(const MeshDev &m, int ind_next)
  {
#if defined(__CUDA_ARCH__)
    if (gmode == 1 && ind_next >= 1) bulk_issue(m.rec44, ind_next);
    if (gmode == 2) coop_issue<PHI>(m.rec44, ind_next, wmask);
#else
    (void)m; (void)ind_next;
#endif
  }
  GB_HD double x1s(int i) const { return st[i * sts]; }
  GB_HD static int32_t word_of(double d, int hi)
  {
#if defined(__CUDA_ARCH__)
    return hi ? __double2hiint(d) : __double2loint(d);
#else
    union { double d; int32_t i[2]; } u;
    u.d = d;
    return u.i[hi];
#endif
  }
  GB_HD int32_t nb(int f) const { return word_of(st[(3 + (f >> 1)) * sts], f & 1); }
  GB_HD uint32_t flags() const { return (uint32_t)word_of(st[5 * sts], 0); }

  GB_HD void load(const MeshDev &m, int ind_tetr /*1-based*/)
  {
    const int64_t t = (int64_t)ind_tetr - 1;
    double g[GEOM_ND], b[BPART_ND];
    const double *pg = m.geom + t * GEOM_ND, *pb = m.bpart + t * BPART_ND;
#if defined(__CUDA_ARCH__)
    bool from_slot = false;
    if (gmode == 1) {
      bulk_acquire(m.rec44, ind_tetr);
      from_slot = true;
    } else if (gmode == 2) {
      from_slot = lds_i32(coop_tag<PHI>(gb_tid_now())) == ind_tetr;   // the kernel has called coop_wait at the top of the push
    }
    const unsigned slot = gmode == 2 ? coop_slot<PHI>() : bulk_slot();
    const bool all_staged = from_slot && gmode == 2 && PHI >= 1;   // Phi (and strong-E) sub-records are in the slot as well
    if (from_slot) {
#pragma unroll
      for (int i = 0; i < GEOM_ND; i += 2) lds2(slot + 8u * i, g[i], g[i + 1]);
#pragma unroll
      for (int i = 0; i < BPART_ND; i += 2) lds2(slot + 128u + 8u * i, b[i], b[i + 1]);
    } else
#endif
    {
#pragma unroll
      for (int i = 0; i < GEOM_ND; i += 4) ld4(pg + i, g[i], g[i + 1], g[i + 2], g[i + 3]);
#pragma unroll
      for (int i = 0; i < BPART_ND; i += 4) ld4(pb + i, b[i], b[i + 1], b[i + 2], b[i + 3]);
    }
    x1[0] = g[0]; x1[1] = g[1]; x1[2] = g[2]; dist_ref = g[3];
#pragma unroll
    for (int f = 0; f < 4; f++)
#pragma unroll
      for (int i = 0; i < 3; i++) an[f][i] = g[4 + 3 * f + i];
    bmod1 = b[B_BMOD1];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      gB[i] = b[B_GB + i];
      curlA[i] = b[B_CURLA + i];
      curlh[i] = b[B_CURLH + i];
      gBxh1[i] = b[B_GBXH1 + i];
    }
    gBxcurlA = b[B_GBXCURLA];
#pragma unroll
    for (int i = 0; i < 9; i++) alp[i] = b[B_ALP + i];
    spalp = b[B_SPALP];
    dtdtau = b[B_DTDTAU];
    st[0] = g[0]; st[sts] = g[1]; st[2 * sts] = g[2];
    st[3 * sts] = b[B_TOPO]; st[4 * sts] = b[B_TOPO + 1]; st[5 * sts] = b[B_TOPO + 2];
    if (PHI) {
      double p[PHI_ND];
      const double *pp = m.phi + t * PHI_ND;
#if defined(__CUDA_ARCH__)
      if (all_staged) {
#pragma unroll
        for (int i = 0; i < PHI_ND; i += 2) lds2(slot + 8u * (GEOM_ND + BPART_ND + i), p[i], p[i + 1]);
      } else
#endif
      {
#pragma unroll
        for (int i = 0; i < PHI_ND; i += 4) ld4(pp + i, p[i], p[i + 1], p[i + 2], p[i + 3]);
      }
      Phi1 = p[P_PHI1];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        gPhi[i] = p[P_GPHI + i];
        gPhixh1[i] = p[P_GPHIXH1 + i];
      }
      gPhixcurlA = p[P_GPHIXCURLA];
#pragma unroll
      for (int i = 0; i < 9; i++) bet[i] = p[P_BET + i];
      spbet = p[P_SPBET];
    }
    if (PHI == 2) {
      double q[28];  // 26 hot doubles, fetched as seven 32-byte sectors
      const double *ps = m.se + t * SE_ND;
#if defined(__CUDA_ARCH__)
      if (all_staged) {
#pragma unroll
        for (int i = 0; i < S_HOT_ND; i += 2) lds2(slot + 8u * (GEOM_ND + BPART_ND + PHI_ND + i), q[i], q[i + 1]);
      } else
#endif
      {
#pragma unroll
        for (int i = 0; i < 28; i += 4) ld4(ps + i, q[i], q[i + 1], q[i + 2], q[i + 3]);
      }
      v2Emod1 = q[S_V2EMOD1];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        gv2Emod[i] = q[S_GV2EMOD + i];
        gv2Emodxh1[i] = q[S_GV2EMODXH1 + i];
        curlvE[i] = q[S_CURLVE + i];
      }
      gBxcurlvE = q[S_GBXCURLVE];
      gPhixcurlvE = q[S_GPHIXCURLVE];
      gv2EmodxcurlvE = q[S_GV2EMODXCURLVE];
      gv2EmodxcurlA = q[S_GV2EMODXCURLA];
#pragma unroll
      for (int i = 0; i < 9; i++) gam[i] = q[S_GAMMAT + i];
      spgam = q[S_SPGAMMAT];
      vE_mod_avg = q[S_VE_MOD_AVG];
    }
  }
};

} // namespace gb
