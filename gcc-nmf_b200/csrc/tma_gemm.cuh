// TMA-fed 3xBF16 GEMM over pre-split operand planes on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a only.
//
//   D[m, n] = sum_k A(m, k) * B(n, k)        A = A_hi + A_lo, B = B_hi + B_lo   (bf16 planes, hi = bf16(x), lo = bf16(x - hi))
//
// Each operand lives in global memory as TWO bf16 planes [2][rows][pitch] written by the kernel that produced it
// (the KL-NMF epilogues / the W update), in ONE orientation; the contraction picks the matching shared-memory layout:
//   K-major  : element (r, k) at rows r, k contiguous        TMA boxes {KB, rows / cluster extent, 1 plane}, SWIZZLE_64B rows
//   MN-major : element (r, k) at rows k, r contiguous        TMA boxes {64, KB, 2 planes} = 64-wide atoms, SWIZZLE_128B
// so no matrix is ever transposed or re-split inside the loop (round 1's kernel converted float32 operands on the fly:
// 32 KB of L2->SM traffic and 32 KB of shared-memory stores per k-block per CTA, its limiter).
// Per 16-deep k-step the products lo.hi + hi.lo + hi.hi accumulate in float32 TMEM: three tcgen05.mma.kind::f16 of width BN,
// or -- dual-N loop, K-major B with 2 BN <= 256 -- two of width 2 BN against the adjacent [B_hi; B_lo] planes of the stage,
// which also yields lo.lo; the epilogue adds the two accumulator halves.  (Measured: an M = 128 MMA costs ~128 cycles for
// any N <= 208 and ~168 at N = 224 / 256, so two wide MMAs beat three narrow ones.)
//
// CTA = one 128 x BN accumulator tile, 10 warps:
//   warp 0     lane 0 is the TMA producer: waits empty[s], arms full[s] with the stage's byte count, issues the boxes
//   warp 1     allocates TMEM; lane 0 issues every tcgen05.mma and tcgen05.commit (-> empty[s], accum_full)
//   warps 2-9  epilogue: tcgen05.ld (32 lanes x 32 columns per instruction) -> shared tile -> functor by columns (all 10 warps)
// Rows past the last full 128-row tile (F = 513 = 4 x 128 + 1) are computed in float32 SIMT by the epilogue warps while
// they wait for the accumulator.  A cluster of CN x CM CTAs (n tiles x m tiles) shares operand tiles: each CTA loads 1 / CN
// of its A tile and 1 / CM of its B tile and TMA-multicasts the slice to the CTAs of its cluster row / column.  Measured
// (DESIGN.md 4.1): that pays for the 128 x 208 tiles of the H update (1 x 2), not for the 128 x 128 tiles.
// PAIR: compile-time cta_group::2 mode (256-row MMAs over a CTA pair, B split between the two CTAs); validated on hardware,
// bit-identical to the single-CTA kernel, and measured no faster (the MMA rate, not operand staging, bounds the loop).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <type_traits>

#include "umma_ptx.cuh"

namespace tgemm {

using umma::mbar_init;
using umma::mbar_wait;
using umma::smem_u32;

constexpr int kBM = 128;
constexpr int kEpiWarps = 8;
constexpr int kThreads = (2 + kEpiWarps) * 32;
constexpr int kColumnsInFlight = 8;        // epilogue: independent column loads per thread before the first dependent store
constexpr int kMaxRowValues = 4;           // per-row values an epilogue functor may stage in shared memory
constexpr int kSmemBudget = 216 * 1024;    // stages; + barriers + epilogue scratch + alignment slack stays under the 227 KB per-CTA limit

__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// One 3-D box (inner, rows, planes) -> shared memory; completion is signalled on `bar` as transaction bytes.
// Coordinates outside the tensor are zero-filled (and still counted in the transaction bytes).
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// Same, delivered to the same shared-memory offset (and signalled on the same barrier offset) of every CTA of the cluster in `mask`.
template <bool MULTICAST>
__device__ __forceinline__ void tma_load_3d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, uint16_t mask) {
  if (MULTICAST) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
        : "memory");
  } else {
    tma_load_3d(dst, map, bar, c0, c1, c2);
  }
}
// tcgen05.commit arriving on the barrier at this offset in every CTA of `mask`.
__device__ __forceinline__ void mma_commit_multicast(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
// ---- CTA pair (cta_group::2): two CTAs of a cluster (ranks 0 / 1 = even / odd m tile) issue ONE 256 x BN MMA; each holds its
// own 128 rows of A and half of the B tile.  NOT YET RUN ON HARDWARE ("next").
// The leader's (even rank) barrier as seen from either CTA of the pair: clear the peer bit of the shared-window address.
__device__ __forceinline__ uint32_t pair_leader_addr(uint32_t smem_addr) { return smem_addr & 0xFEFFFFFFu; }
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void mma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit_pair(uint32_t bar) {     // arrives on the barrier at this offset in both CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 16 bytes from the shared memory of CTA `rank` of my cluster, at the offset of my own `local` pointer
__device__ __forceinline__ float4 ld_cluster_f32x4(const float* local, uint32_t rank) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(umma::smem_u32(local)), "r"(rank));
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(remote) : "memory");
  return v;
}
__device__ __forceinline__ void tma_prefetch_descriptor(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t) :: "memory");
  return t;
}
// Programmatic dependent launch: both are no-ops for a kernel launched without the attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait_prior_grids() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4 | LBO >> 4 at bit 16 | SBO >> 4 at bit 32 |
// version 1 at bit 46 | layout type at bit 61 (2 = SWIZZLE_128B, 4 = SWIZZLE_64B).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

// 32 lanes x 16 consecutive 32-bit columns.
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float (&v)[32]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 lanes x 8 consecutive 32-bit columns.
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, float (&v)[32]) {
  uint32_t r[8];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// barrier 1: the 8 epilogue warps only
__device__ __forceinline__ void epilogue_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory"); }

// hi = bf16(x) (round to nearest even), lo = bf16(x - hi): x = hi + lo up to 2^-17 |x|.
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// 8 consecutive elements of a plane pair (one 16-byte load per plane) as floats hi + lo.
__device__ __forceinline__ void load_planes8(const __nv_bfloat16* hi, const __nv_bfloat16* lo, float (&out)[8]) {
  const uint4 h = __ldg(reinterpret_cast<const uint4*>(hi)), l = __ldg(reinterpret_cast<const uint4*>(lo));
  const uint32_t hw[4] = {h.x, h.y, h.z, h.w}, lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    out[2 * i] = __uint_as_float(hw[i] << 16) + __uint_as_float(lw[i] << 16);
    out[2 * i + 1] = __uint_as_float(hw[i] & 0xFFFF0000u) + __uint_as_float(lw[i] & 0xFFFF0000u);
  }
}

// Completion signal of a whole launch to the ranks of a sharded run: the last CTA to finish adds 1 to the arrival counter of every
// rank (peer-mapped addresses).  The data the signal publishes lies in THIS GPU's memory -- peers fetch it over NVLink through this
// GPU's L2 -- so a device-scope fence per CTA is enough and the adds are relaxed (no MEMBAR.SYS: that costs ~5 us).
struct PeerSignal {
  unsigned* counters[8];
  int world;              // 0: no signal
};
__device__ __forceinline__ void signal_peers(const PeerSignal& sig) {
  for (int r = 0; r < sig.world; ++r) asm volatile("red.relaxed.sys.global.add.u32 [%0], %1;" ::"l"(sig.counters[r]), "r"(1u) : "memory");
}

struct PlaneGemmArgs {
  int M, N, Kc;
  int m_tiles;             // 128-row tiles on the tensor cores (= gridDim.y)
  int tail_rows;           // rows [128 m_tiles, M) computed in float32 SIMT by the epilogue warps while the main loop runs (K-major operands)
  int tail_cols;           // columns of the n tile each m tile's CTA takes for those rows (multiple of 2)
  int kblocks_per_split;   // k-blocks of KB handled by one blockIdx.z
  int preload;             // 1: the by-column epilogue's global operands are fetched into registers while the main loop runs
  int m_fastest;           // 0: grid (n tiles, m tiles, splits); 1: grid (m tiles, n tiles, splits) -- the CTAs that share a B tile are
                           // launched together, so a large B operand is read from HBM once (cluster shapes with CN == 1 only)
  int z_cluster;           // S > 1: the S k-splits of a tile form a (1, 1, S) cluster and are summed through distributed shared memory
                           // (CTA z finishes the z-th slice of the tile's columns, rank order 0 .. S - 1); the functor sees z = 0
  // SIMT tail rows (both operands K-major only): element (r, k) of plane p at ptr[p * plane + r * ld + k]
  const __nv_bfloat16* A; int64_t a_plane, lda;
  const __nv_bfloat16* B; int64_t b_plane, ldb;
  unsigned* done_counter;       // with `signal`: CTA completion count of this launch (zero before and after)
  PeerSignal signal;
  unsigned long long* timing;   // optional diagnostics, 8 slots per CTA: [0] / [7] globaltimer (ns) at CTA start / end (tail CTAs too),
                                // [1..6] clock64: start, first stage full, last MMA issued, producer done, accumulator complete, epilogue end
};

template <int BN, int KB, bool A_MN, bool B_MN, int BDIV = 1, int NACC = 1>     // BDIV = 2: CTA pair (cta_group::2), each CTA holds half of the B tile;
struct Config {                                                                  // NACC = 2: dual-N mode, two BN-column accumulators
  static_assert(KB == 32 || KB == 64, "k-block of 32 (SWIZZLE_64B K-major rows) or 64 (SWIZZLE_128B)");
  static_assert(BN % 8 == 0 && (BN * NACC) % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N (= BN, or 2 BN in the dual-N loop) for M = 128");
  static constexpr int kAAtoms = kBM / 64;
  static constexpr int kBAtoms = (BN + 63) / 64;
  static constexpr int kAtomBytes = 2 * KB * 128;                 // one MN-major atom: 64 elements x KB k-rows x 2 planes
  static constexpr int kABytes = 2 * kBM * KB * 2;                // both layouts: 512 KB
  static_assert(BDIV == 1 || (B_MN ? kBAtoms % 2 == 0 : (BN / 2) % 8 == 0), "a CTA pair splits the B tile into two halves of whole swizzle atoms");
  static constexpr int kBBytes = (B_MN ? kBAtoms * kAtomBytes : 2 * BN * KB * 2) / BDIV;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static_assert(kABytes % 1024 == 0 && kBBytes % 1024 == 0, "operand blocks must keep 1024-byte alignment");
  static constexpr int kStagesRaw = kSmemBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static_assert(kStages >= 2, "tile too large for a 2-stage pipeline");
  static constexpr int kAccCols = BN * NACC;
  static_assert(kAccCols <= 256, "accumulator columns");
  static constexpr int kTmemCols = kAccCols <= 32 ? 32 : kAccCols <= 64 ? 64 : kAccCols <= 128 ? 128 : 256;
  static constexpr int kBarrierBytes = 256;
  static constexpr int kScratchBytes = kMaxRowValues * kBM * 4 + (kThreads / 32) * 32 * 16;
  static constexpr int kTotal = kStages * kStageBytes + kBarrierBytes + kScratchBytes + 1024;   // + alignment slack
  static_assert(kStages * kStageBytes >= BN * kBM * 4, "the epilogue stages the accumulator tile in the pipeline buffers");
  static_assert(kTotal <= 227 * 1024, "shared memory per CTA");
  // K-major rows: KB bf16 = 64 B (SWIZZLE_64B, 8-row groups of 512 B) or 128 B (SWIZZLE_128B, groups of 1024 B)
  static constexpr uint32_t kKRowBytes = KB * 2;
  static constexpr uint32_t kKLayout = (KB == 32) ? 4u : 2u;
  static constexpr uint32_t kKSbo = 8 * kKRowBytes;
  // epilogue: the two warps of a TMEM lane quarter split the BN columns at a multiple of 16
  static constexpr int kCols0 = ((BN / 2 + 15) / 16) * 16;
};

// D = f32, A = B = bf16 (kind::f16), M = 128; bit 15 / 16: A / B is MN-major.
__host__ __device__ constexpr uint32_t make_idesc(int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(kBM >> 4) << 24);
}

// Epilogues with `static constexpr bool kTileEpilogue = true` take the whole staged tile:
//   __device__ void tile_epilogue(const float* tile /* [n][128] */, int m0, int n0, int n_valid, int z) const;   (all threads)
template <class E, class = void>
struct has_tile_epilogue { static constexpr bool value = false; };
template <class E>
struct has_tile_epilogue<E, decltype((void)E::kTileEpilogue)> { static constexpr bool value = E::kTileEpilogue; };

// Epilogues with `static constexpr bool kDualN = true` ask for the dual-N main loop where the shape allows it (K-major B, 2 BN <= 256,
// no CTA pair).  Measured (profiles/r02_cta_phase_stamps.md): a tcgen05.mma with M = 128 costs ~128 cycles however narrow N is
// (391 / 388 / 378 cycles per 3-MMA k-step at N = 128 / 208 / 176, and no change at all with cta_group::2, which halves the B bytes
// read per CTA) -- the instruction is bound by its 128 x 16 A slab, so it only reaches the tensor-pipe floor of N / 2 cycles at
// N = 256.  The hi and lo planes of a K-major B tile are adjacent in the stage, so ONE MMA with N = 2 BN multiplies an A plane with
// [B_hi ; B_lo] into two accumulators: 2 MMAs per k-step (A_hi, then A_lo) compute all FOUR products hi.hi + lo.hi | hi.lo + lo.lo
// instead of three products in three MMAs, and the epilogue adds the two accumulator halves (measured: 391 -> 334 cycles per k-step,
// now at the shared-memory port: 24 KB of operand reads + 16 KB of TMA writes).  With a CTA pair (cta_group::2, M = 256) on top, the
// N = 2 BN operand is split between the two CTAs by plane, which halves the B bytes each CTA reads and writes.
// Epilogues with `static constexpr bool kPreloadOperands = true` have their by-column global operands fetched into registers while
// the main loop runs (opt-in: measured to pay for the ratio epilogue of the W.H contractions, to cost for the H update).
template <class E, class = void>
struct wants_preload { static constexpr bool value = false; };
template <class E>
struct wants_preload<E, decltype((void)E::kPreloadOperands)> { static constexpr bool value = E::kPreloadOperands; };

template <class E, class = void>
struct wants_dual_n { static constexpr bool value = false; };
template <class E>
struct wants_dual_n<E, decltype((void)E::kDualN)> { static constexpr bool value = E::kDualN; };

// Epilogue concept (functors in klnmf_tma.cu).  The kernel stages the accumulator tile in shared memory and hands it out by
// columns: a warp owns column n, lane l rows m .. m + 3 with m = m0 + 4 l (contiguous in every output of the KL-NMF loop).
//   static constexpr int kRowValues (<= kMaxRowValues);  __device__ void row_values(int m, float* v) const;
//       per-row constants, fetched by one thread per row while the main loop runs and staged in shared memory
//   struct State;   __device__ void init(State&, int m, const float* rowvals) const;    rowvals[i * 128 + 0..3] = value i of rows m .. m + 3
//   struct Loaded;  __device__ Loaded load(int m, int n) const;          the column's global operands (issued kColumnsInFlight deep)
//   static constexpr bool kPrefetch;  __device__ void prefetch(int m, int n) const;     L2 prefetch of the line load(m, n) will read
//   __device__ void store(int m, int n, float4 acc, const Loaded&, int z, State&) const;
//   static constexpr bool kRowReduce;  __device__ float4 row_partial(const State&) const;  __device__ void row_total(int m, int tile_n, float) const;
//   __device__ void elem(int m, int n, float acc, int z) const;          SIMT tail rows (one column per lane)
template <int BN, int KB, bool A_MN, bool B_MN, int CN, int CM, bool PAIR, class Epilogue>
__global__ void __launch_bounds__(kThreads, 1)
plane_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, PlaneGemmArgs args, Epilogue epi) {
  constexpr bool DUAL = wants_dual_n<Epilogue>::value && !B_MN && 2 * BN <= 256;
  using C = Config<BN, KB, A_MN, B_MN, PAIR ? 2 : 1, DUAL ? 2 : 1>;
  constexpr int kCluster = CN * CM;
  static_assert(!PAIR || (CN == 1 && CM == 2), "a CTA pair is a 1 x 2 cluster (two m tiles)");
  static_assert((CN == 1 || CN == 2) && (CM == 1 || CM == 2), "cluster of CN n-tiles x CM m-tiles");
  static_assert(A_MN || (kBM / CN) % 8 == 0, "A row slices keep the swizzle atoms whole");
  static_assert(B_MN || (BN / CM) % 8 == 0, "B row slices keep the swizzle atoms whole");
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // PAIR: the two CTAs of a pair must be neighbours along x (cta_group::2 pairs are formed along the cluster's x dimension: a
  // (1, 2, 1) cluster is refused at launch with "cluster misconfiguration"), so a pair always runs on the m-fastest grid
  // (m tiles, n tiles, splits) with (2, 1, 1) clusters.
  const bool mf = PAIR || args.m_fastest != 0;
  const int tile_n = mf ? (int)blockIdx.y : (int)blockIdx.x;
  const int tile_m = mf ? (int)blockIdx.x : (int)blockIdx.y;
  const int z = blockIdx.z;
  const int n0 = tile_n * BN;
  const int total_kblocks = (args.Kc + KB - 1) / KB;
  const int kb_begin = z * args.kblocks_per_split;
  const int kb_end = min(total_kblocks, kb_begin + args.kblocks_per_split);
  const int num_kb = max(0, kb_end - kb_begin);
  // position inside the cluster (x = n tile, y = m tile); rank = x + CN y (%cluster_ctarank)
  const int cx = (!mf && CN > 1) ? (int)(blockIdx.x % CN) : 0;     // (the m-fastest grid is used with CN == 1 only)
  const int cy = (CM > 1) ? tile_m % CM : 0;
  // CTAs that receive my slice of A (same m tile: my cluster row) / of B (same n tile: my cluster column)
  const uint16_t mask_row = (uint16_t)(((1u << CN) - 1u) << (CN * cy));
  const uint16_t mask_col = (uint16_t)((CM > 1 ? ((1u << cx) | (1u << (cx + CN))) : (1u << cx)));

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* full = bars;                    // [kStages]  TMA -> MMA
  uint64_t* empty = bars + C::kStages;      // [kStages]  tcgen05.commit (of every CTA that shares a slice with me) -> TMA
  uint64_t* accum_full = bars + 2 * C::kStages;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::kStages + 1);
  // epilogue scratch behind the barriers (never touched by TMA): per-row functor values, row-sum partials of the 10 warps
  float* rowvals = reinterpret_cast<float*>(smem + C::kStages * C::kStageBytes + C::kBarrierBytes);   // [kMaxRowValues][128]
  float4* red = reinterpret_cast<float4*>(rowvals + kMaxRowValues * kBM);                             // [10 warps][32 lanes]
  float* tile = reinterpret_cast<float*>(smem);        // epilogue: [BN][128] float32, aliases the pipeline stages
  const int m0 = tile_m * kBM;

  const int cta_linear = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  if (args.timing && tid == 0) {
    args.timing[cta_linear * 8 + 0] = globaltimer_ns();
    args.timing[cta_linear * 8 + 1] = clock64();
  }
  if (tid == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(smem_u32(&full[s]), 1);
      mbar_init(smem_u32(&empty[s]), PAIR ? 1 : CN + CM - 1);     // one release per CTA whose multicast lands in this stage (incl. myself); pair: the leader's commit
    }
    mbar_init(smem_u32(accum_full), 1);
    umma::fence_barrier_init();
    tma_prefetch_descriptor(&map_a);
    tma_prefetch_descriptor(&map_b);
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_pair(smem_u32(tmem_base_slot), C::kTmemCols);
    else umma::tmem_alloc(smem_u32(tmem_base_slot), C::kTmemCols);
  }
  umma::tc_fence_before_sync();
  if (kCluster > 1) cluster_sync();     // no peer may signal my barriers or write my stages before they are initialised
  else __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_base_slot;
  // Everything above overlaps the previous kernel's tail under programmatic dependent launch; nothing below may
  // touch global memory before the prior grids have completed.
  pdl_launch_dependents();
  pdl_wait_prior_grids();

  // Early loads of the epilogue's global operands (V^T for the ratio, the old H^T for the H update): every warp fetches the values
  // of its first kPre columns into registers BEFORE it waits for the accumulator -- while the tensor pipe is still busy -- so the
  // by-column pass after the main loop starts with its operands already there (they were one or two L2 / HBM round trips on the
  // critical path: 6.2 k of a 31 k-cycle CTA for the ratio, 12 k of 29 k for the H update).
  constexpr int kWarpsAll = kThreads / 32;
  constexpr int kLoadedWords = (int)((sizeof(typename Epilogue::Loaded) + 3) / 4);
  constexpr bool kPreloads = wants_preload<Epilogue>::value && !has_tile_epilogue<Epilogue>::value && !std::is_empty<typename Epilogue::Loaded>::value;
  constexpr int kPreCap = 64 / (kLoadedWords > 0 ? kLoadedWords : 1);                      // register budget: 64 words per thread
  constexpr int kPre = kPreloads ? ((BN + kWarpsAll - 1) / kWarpsAll < kPreCap ? (BN + kWarpsAll - 1) / kWarpsAll : kPreCap) : 0;
  typename Epilogue::Loaded pre[kPre > 0 ? kPre : 1];
  auto preload = [&]() {
    if constexpr (kPre > 0) {
      if (!args.preload) return;
      const int n_valid = min(BN, args.N - n0);
#pragma unroll
      for (int u = 0; u < kPre; ++u) {
        const int cc = warp + kWarpsAll * u;
        if (cc < n_valid) pre[u] = epi.load(m0 + 4 * lane, n0 + cc);
      }
    }
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one thread)
    if (lane == 0) {
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % C::kStages;
        const uint32_t use = i / C::kStages;
        if (use > 0) mbar_wait(smem_u32(&empty[s]), (use - 1) & 1);   // the MMAs (mine and my peers') that read this stage have retired
        const uint32_t bar = smem_u32(&full[s]);
        const uint32_t a_dst = smem_u32(smem + (size_t)s * C::kStageBytes), b_dst = a_dst + C::kABytes;
        const int k0 = (kb_begin + i) * KB;
        if constexpr (PAIR) {
          // both CTAs of the pair fill their own stage and signal the LEADER's barrier, which expects the bytes of both
          const uint32_t leader_bar = pair_leader_addr(bar);
          if (cy == 0) mbar_arrive_expect_tx(bar, 2 * C::kStageBytes);
          if (A_MN) {
#pragma unroll
            for (int a = 0; a < C::kAAtoms; ++a) tma_load_3d_pair(a_dst + a * C::kAtomBytes, &map_a, leader_bar, m0 + 64 * a, k0, 0);
          } else {
#pragma unroll
            for (int p = 0; p < 2; ++p) tma_load_3d_pair(a_dst + p * (kBM * KB * 2), &map_a, leader_bar, k0, m0, p);
          }
          if (B_MN) {
            constexpr int kHalfAtoms = C::kBAtoms / 2;
#pragma unroll
            for (int a = 0; a < kHalfAtoms; ++a)
              tma_load_3d_pair(b_dst + a * C::kAtomBytes, &map_b, leader_bar, n0 + 64 * (cy * kHalfAtoms + a), k0, 0);
          } else if constexpr (DUAL) {
            // pair + dual-N: the N = 2 BN operand [B_hi ; B_lo] is split between the CTAs by PLANE -- the leader holds all BN rows of
            // the hi plane, its peer those of the lo plane (same bytes per CTA as half of both planes)
            tma_load_3d_pair(b_dst, &map_b, leader_bar, k0, n0, cy);
          } else {
            constexpr int kRows = BN / 2;
#pragma unroll
            for (int p = 0; p < 2; ++p) tma_load_3d_pair(b_dst + p * (kRows * KB * 2), &map_b, leader_bar, k0, n0 + cy * kRows, p);
          }
          continue;
        }
        mbar_arrive_expect_tx(bar, C::kStageBytes);                    // my slices + the ones my peers multicast to me
        if (A_MN) {
#pragma unroll
          for (int a = 0; a < C::kAAtoms; ++a)
            if (a % CN == cx) tma_load_3d_mc<(CN > 1)>(a_dst + a * C::kAtomBytes, &map_a, bar, m0 + 64 * a, k0, 0, mask_row);
        } else {
          constexpr int kRows = kBM / CN;      // my row slice of the A tile, one box per plane
#pragma unroll
          for (int p = 0; p < 2; ++p)
            tma_load_3d_mc<(CN > 1)>(a_dst + p * (kBM * KB * 2) + cx * (kRows * KB * 2), &map_a, bar, k0, m0 + cx * kRows, p, mask_row);
        }
        if (B_MN) {
#pragma unroll
          for (int a = 0; a < C::kBAtoms; ++a)
            if (a % CM == cy) tma_load_3d_mc<(CM > 1)>(b_dst + a * C::kAtomBytes, &map_b, bar, n0 + 64 * a, k0, 0, mask_col);
        } else {
          constexpr int kRows = BN / CM;
#pragma unroll
          for (int p = 0; p < 2; ++p)
            tma_load_3d_mc<(CM > 1)>(b_dst + p * (BN * KB * 2) + cy * (kRows * KB * 2), &map_b, bar, k0, n0 + cy * kRows, p, mask_col);
        }
      }
      if (args.timing) args.timing[cta_linear * 8 + 4] = clock64();   // producer done issuing
    }
    __syncwarp();
    preload();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0 && (!PAIR || cy == 0)) {      // pair: the leader issues for both CTAs
      constexpr uint32_t idesc = make_idesc(DUAL ? 2 * BN : BN, A_MN, B_MN) + (PAIR ? ((uint32_t)(kBM >> 4) << 24) : 0u);   // pair: M = 256
      // plane offsets inside an operand block and the per-k16 start-address advance
      constexpr uint32_t a_lo_off = A_MN ? KB * 128 : kBM * KB * 2;
      constexpr uint32_t b_lo_off = B_MN ? KB * 128 : (BN / (PAIR ? 2 : 1)) * KB * 2;
      constexpr uint32_t a_step = A_MN ? 2048u : 32u, b_step = B_MN ? 2048u : 32u;
      const uint16_t mask_release = mask_row | mask_col;
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % C::kStages;
        mbar_wait(smem_u32(&full[s]), (i / C::kStages) & 1);
        umma::tc_fence_after_sync();
        if (args.timing && i == 0) args.timing[cta_linear * 8 + 2] = clock64();
        const uint32_t a_base = smem_u32(smem + (size_t)s * C::kStageBytes), b_base = a_base + C::kABytes;
        const int k_left = args.Kc - (kb_begin + i) * KB;
        const int steps = min(KB / 16, (k_left + 15) / 16);              // the k tail issues only the 16-deep steps that hold data
#pragma unroll
        for (int kk = 0; kk < KB / 16; ++kk) {
          if (kk < steps) {
            const uint32_t a_addr = a_base + kk * a_step, b_addr = b_base + kk * b_step;
            const uint64_t a_hi = A_MN ? make_desc(a_addr, C::kAtomBytes, 1024, 2) : make_desc(a_addr, 16, C::kKSbo, C::kKLayout);
            const uint64_t a_lo = A_MN ? make_desc(a_addr + a_lo_off, C::kAtomBytes, 1024, 2) : make_desc(a_addr + a_lo_off, 16, C::kKSbo, C::kKLayout);
            const uint64_t b_hi = B_MN ? make_desc(b_addr, C::kAtomBytes, 1024, 2) : make_desc(b_addr, 16, C::kKSbo, C::kKLayout);
            const uint64_t b_lo = B_MN ? make_desc(b_addr + b_lo_off, C::kAtomBytes, 1024, 2) : make_desc(b_addr + b_lo_off, 16, C::kKSbo, C::kKLayout);
            if constexpr (PAIR && DUAL) {
              // M = 256 (both m tiles), N = 2 BN ([leader: B_hi ; peer: B_lo]): all four products of both tiles in two MMAs
              mma_bf16_pair(tmem_base, a_lo, b_hi, idesc, (i | kk) != 0);
              mma_bf16_pair(tmem_base, a_hi, b_hi, idesc, 1);
              (void)b_lo;
            } else if constexpr (PAIR) {
              mma_bf16_pair(tmem_base, a_lo, b_hi, idesc, (i | kk) != 0);
              mma_bf16_pair(tmem_base, a_hi, b_lo, idesc, 1);
              mma_bf16_pair(tmem_base, a_hi, b_hi, idesc, 1);
            } else if constexpr (DUAL) {
              // N = 2 BN: the descriptor at b_hi walks the BN rows of the hi plane and on into the lo plane behind it
              umma::mma_bf16(tmem_base, a_lo, b_hi, idesc, (i | kk) != 0);
              umma::mma_bf16(tmem_base, a_hi, b_hi, idesc, 1);
              (void)b_lo;
            } else {
              umma::mma_bf16(tmem_base, a_lo, b_hi, idesc, (i | kk) != 0);
              umma::mma_bf16(tmem_base, a_hi, b_lo, idesc, 1);
              umma::mma_bf16(tmem_base, a_hi, b_hi, idesc, 1);
            }
          }
        }
        // stage reusable once these MMAs retire: tell every CTA whose multicast lands in it
        if (PAIR) mma_commit_pair(smem_u32(&empty[s]));
        else if (kCluster > 1) mma_commit_multicast(smem_u32(&empty[s]), mask_release);
        else umma::mma_commit(smem_u32(&empty[s]));
      }
      if (num_kb > 0) {
        if (PAIR) mma_commit_pair(smem_u32(accum_full));
        else umma::mma_commit(smem_u32(accum_full));
      }
      if (args.timing) args.timing[cta_linear * 8 + 3] = clock64();
    } else if (PAIR && lane == 0 && args.timing) {
      // the non-leader CTA of a pair issues no MMA: its slots [2] / [3] record the end of its prologue (diagnostics only)
      args.timing[cta_linear * 8 + 2] = clock64();
      args.timing[cta_linear * 8 + 3] = 0;
    }
    __syncwarp();
    preload();
  } else {
    // ------------------------------------------------------------------ epilogue warps, while the main loop runs
    const int e = warp - 2;
    const int quarter = warp & 3;                        // TMEM lanes 32 (warp % 4) .. + 31 are the ones this warp may read
    const int half = e >> 2;
    if (Epilogue::kRowValues > 0 && e < 4) {             // the functor's per-row values, one row per thread
      float rv[Epilogue::kRowValues > 0 ? Epilogue::kRowValues : 1];
      epi.row_values(m0 + e * 32 + lane, rv);
#pragma unroll
      for (int i = 0; i < Epilogue::kRowValues; ++i) rowvals[i * kBM + e * 32 + lane] = rv[i];
    }
    if (Epilogue::kPrefetch) {
      // pull the epilogue's global operands of this tile towards L2 while the main loop runs (they were last touched an
      // iteration ago and have partly been evicted to HBM since): one prefetch per 128-byte line, 4 lines per column
      const int n_valid = min(BN, args.N - n0);
      if ((lane & 7) == 0)
        for (int c = e; c < n_valid; c += kEpiWarps) epi.prefetch(m0 + 4 * lane, n0 + c);
    }
    if (!A_MN && !B_MN && args.tail_rows > 0) {
      // Rows past the last full 128-row tile (F = 513 = 4 x 128 + 1), float32 SIMT from the K-major planes: the m tiles of
      // this n tile share its columns (tail_cols <= 256 each), each warp takes two columns per step, each lane 8 consecutive k per
      // 16-byte load (hi and lo plane), four k-chunks in flight.  The result of step i stays in lanes 2i / 2i + 1 and the
      // functor runs with one column per lane, so its global loads overlap.
      // (k-splits summed inside a cluster, z_cluster > 1: the functor is not linear in the accumulator, so the tail rows are computed
      // over the WHOLE contraction, each split taking its share of the tile's tail columns)
      const bool z_red = kCluster == 1 && args.z_cluster > 1;
      const int k_begin = z_red ? 0 : kb_begin * KB, k_end = z_red ? args.Kc : min(args.Kc, kb_end * KB);
      const int tcols = z_red ? (((args.tail_cols + args.z_cluster - 1) / args.z_cluster) + 1) & ~1 : args.tail_cols;
      const int c_begin = n0 + (z_red ? tile_m * args.z_cluster + z : tile_m) * tcols;
      const int c_end = min(min(args.N, n0 + BN), c_begin + tcols);
      for (int m = args.m_tiles * kBM; m < args.M; ++m) {
        const __nv_bfloat16* a_hi = args.A + (int64_t)m * args.lda;
        const __nv_bfloat16* a_lo = a_hi + args.a_plane;
        float keep = 0.f;
#pragma unroll 1
        for (int i = 0; i < 16; ++i) {
          const int n = c_begin + 2 * (e + i * kEpiWarps);
          if (n >= c_end) break;
          const bool two = n + 1 < c_end;
          const __nv_bfloat16* b_hi = args.B + (int64_t)n * args.ldb;
          const __nv_bfloat16* b_lo = b_hi + args.b_plane;
          const int64_t next = two ? args.ldb : 0;
          float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 4
          for (int k = k_begin + 8 * lane; k < k_end; k += 256) {     // pitches are multiples of 8; pad columns hold zeros
            float a[8], b0[8], b1[8];
            load_planes8(a_hi + k, a_lo + k, a);
            load_planes8(b_hi + k, b_lo + k, b0);
            load_planes8(b_hi + next + k, b_lo + next + k, b1);
#pragma unroll
            for (int j = 0; j < 8; ++j) { acc0 = fmaf(a[j], b0[j], acc0); acc1 = fmaf(a[j], b1[j], acc1); }
          }
          for (int o = 16; o > 0; o >>= 1) { acc0 += __shfl_xor_sync(0xffffffffu, acc0, o); acc1 += __shfl_xor_sync(0xffffffffu, acc1, o); }
          if (lane == 2 * i) keep = acc0;
          if (lane == 2 * i + 1) keep = acc1;
        }
        const int n_mine = c_begin + 2 * (e + (lane >> 1) * kEpiWarps) + (lane & 1);
        if (n_mine < c_end) epi.elem(m, n_mine, keep, z_red ? 0 : z);
      }
    }
    preload();
    // ------------------------------------------------------------------ epilogue, phase 1 (8 warps)
    // TMEM -> registers -> shared tile[n][m] (the pipeline stages are idle by then: every TMA box has landed -- mine and the
    // ones my peers multicast to me were all consumed by my MMAs -- and every MMA has retired).
    if (num_kb > 0) {
      mbar_wait(smem_u32(accum_full), 0);   // every MMA has retired: accumulator complete
      umma::tc_fence_after_sync();
      if (args.timing && tid == 64) args.timing[cta_linear * 8 + 5] = clock64();
    }
    const int col0 = half ? C::kCols0 : 0;
    const int ncols = half ? BN - C::kCols0 : C::kCols0;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)col0;
    float* dst = tile + (size_t)col0 * kBM + quarter * 32 + lane;
    float v[32];
    int c = 0;
#pragma unroll 1
    for (; c + 32 <= ncols; c += 32) {
      if (num_kb > 0) {
        umma::tmem_ld_32x32(taddr + (uint32_t)c, v);
        if constexpr (DUAL) {      // + the [hi ; lo] . B_lo half: columns BN .. 2 BN - 1
          float v2[32];
          umma::tmem_ld_32x32(taddr + (uint32_t)(BN + c), v2);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += v2[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) dst[(size_t)(c + j) * kBM] = v[j];
    }
    if (c + 16 <= ncols) {   // 16-column remainder (BN = 176, 208)
      if (num_kb > 0) {
        tmem_ld_32x16(taddr + (uint32_t)c, v);
        if constexpr (DUAL) {
          float v2[32];
          tmem_ld_32x16(taddr + (uint32_t)(BN + c), v2);
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] += v2[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) dst[(size_t)(c + j) * kBM] = v[j];
      c += 16;
    }
    if (c < ncols) {   // 8-column remainder (BN = 104: halves of 64 and 40 columns)
      if (num_kb > 0) {
        tmem_ld_32x8(taddr + (uint32_t)c, v);
        if constexpr (DUAL) {
          float v2[32];
          tmem_ld_32x8(taddr + (uint32_t)(BN + c), v2);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] += v2[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) dst[(size_t)(c + j) * kBM] = v[j];
    }
  }
  // ------------------------------------------------------------------ epilogue, phase 2 (all 10 warps)
  // Each warp owns whole columns n -- 128 consecutive m are contiguous in every output -- one float4 of m per lane, so each
  // global access of a warp is one 512-byte (float32) or 256-byte (bf16 plane) row segment, with kColumnsInFlight
  // independent loads per thread before the first dependent store.
  umma::tc_fence_before_sync();
  __syncthreads();
  const int zc = (kCluster == 1 && !has_tile_epilogue<Epilogue>::value) ? args.z_cluster : 0;
  if (zc > 1) cluster_sync();             // every k-split of this tile has staged its partial accumulator
  if constexpr (has_tile_epilogue<Epilogue>::value) {
    // whole-tile epilogue: the functor reads the staged accumulator tile[n][m] itself (reductions ACROSS columns, e.g. the argmax
    // over the TDOAs of a frame, which the by-column hand-out below cannot express)
    epi.tile_epilogue(tile, m0, n0, min(BN, args.N - n0), z);
    if (args.timing && tid == 64) args.timing[cta_linear * 8 + 6] = clock64();
  } else {
    constexpr int kWarps = kThreads / 32;
    const int m_first = m0 + 4 * lane;
    typename Epilogue::State st;
    epi.init(st, m_first, rowvals + 4 * lane);
    int n_valid = min(BN, args.N - n0);
    constexpr int U = kColumnsInFlight;
    int c_first = warp;
    if (zc > 1) {                          // my slice of the tile's columns
      const int per = (n_valid + zc - 1) / zc;
      c_first = z * per + warp;
      n_valid = min(n_valid, (z + 1) * per);
    }
    if constexpr (kPre > 0) {
      if (args.preload && zc <= 1) {
#pragma unroll
        for (int u = 0; u < kPre; ++u) {
          const int cc = warp + kWarps * u;
          if (cc < n_valid) {
            const float4 acc = *reinterpret_cast<const float4*>(tile + (size_t)cc * kBM + 4 * lane);
            epi.store(m_first, n0 + cc, acc, pre[u], z, st);
          }
        }
        c_first = warp + kWarps * kPre;
      }
    }
#pragma unroll 1
    for (int c = c_first; c < n_valid; c += kWarps * U) {
      typename Epilogue::Loaded loaded[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int cc = c + kWarps * u;
        if (cc < n_valid) loaded[u] = epi.load(m_first, n0 + cc);
      }
      if (zc > 1) {
        // sum of the k-splits in split order (the order the W update used for the slabs): the distributed-shared-memory loads of ALL
        // U columns of a split are issued before the first add (one at a time they cost a remote-SM round trip per column: the
        // epilogue took 15 k cycles instead of 6 k)
        float4 sum[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int cc = c + kWarps * u;
          const float* src = tile + (size_t)(cc < n_valid ? cc : 0) * kBM + 4 * lane;
          sum[u] = z == 0 ? *reinterpret_cast<const float4*>(src) : ld_cluster_f32x4(src, 0);
        }
        for (int r = 1; r < zc; ++r) {
          float4 t[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int cc = c + kWarps * u;
            const float* src = tile + (size_t)(cc < n_valid ? cc : 0) * kBM + 4 * lane;
            t[u] = r == z ? *reinterpret_cast<const float4*>(src) : ld_cluster_f32x4(src, (uint32_t)r);
          }
#pragma unroll
          for (int u = 0; u < U; ++u) { sum[u].x += t[u].x; sum[u].y += t[u].y; sum[u].z += t[u].z; sum[u].w += t[u].w; }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int cc = c + kWarps * u;
          if (cc < n_valid) epi.store(m_first, n0 + cc, sum[u], loaded[u], 0, st);
        }
      } else {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int cc = c + kWarps * u;
          if (cc < n_valid) {
            const float4 acc = *reinterpret_cast<const float4*>(tile + (size_t)cc * kBM + 4 * lane);
            epi.store(m_first, n0 + cc, acc, loaded[u], z, st);
          }
        }
      }
    }
    if constexpr (Epilogue::kRowReduce) {   // per-row sums over the tile's columns: 10 warp partials -> one value per row
      red[warp * 32 + lane] = epi.row_partial(st);
      __syncthreads();
      if (tid < kBM) {
        const float* r = reinterpret_cast<const float*>(red);
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) sum += r[w * kBM + tid];
        epi.row_total(m0 + tid, tile_n, sum);
      }
    }
    if (args.timing && tid == 64) args.timing[cta_linear * 8 + 6] = clock64();
  }
  // No CTA of a cluster may exit while a peer can still signal its barriers (my last releases have been delivered by now:
  // they precede accum_full, which the epilogue waited for).
  if (kCluster > 1 || zc > 1) cluster_sync();          // (z-cluster: no CTA exits while a peer still reads its tile)
  else __syncthreads();
  if (args.signal.world > 0 && tid == 0) {             // every store of this CTA precedes the barrier above
    __threadfence();
    const unsigned prev = atomicAdd(args.done_counter, 1u);
    if (prev == gridDim.x * gridDim.y * gridDim.z - 1) {
      *args.done_counter = 0;
      __threadfence();
      signal_peers(args.signal);
    }
  }
  if (args.timing && tid == 0) args.timing[cta_linear * 8 + 7] = globaltimer_ns();
  if (warp == 1) {
    umma::tc_fence_after_sync();
    if (PAIR) tmem_dealloc_pair(tmem_base, C::kTmemCols);
    else umma::tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

}  // namespace tgemm
