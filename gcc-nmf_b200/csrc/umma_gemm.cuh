// 3xTF32 error-compensated GEMM on the 5th-generation tensor cores (tcgen05 / TMEM), sm_100a only.
//
//   D[m, n] = sum_k A[m, k] * B[n, k]        A (M x Kc), B (N x Kc): float32, row-major, k contiguous
//
// KL-NMF needs float32-level accuracy over 100 multiplicative iterations (plain TF32 drifts to 1e-3,
// SURVEY.md section 7), so every float32 operand x is split into hi = tf32(x) and lo = tf32(x - hi)
// and each k-step issues three tcgen05.mma.kind::tf32 instructions into the same float32 TMEM
// accumulator:  lo.hi + hi.lo + hi.hi  (the lo.lo term is below float32 resolution).
//
// CTA = one 128 x BN output tile.  Warp roles:
//   warps 0 .. LW-1  loaders: coalesced 16-byte global loads (register double-buffered), split into
//                    hi / lo, stored to shared memory in the canonical K-major SWIZZLE_128B UMMA layout
//                    (rows of 32 floats = 128 B, 8-row groups of 1024 B, 16-byte chunk index XOR row%8);
//                    after the main loop the same warps are the epilogue (tcgen05.ld -> functor).
//   warp LW          allocates TMEM, and its lane 0 issues every tcgen05.mma / tcgen05.commit.
// Pipelines: shared-memory stages guarded by full[] (loaders -> MMA, 1 arrival per loader thread after
// fence.proxy.async) and empty[] (tcgen05.commit -> loaders); accum_full (tcgen05.commit -> epilogue).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace umma {

constexpr int kBM = 128;      // UMMA M (rows of the accumulator = TMEM lanes)
constexpr int kBK = 32;       // floats per k-block = one 128-byte swizzle row
constexpr int kUmmaK = 8;     // K of one kind::tf32 instruction (32 bytes)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must trap (sticky CUDA error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins) {
    if (spins > (1u << 24)) __trap();
  }
}

__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, issued by ONE thread.
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier when all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp receives row (lane base + i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory matrix descriptor, K-major, SWIZZLE_128B (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (unused for swizzled K-major: 1)
//   [32,46) stride byte offset >> 4 (1024 B between 8-row groups) | [46,48) version = 1 | [61,64) layout = 2
__device__ __forceinline__ uint64_t make_desc_kmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// K-major SWIZZLE_64B (rows of 64 bytes = 32 bf16, 8-row groups of 512 B, 16-byte chunk index XOR (row / 2) % 4): layout = 4.
__device__ __forceinline__ uint64_t make_desc_kmajor_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
// D = f32, A = B = bf16 (kind::f16, format 1), both K-major, M x N; K = 16 per instruction.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// two floats -> packed bf16x2 (round to nearest even); low half = first argument
__device__ __forceinline__ uint32_t pack_bf16x2(float lo_elem, float hi_elem) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): D = f32, A = B = tf32, both K-major, M x N.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Round-to-nearest (ties away) to TF32 with two integer ops: add half an ulp of the 10-bit mantissa,
// clear the 13 low bits.  Same result as cvt.rna.tf32.f32 for finite inputs that do not round up to
// infinity (the cvt compiles to a ~7-instruction sequence because it also handles those).
__device__ __forceinline__ float tf32_round(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}

struct GemmArgs {
  const float* A;   // (M, Kc) row-major, leading dimension lda (multiple of 4 floats, 16-byte aligned rows)
  const float* B;   // (N, Kc) row-major, leading dimension ldb
  int M, N, Kc;     // Kc may be any value; columns in [Kc, round_up(Kc, 4)) must hold zeros
  int64_t lda, ldb;
  int kblocks_per_split;  // k-blocks of 32 handled by one blockIdx.z
  int m_tiles;            // 128-row tiles on the tensor cores; blockIdx.y == m_tiles -> SIMT tail rows [128 m_tiles, M)
  const float* b_scale;   // optional per-k scale of B (length >= round_up(Kc, 4)): B[n][k] * b_scale[k], or NULL
  unsigned long long* timing;   // optional (diagnostics): 6 clock64 stamps per tensor-core CTA, or NULL
  int m_fastest;          // 0: grid (n tiles, m tiles + tail, splits); 1: grid (m tiles, n tiles, splits) -- CTAs sharing a B slab are co-scheduled
};

// Operand split modes (both: three tensor-core products hi.hi + hi.lo + lo.hi into a float32 accumulator):
//   kSplitTF32  hi = tf32(x), lo = x - hi            kind::tf32, K = 8,  128-byte shared rows   error ~2^-21 per product
//   kSplitBF16  hi = bf16(x), lo = bf16(x - hi)      kind::f16,  K = 16, 64-byte shared rows    error ~2^-17 per product
// BF16 halves the shared-memory bytes per k-block (the measured limiter of the TF32 main loop) and runs at twice the
// tensor rate; whether 2^-17 is enough is a parity question answered by the tests (KL-NMF, 100 iterations, 1e-4 bar).
constexpr int kSplitTF32 = 0, kSplitBF16 = 1;
constexpr int kLoaderWarps = 8;                 // default worker-warp count (loaders + epilogue); the KL-NMF GEMMs use 16
constexpr int kThreads = kLoaderWarps * 32 + 32;

template <int BN, int SPLIT = kSplitTF32>
struct GemmSmem {
  static constexpr int kElemBytes = (SPLIT == kSplitBF16) ? 2 : 4;
  static constexpr int kRowBytes = kBK * kElemBytes;              // 128 (SWIZZLE_128B) or 64 (SWIZZLE_64B)
  static constexpr int kStageBytes = (kBM + BN) * kRowBytes * 2;  // A and B tiles, hi and lo
  static constexpr int kStages = (SPLIT == kSplitBF16) ? ((BN <= 128) ? 4 : 3) : ((BN <= 128) ? 3 : 2);
  static constexpr int kBarrierBytes = 256;
  static constexpr int kTotal = kStages * kStageBytes + kBarrierBytes + 1024;  // + alignment slack
};

// One thread's share of a [ROWS x 32] float tile loaded by NT threads: chunk column c = tid % 8 (4 floats),
// rows tid / 8 + (NT / 8) i.
// Rows past the end of the matrix are clamped to its last row (their products land in accumulator rows /
// columns the epilogue never stores), so the only predicate left is the k tail, uniform per thread.
template <int ROWS, int NT, int SPLIT = kSplitTF32>
struct TileLoader {
  static constexpr int kChunks = ROWS * 8 / NT;
  static constexpr int kRowStep = NT / 8;                 // rows between consecutive chunks of one thread
  static constexpr int kRowBytes = (SPLIT == kSplitBF16) ? 64 : 128;
  static constexpr int kSmemStep = kRowStep * kRowBytes;   // bytes (kRowStep is a multiple of the 8-row swizzle groups)
  const float* ptr;            // chunk 0 of this thread at the current k-block; advanced by 32 floats per fetch
  int64_t stride;              // floats between consecutive chunks (kRowStep rows)
  int imax;                    // last chunk index whose row is inside the matrix (chunks past it re-read that row)
  int kcol;                    // c * 4
  uint32_t smem_off;           // swizzled byte offset of chunk 0 inside the tile; chunk i is + kSmemStep i

  __device__ __forceinline__ void init(const float* src, int64_t ld, int row0, int rows_valid, int k0, int tid) {
    const int r = tid >> 3, c = tid & 7;
    kcol = c * 4;
    const int first = min(row0 + r, rows_valid - 1);
    ptr = src + (int64_t)first * ld + k0 + kcol;
    stride = (int64_t)kRowStep * ld;
    imax = max(0, (rows_valid - 1 - first) / kRowStep);
    if (SPLIT == kSplitBF16)   // 4 floats -> 8 bytes: half of 16-byte chunk c/2, swizzled with (row / 2) % 4
      smem_off = (uint32_t)(r >> 3) * 512u + (uint32_t)(r & 7) * 64u + (uint32_t)(((c >> 1) ^ ((r >> 1) & 3)) << 4) + (uint32_t)(c & 1) * 8u;
    else
      smem_off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u + (uint32_t)((c ^ (r & 7)) << 4);
  }
  // loads the k-block starting at column k0 (the pointer already points there) and advances to the next one
  __device__ __forceinline__ void fetch(float4 (&regs)[kChunks], int k0, int kc4) {
    if (k0 + kcol < kc4) {
#pragma unroll
      for (int i = 0; i < kChunks; ++i) regs[i] = __ldg(reinterpret_cast<const float4*>(ptr + (int64_t)min(i, imax) * stride));
    } else {
#pragma unroll
      for (int i = 0; i < kChunks; ++i) regs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    ptr += kBK;
  }
  __device__ __forceinline__ void skip(int kblocks) { ptr += kblocks * kBK; }
  // hi = tf32(x) (round to nearest), lo = x - hi exactly (|lo| <= 2^-11 |x|; the tensor core drops its bits
  // below tf32 resolution: <= 2^-21 |x|, sign-symmetric)
  __device__ __forceinline__ void stash(const float4 (&regs)[kChunks], unsigned char* hi, unsigned char* lo) const {
#pragma unroll
    for (int i = 0; i < kChunks; ++i) {
      const float4 x = regs[i];
      if (SPLIT == kSplitBF16) {
        const uint32_t h01 = pack_bf16x2(x.x, x.y), h23 = pack_bf16x2(x.z, x.w);
        const float l0 = x.x - __uint_as_float(h01 << 16), l1 = x.y - __uint_as_float(h01 & 0xFFFF0000u);
        const float l2 = x.z - __uint_as_float(h23 << 16), l3 = x.w - __uint_as_float(h23 & 0xFFFF0000u);
        *reinterpret_cast<uint2*>(hi + smem_off + i * kSmemStep) = make_uint2(h01, h23);
        *reinterpret_cast<uint2*>(lo + smem_off + i * kSmemStep) = make_uint2(pack_bf16x2(l0, l1), pack_bf16x2(l2, l3));
        continue;
      }
      float4 h, l;
      h.x = tf32_round(x.x); h.y = tf32_round(x.y); h.z = tf32_round(x.z); h.w = tf32_round(x.w);
      l.x = x.x - h.x; l.y = x.y - h.y; l.z = x.z - h.z; l.w = x.w - h.w;
      *reinterpret_cast<float4*>(hi + smem_off + i * kSmemStep) = h;
      *reinterpret_cast<float4*>(lo + smem_off + i * kSmemStep) = l;
    }
  }
};

// Boilerplate for epilogues without cross-chunk state (kept as members so the functors stay aggregates).
#define UMMA_EPILOGUE_STATELESS \
  struct State {};            \
  __device__ void init(State&) const {}

// 32 x 32 register transpose through a per-warp shared scratch (33-float rows): in: v[j] = D[row lane][col j];
// out: v[i] = D[row i][col lane].
__device__ __forceinline__ void warp_transpose_32x32(float (&v)[32], float* scratch, int lane) {
#pragma unroll
  for (int j = 0; j < 32; ++j) scratch[lane * 33 + j] = v[j];
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = scratch[i * 33 + lane];
  __syncwarp();
}

// Epilogue concept (see klnmf_tc.cu for the functors):
//   struct State;  __device__ void init(State&) const;
//   __device__ void tile(int m_base, int lane, int n0, float (&v)[32], int z, int slot, float* scratch, State&) const;
//       called by each epilogue warp per 32-column chunk, in increasing n0 over the warp's BN/2 columns; the
//       thread holds row m_base + lane, columns n0 .. n0 + 31; slot = which column half of which tile this
//       warp owns (for per-CTA partial outputs); scratch = 32 x 33 floats of shared memory private to the
//       warp (for warp_transpose_32x32); State = per-thread registers carried across the chunks.
//   __device__ void elem(int m, int n, float acc, int z) const;      // SIMT tail rows
template <int BN, bool SCALE_B, int SPLIT, int LW, class Epilogue>
__global__ void __launch_bounds__(LW * 32 + 32, 1)
gemm_tn_3xtf32_kernel(GemmArgs args, Epilogue epi) {
  using S = GemmSmem<BN, SPLIT>;
  constexpr int kLoaderWarps = LW, kLoaderThreads = LW * 32, kThreads = LW * 32 + 32;
  static_assert(LW == 8 || LW == 16, "worker warps");
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tile_m = args.m_fastest ? blockIdx.x : blockIdx.y;
  const int tile_n = args.m_fastest ? blockIdx.y : blockIdx.x;
  const int n0 = tile_n * BN;
  const int kc4 = (args.Kc + 3) & ~3;
  const int total_kblocks = (args.Kc + kBK - 1) / kBK;
  const int kb_begin = blockIdx.z * args.kblocks_per_split;
  const int kb_end = min(total_kblocks, kb_begin + args.kblocks_per_split);
  const int num_kb = max(0, kb_end - kb_begin);

  if (tile_m >= args.m_tiles) {
    // ---------------------------------------------------------------- SIMT tail rows (runs on SMs the tile grid leaves idle)
    const int k_begin = kb_begin * kBK, k_end = min(kc4, kb_end * kBK);
    for (int m = args.m_tiles * kBM; m < args.M; ++m) {
      const float4* a = reinterpret_cast<const float4*>(args.A + (int64_t)m * args.lda);
      for (int n = n0 + warp; n < min(args.N, n0 + BN); n += kThreads / 32) {
        const float4* b = reinterpret_cast<const float4*>(args.B + (int64_t)n * args.ldb);
        float acc = 0.f;
        for (int k4 = k_begin / 4 + lane; k4 < k_end / 4; k4 += 32) {
          const float4 x = __ldg(a + k4);
          float4 y = __ldg(b + k4);
          if (SCALE_B) {
            const float4 sc = __ldg(reinterpret_cast<const float4*>(args.b_scale) + k4);
            y.x *= sc.x; y.y *= sc.y; y.z *= sc.z; y.w *= sc.w;
          }
          acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc); acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
        }
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) epi.elem(m, n, acc, (int)blockIdx.z);
      }
    }
    return;
  }

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kStages * S::kStageBytes);
  uint64_t* full = bars;                    // [kStages]
  uint64_t* empty = bars + S::kStages;      // [kStages]
  uint64_t* accum_full = bars + 2 * S::kStages;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 2 * S::kStages + 1);
  const int m0 = tile_m * kBM;

  const int cta_linear = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  if (args.timing && tid == 0) args.timing[cta_linear * 6 + 0] = clock64();
  if (tid == 0) {
    for (int s = 0; s < S::kStages; ++s) {
      mbar_init(smem_u32(&full[s]), (BN <= 128) ? 128 : 256);   // one loader group fills a stage
      mbar_init(smem_u32(&empty[s]), 1);
    }
    mbar_init(smem_u32(accum_full), 1);
    fence_barrier_init();
  }
  if (warp == kLoaderWarps) tmem_alloc(smem_u32(tmem_base_slot), BN);   // BN float32 accumulator columns (power of two >= 32)
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_base_slot;

  auto stage_ptr = [&](int s) { return smem + (size_t)s * S::kStageBytes; };
  // stage layout: A_hi | A_lo | B_hi | B_lo
  constexpr int kATile = kBM * S::kRowBytes, kBTile = BN * S::kRowBytes;

  if (warp < kLoaderWarps) {
    // ------------------------------------------------------------------ loaders
    // The 8 loader warps form kGroups groups of 128 threads; group g produces k-blocks g, g + kGroups, ...  Each
    // thread keeps ONE register set: the loads of its next k-block are issued right after it has stored the
    // current one, and while it waits for them the other group's warps (same SM sub-partitions) work -- the
    // overlap comes from warp scheduling.  (A two-register-set software pipeline inside one thread does not
    // overlap: ptxas tracks all the LDGs of both sets on one scoreboard slot, so waiting for the older set waits
    // for the newer one as well -- seen in the SASS control codes, and as 38 % long-scoreboard stalls in ncu.)
    // BN = 256: one group (a k-block there is 1536 MMA cycles, so a prefetch distance of one k-block is enough and the
    // 12 chunks per thread fit the register budget; two groups would need 24).
    constexpr int kGroups = (BN <= 128) ? LW / 4 : LW / 8;      // groups of 128 (BN = 128) / 256 (BN = 256) threads
    constexpr int kGroupThreads = kLoaderThreads / kGroups;
    const int group = tid / kGroupThreads, gtid = tid % kGroupThreads;
    TileLoader<kBM, kGroupThreads, SPLIT> la;
    TileLoader<BN, kGroupThreads, SPLIT> lb;
    const int kb_first = kb_begin + group;
    la.init(args.A, args.lda, m0, min(args.M, args.m_tiles * kBM), kb_first * kBK, gtid);
    lb.init(args.B, args.ldb, n0, args.N, kb_first * kBK, gtid);
    float4 ar[TileLoader<kBM, kGroupThreads, SPLIT>::kChunks], br[TileLoader<BN, kGroupThreads, SPLIT>::kChunks];
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f);
    auto fetch = [&](int kb) {
      la.fetch(ar, kb * kBK, kc4);
      lb.fetch(br, kb * kBK, kc4);
      if (SCALE_B) sc = (kb * kBK + lb.kcol < kc4) ? __ldg(reinterpret_cast<const float4*>(args.b_scale + kb * kBK + lb.kcol))
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
      la.skip(kGroups - 1);            // the pointers advance by one k-block per fetch; jump over the other groups' k-blocks
      lb.skip(kGroups - 1);
    };
    if (group < num_kb) fetch(kb_first);
    for (int i = group; i < num_kb; i += kGroups) {
      const int s = i % S::kStages;
      const uint32_t use = i / S::kStages;
      if (SCALE_B) {
#pragma unroll
        for (int j = 0; j < TileLoader<BN, kGroupThreads, SPLIT>::kChunks; ++j) { br[j].x *= sc.x; br[j].y *= sc.y; br[j].z *= sc.z; br[j].w *= sc.w; }
      }
      if (use > 0) mbar_wait(smem_u32(&empty[s]), (use - 1) & 1);   // MMAs that read this stage have retired
      unsigned char* st = stage_ptr(s);
      la.stash(ar, st, st + kATile);
      lb.stash(br, st + 2 * kATile, st + 2 * kATile + kBTile);
      if (i + kGroups < num_kb) fetch(kb_begin + i + kGroups);
      fence_proxy_async_smem();          // generic-proxy stores -> visible to the tensor core (async proxy)
      mbar_arrive(smem_u32(&full[s]));
    }
    // ------------------------------------------------------------------ epilogue
    if (num_kb > 0) {
      if (args.timing && tid == 0) args.timing[cta_linear * 6 + 3] = clock64();   // loaders done producing
      mbar_wait(smem_u32(accum_full), 0);   // every MMA has retired: accumulator complete, pipeline stages free
      tc_fence_after_sync();
      if (args.timing && tid == 0) args.timing[cta_linear * 6 + 4] = clock64();
    }
    const int quarter = warp & 3;                         // TMEM lane quarter this warp may read
    constexpr int kColsPerWarp = BN / (LW / 4);           // the LW / 4 warps of a lane quarter split the columns
    const int slot = warp >> 2;
    const int col0 = slot * kColsPerWarp;
    float* scratch = reinterpret_cast<float*>(smem) + warp * (32 * 33);   // aliases stage 0 (idle now)
    typename Epilogue::State epi_state;
    epi.init(epi_state);
#pragma unroll 1
    for (int c = 0; c < kColsPerWarp; c += 32) {
      float v[32];
      if (num_kb > 0) {
        tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(col0 + c), v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      epi.tile(m0 + quarter * 32, lane, n0 + col0 + c, v, (int)blockIdx.z, tile_n * (LW / 4) + slot, scratch, epi_state);
    }
    if (args.timing && tid == 0) args.timing[cta_linear * 6 + 5] = clock64();
    tc_fence_before_sync();
  } else {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      constexpr uint32_t idesc = (SPLIT == kSplitBF16) ? make_idesc_bf16(kBM, BN) : make_idesc_tf32(kBM, BN);
      for (int i = 0; i < num_kb; ++i) {
        const int s = i % S::kStages;
        mbar_wait(smem_u32(&full[s]), (i / S::kStages) & 1);
        tc_fence_after_sync();
        if (args.timing && i == 0) args.timing[cta_linear * 6 + 1] = clock64();
        const uint32_t base = smem_u32(stage_ptr(s));
        if (SPLIT == kSplitBF16) {
          const uint64_t a_hi = make_desc_kmajor_sw64(base), a_lo = make_desc_kmajor_sw64(base + kATile);
          const uint64_t b_hi = make_desc_kmajor_sw64(base + 2 * kATile), b_lo = make_desc_kmajor_sw64(base + 2 * kATile + kBTile);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t adv = (uint64_t)((k * 16 * 2) >> 4);       // +32 bytes per 16-element k-step inside the 64-byte swizzle row
            mma_bf16(tmem_base, a_lo + adv, b_hi + adv, idesc, (i | k) != 0);
            mma_bf16(tmem_base, a_hi + adv, b_lo + adv, idesc, 1);
            mma_bf16(tmem_base, a_hi + adv, b_hi + adv, idesc, 1);
          }
        } else {
          const uint64_t a_hi = make_desc_kmajor_sw128(base), a_lo = make_desc_kmajor_sw128(base + kATile);
          const uint64_t b_hi = make_desc_kmajor_sw128(base + 2 * kATile), b_lo = make_desc_kmajor_sw128(base + 2 * kATile + kBTile);
#pragma unroll
          for (int k = 0; k < kBK / kUmmaK; ++k) {
            const uint64_t adv = (uint64_t)((k * kUmmaK * 4) >> 4);   // +32 bytes per k-step inside the 128-byte swizzle row
            mma_tf32(tmem_base, a_lo + adv, b_hi + adv, idesc, (i | k) != 0);
            mma_tf32(tmem_base, a_hi + adv, b_lo + adv, idesc, 1);
            mma_tf32(tmem_base, a_hi + adv, b_hi + adv, idesc, 1);
          }
        }
        mma_commit(smem_u32(&empty[s]));      // stage reusable once these MMAs retire
      }
      if (num_kb > 0) mma_commit(smem_u32(accum_full));
      if (args.timing) args.timing[cta_linear * 6 + 2] = clock64();
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == kLoaderWarps) {
    tc_fence_after_sync();
    tmem_dealloc(tmem_base, BN);
  }
}

}  // namespace umma
