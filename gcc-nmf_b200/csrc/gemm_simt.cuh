// Generic shared-memory tiled SIMT GEMM  D[m, n] = sum_k A(m, k) * B(n, k)  with functor operand
// loaders (so operands can be generated on the fly: masked H, Re(coherence * E) ...) and a functor
// epilogue (so V/(W.H), the multiplicative updates, argmax ... are fused and the product matrix
// never round-trips HBM).  Templated on the accumulator type: float for the NMF contractions that
// the reference does in float32, double for the contractions the reference does in float64.
//
// Tile BM x BN x BK, (BM/TM) x (BN/TN) threads, each thread owns a TM x TN register tile split
// into two halves per dimension (rows ty*TM/2 + {0..} and BM/2 + ...) so that shared-memory reads
// are contiguous 16-byte vectors across a quarter warp (no bank conflicts).  Global loads of the
// next tile are issued into registers before the FMA loop of the current tile (software pipeline).
#pragma once
#include <cuda_runtime.h>

// Loader concept:
//   struct L { static constexpr bool kContigK; __device__ T operator()(int row, int k) const; };
// `row` is m for A, n for B; the loader must return 0 for out-of-range (row, k).
// Epilogue concept:
//   struct E { __device__ void operator()(int m, int n, T acc) const; };  // called only in range
//   optional: tile_begin / reductions are handled by specialised kernels, not here.

template <typename T, int BM, int BN, int BK, int TM, int TN>
struct GemmSimtConfig {
  static constexpr int kThreadsM = BM / TM;
  static constexpr int kThreadsN = BN / TN;
  static constexpr int kThreads = kThreadsM * kThreadsN;
  static constexpr int kALoads = BM * BK / kThreads;
  static constexpr int kBLoads = BN * BK / kThreads;
  static_assert(BM % TM == 0 && BN % TN == 0, "tile shape");
  static_assert((BM * BK) % kThreads == 0 && (BN * BK) % kThreads == 0, "loader shape");
  static_assert(TM % 2 == 0 && TN % 2 == 0, "split register tile");
  static_assert(((TM / 2) * sizeof(T)) % 16 == 0 && ((TN / 2) * sizeof(T)) % 16 == 0, "16-byte smem vectors");
  // row padding: keeps 16-byte alignment and spreads k-major stashes over banks
  static constexpr int kPad = 16 / sizeof(T);
  static constexpr int kLdA = BM + kPad;
  static constexpr int kLdB = BN + kPad;
};

// 16-byte shared-memory vector read of `COUNT` consecutive elements.
template <typename T, int COUNT>
__device__ __forceinline__ void smem_vec_read(T* dst, const T* src) {
  constexpr int kVecs = COUNT * sizeof(T) / 16;
#pragma unroll
  for (int v = 0; v < kVecs; ++v)
    reinterpret_cast<int4*>(dst)[v] = reinterpret_cast<const int4*>(src)[v];
}

template <typename T, int ROWS, int BK, int THREADS, bool CONTIG_K, class Loader>
__device__ __forceinline__ void gemm_tile_fetch(T (&regs)[ROWS * BK / THREADS], const Loader& ld,
                                                int row0, int k0, int tid) {
#pragma unroll
  for (int i = 0; i < ROWS * BK / THREADS; ++i) {
    const int e = tid + i * THREADS;
    int r, k;
    if (CONTIG_K) { k = e % BK; r = e / BK; } else { r = e % ROWS; k = e / ROWS; }
    regs[i] = ld(row0 + r, k0 + k);
  }
}

template <typename T, int ROWS, int LD, int BK, int THREADS, bool CONTIG_K>
__device__ __forceinline__ void gemm_tile_stash(T (*smem)[LD], const T (&regs)[ROWS * BK / THREADS], int tid) {
#pragma unroll
  for (int i = 0; i < ROWS * BK / THREADS; ++i) {
    const int e = tid + i * THREADS;
    int r, k;
    if (CONTIG_K) { k = e % BK; r = e / BK; } else { r = e % ROWS; k = e / ROWS; }
    smem[k][r] = regs[i];
  }
}

// Computes the accumulator tile for block (blockIdx.y -> m tile, blockIdx.x -> n tile).
// acc[i][j] belongs to row  m0 + (i < TM/2 ? ty*TM/2 + i : BM/2 + ty*TM/2 + i - TM/2)  and the
// matching column formula; use gemm_row()/gemm_col() to recover them.
template <typename T, int BM, int BN, int BK, int TM, int TN, class ALoad, class BLoad>
__device__ __forceinline__ void gemm_simt_mainloop(T (&acc)[TM][TN], int m0, int n0, int Kc,
                                                   const ALoad& aload, const BLoad& bload) {
  using Cfg = GemmSimtConfig<T, BM, BN, BK, TM, TN>;
  __shared__ __align__(16) T As[2][BK][Cfg::kLdA];
  __shared__ __align__(16) T Bs[2][BK][Cfg::kLdB];
  const int tid = threadIdx.x;
  const int tx = tid % Cfg::kThreadsN, ty = tid / Cfg::kThreadsN;

#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = T(0);

  T areg[Cfg::kALoads], breg[Cfg::kBLoads];
  gemm_tile_fetch<T, BM, BK, Cfg::kThreads, ALoad::kContigK>(areg, aload, m0, 0, tid);
  gemm_tile_fetch<T, BN, BK, Cfg::kThreads, BLoad::kContigK>(breg, bload, n0, 0, tid);
  gemm_tile_stash<T, BM, Cfg::kLdA, BK, Cfg::kThreads, ALoad::kContigK>(As[0], areg, tid);
  gemm_tile_stash<T, BN, Cfg::kLdB, BK, Cfg::kThreads, BLoad::kContigK>(Bs[0], breg, tid);
  __syncthreads();

  const int num_tiles = (Kc + BK - 1) / BK;
  for (int t = 0; t < num_tiles; ++t) {
    const int cur = t & 1;
    if (t + 1 < num_tiles) {
      gemm_tile_fetch<T, BM, BK, Cfg::kThreads, ALoad::kContigK>(areg, aload, m0, (t + 1) * BK, tid);
      gemm_tile_fetch<T, BN, BK, Cfg::kThreads, BLoad::kContigK>(breg, bload, n0, (t + 1) * BK, tid);
    }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      __align__(16) T a[TM];
      __align__(16) T b[TN];
      smem_vec_read<T, TM / 2>(a, &As[cur][kk][ty * (TM / 2)]);
      smem_vec_read<T, TM / 2>(a + TM / 2, &As[cur][kk][BM / 2 + ty * (TM / 2)]);
      smem_vec_read<T, TN / 2>(b, &Bs[cur][kk][tx * (TN / 2)]);
      smem_vec_read<T, TN / 2>(b + TN / 2, &Bs[cur][kk][BN / 2 + tx * (TN / 2)]);
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    if (t + 1 < num_tiles) {
      gemm_tile_stash<T, BM, Cfg::kLdA, BK, Cfg::kThreads, ALoad::kContigK>(As[cur ^ 1], areg, tid);
      gemm_tile_stash<T, BN, Cfg::kLdB, BK, Cfg::kThreads, BLoad::kContigK>(Bs[cur ^ 1], breg, tid);
    }
    __syncthreads();
  }
}

template <int BM, int TM, int THREADS_N>
__device__ __forceinline__ int gemm_row(int m0, int i) {
  const int ty = threadIdx.x / THREADS_N;
  return m0 + (i < TM / 2 ? ty * (TM / 2) + i : BM / 2 + ty * (TM / 2) + (i - TM / 2));
}
template <int BN, int TN, int THREADS_N>
__device__ __forceinline__ int gemm_col(int n0, int j) {
  const int tx = threadIdx.x % THREADS_N;
  return n0 + (j < TN / 2 ? tx * (TN / 2) + j : BN / 2 + tx * (TN / 2) + (j - TN / 2));
}

// Plain kernel: per-element epilogue.
template <typename T, int BM, int BN, int BK, int TM, int TN, class ALoad, class BLoad, class Epi>
__global__ void __launch_bounds__((BM / TM) * (BN / TN), (sizeof(T) == 4 ? 2 : 1))
gemm_simt_kernel(int M, int N, int Kc, ALoad aload, BLoad bload, Epi epi) {
  T acc[TM][TN];
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  gemm_simt_mainloop<T, BM, BN, BK, TM, TN>(acc, m0, n0, Kc, aload, bload);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = gemm_row<BM, TM, BN / TN>(m0, i);
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = gemm_col<BN, TN, BN / TN>(n0, j);
      if (n < N) epi(m, n, acc[i][j]);
    }
  }
}
