// All-TDOA GCC-NMF argmax on the tensor cores with exact float64 refinement
// (reference: notebooks/offlineSpeechEnhancement.ipynb cells 27+29, :444-467; online :422-423).
//
//   gccNMF[k, tau, t] = sum_f W[f, k] * Re(C[f, t] E[f, tau])        argmax over tau per (k, t)
//
// The reference evaluates this in float64 and the argmax must be the reference's, bit for bit.  A
// float64 contraction is 126 GFLOP at the headline shape (10.9 ms on the SIMT float64 kernel), so:
//   1. build  G[(t, tau), f] = Re(C E)  once, float64 product rounded to float32            (HBM-bound)
//   2. argmax over tau of  W^T . G^T  with the 3xTF32 tcgen05 GEMM (M = atoms, N = (t, tau), over f):
//      the epilogue keeps, per (atom, frame), the best and second-best value and the index of the best;
//   3. every (atom, frame) whose margin best - second is below the worst-case error of step 2
//      (kMarginFactor * sum_f |W[f, atom]|, since |G| <= 1) is appended to a list and recomputed EXACTLY in
//      float64 from C, E and W by a warp (the same arithmetic as the float64 kernel in gcc.cu).
// Decisions with a safe margin cannot differ from the float64 ones; the others are the float64 ones.
#include <algorithm>

#include "common.cuh"
#include "umma_gemm.cuh"

namespace {

using umma::GemmArgs;

// Error budget of step 2 relative to sum_f |W| (|G| <= 1): 3xBF16 operand split <= 2^-17 per product (worst case, all
// coherent), float32 rounding of G 2^-24, accumulator truncation <= 2^-24 per accumulation x (3 F / 16) accumulations
// (measured 1.2e-5 at 384 accumulations, tests/test_gpu_umma.py): < 2e-5 per value.  Margin = 2 x that + slack: two
// values each off by the bound.
constexpr float kMarginFactor = 6e-5f;

// ------------------------------------------------------------------ step 1: G[(t, tau)][f]
// CTA = 32 bins x 32 frames x all TDOAs.  Warp w owns the TDOAs d = w, w + 8, ... ; lane = bin f, so every store is one
// 128-byte row segment of G.  E[f][d] is loaded once per (thread, d) -- its rows are D * 16 bytes apart, so a warp's load
// touches 32 lines: doing it inside the frame loop made the kernel L1-wavefront-bound (321 us for 247 MB) -- and reused for
// the 32 frames of the tile, whose coherence values sit in shared memory.
__global__ void __launch_bounds__(256)
build_gcc_matrix_kernel(const float2* __restrict__ coh, int F, int T, const double2* __restrict__ E, int D, float* __restrict__ G, int64_t ldg) {
  __shared__ float2 Cs[32][33];   // [f][t]
  const int f0 = blockIdx.x * 32, t0 = blockIdx.y * 32;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int i = w; i < 32; i += 8) {
    const int f = f0 + i, t = t0 + lane;
    Cs[i][lane] = (f < F && t < T) ? coh[(int64_t)f * T + t] : float2{0.f, 0.f};
  }
  __syncthreads();
  const int f = f0 + lane;
  if (f >= ldg) return;
  const int t_end = min(32, T - t0);
  for (int d0 = w; d0 < D; d0 += 32) {            // 4 TDOAs per pass: d0, d0 + 8, d0 + 16, d0 + 24
    double2 e[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int d = d0 + 8 * j;
      e[j] = (f < F && d < D) ? __ldg(E + (int64_t)f * D + d) : double2{0.0, 0.0};
    }
    for (int tt = 0; tt < t_end; ++tt) {
      const float2 c = Cs[lane][tt];
      float* row = G + ((int64_t)(t0 + tt) * D + d0) * ldg + f;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (d0 + 8 * j < D) row[(int64_t)(8 * j) * ldg] = (float)((double)c.x * e[j].x - (double)c.y * e[j].y);
    }
  }
}

// ------------------------------------------------------------------ step 2: epilogue
struct EpiArgmaxTDOA {
  struct State { float best, second; int idx; int nan; };
  int32_t* __restrict__ argmax;        // (K, T)
  const float* __restrict__ colsumW;   // (K) sum_f |W[f,k]| = colsum (W >= 0)
  int2* __restrict__ list; int* __restrict__ count; int capacity;
  int K, T, D, N;
  __device__ void init(State& s) const { s.best = -INFINITY; s.second = -INFINITY; s.idx = 0; s.nan = 0; }
  __device__ void elem(int, int, float, int) const {}   // M = atoms is tiled without SIMT tail rows (K % 128 handled by predication)
  __device__ void tile(int m_base, int lane, int n0, float (&v)[32], int, int, float*, State& s) const {
    const int d0 = n0 % D;
    if (d0 == 0) init(s);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float x = v[j];
      if (x != x) s.nan = 1;                 // NaN anywhere: let float64 decide with numpy's NaN rules
      if (x > s.best) { s.second = s.best; s.best = x; s.idx = d0 + j; }
      else if (x > s.second) s.second = x;
    }
    if (d0 + 32 == D) {
      const int m = m_base + lane, t = n0 / D;
      if (m < K && n0 < N) {
        argmax[(int64_t)m * T + t] = s.idx;
        const float margin = kMarginFactor * colsumW[m];
        if (s.nan || !(s.best - s.second > margin)) {
          const int slot = atomicAdd(count, 1);
          if (slot < capacity) list[slot] = make_int2(m, t);
        }
      }
    }
  }
};

// ------------------------------------------------------------------ step 3: exact float64 recomputation of flagged (atom, frame) pairs
__device__ __forceinline__ bool argmax_better64(double v, int i, double bv, int bi) {   // numpy.argmax: NaN is a maximum, first wins
  const bool vn = v != v, bn = bv != bv;
  if (vn || bn) return vn && (!bn || i < bi);
  return v > bv || (v == bv && i < bi);
}

// Reference version: one warp per flagged (atom, frame), lane = TDOA (+ 32 j), every bin's row of E read from L2 by every
// warp (13 485 pairs x 513 rows x 1 KB = 7 GB of L2 reads at the headline shape: L2-bandwidth bound, 0.8 ms).
__global__ void __launch_bounds__(256)
refine_argmax_kernel(const int2* __restrict__ list, const int* __restrict__ count, int capacity, const float2* __restrict__ coh,
                     int F, int T, const double2* __restrict__ E, int D, const float* __restrict__ W, int K, int32_t* __restrict__ argmax) {
  const int lane = threadIdx.x & 31;
  const int warps = gridDim.x * (blockDim.x >> 5);
  const int n = min(*count, capacity);
  for (int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); p < n; p += warps) {
    const int k = list[p].x, t = list[p].y;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};   // TDOAs lane, lane + 32, lane + 64, lane + 96 (D <= 128)
    for (int f = 0; f < F; ++f) {
      const float2 c = __ldg(coh + (int64_t)f * T + t);
      const double w = (double)__ldg(W + (int64_t)f * K + k);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int d = lane + 32 * j;
        if (d < D) {
          const double2 e = __ldg(E + (int64_t)f * D + d);
          acc[j] = fma((double)c.x * e.x - (double)c.y * e.y, w, acc[j]);
        }
      }
    }
    double bv = acc[0];
    int bi = lane;
#pragma unroll
    for (int j = 1; j < 4; ++j)
      if (lane + 32 * j < D && argmax_better64(acc[j], lane + 32 * j, bv, bi)) { bv = acc[j]; bi = lane + 32 * j; }
    if (lane >= D) { bv = -INFINITY; bi = 1 << 30; }
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (argmax_better64(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) argmax[(int64_t)k * T + t] = bi;
  }
}


// Same arithmetic (bin order, one fma per bin: the float64 kernel's), with E staged through shared memory in chunks of
// kRefineChunk bins and shared by the CTA's 8 warps x 4 pairs per warp: 32 x less L2 traffic for E.  The (strided)
// coherence and W values of a chunk are fetched by the lanes -- lane l < 16: coherence of bin l for the warp's 4 pairs,
// lane 16 + l: W -- and broadcast with shuffles.
constexpr int kRefineChunk = 16, kRefinePairs = 4;
__global__ void __launch_bounds__(256)
refine_argmax_shared_kernel(const int2* __restrict__ list, const int* __restrict__ count, int capacity, const float2* __restrict__ coh,
                            int F, int T, const double2* __restrict__ E, int D, const float* __restrict__ W, int K, int32_t* __restrict__ argmax) {
  __shared__ double2 Es[kRefineChunk][64];        // D <= 64
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = min(*count, capacity);
  const int per_cta = 8 * kRefinePairs;
  for (int base = blockIdx.x * per_cta; base < n; base += gridDim.x * per_cta) {      // (uniform per CTA: the barriers below are safe)
    int pk[kRefinePairs], pt[kRefinePairs];
#pragma unroll
    for (int p = 0; p < kRefinePairs; ++p) {
      const int idx = base + warp * kRefinePairs + p;
      const int2 kt = idx < n ? list[idx] : make_int2(0, 0);
      pk[p] = kt.x; pt[p] = kt.y;
    }
    double acc[kRefinePairs][2];
#pragma unroll
    for (int p = 0; p < kRefinePairs; ++p) { acc[p][0] = 0.0; acc[p][1] = 0.0; }
    for (int f0 = 0; f0 < F; f0 += kRefineChunk) {
      __syncthreads();                            // the previous chunk has been consumed
      for (int i = threadIdx.x; i < kRefineChunk * 64; i += 256) {
        const int fl = i >> 6, d = i & 63, f = f0 + fl;
        Es[fl][d] = (f < F && d < D) ? __ldg(E + (int64_t)f * D + d) : double2{0.0, 0.0};
      }
      // lane l < 16: coherence of bin f0 + l; lane >= 16: W of bin f0 + l - 16 (both for the warp's 4 pairs)
      const int fm = min(f0 + (lane & 15), F - 1);
      float cx[kRefinePairs], cy[kRefinePairs];   // (for lanes >= 16, cx carries W)
#pragma unroll
      for (int p = 0; p < kRefinePairs; ++p) {
        if (lane < 16) {
          const float2 c = __ldg(coh + (int64_t)fm * T + pt[p]);
          cx[p] = c.x; cy[p] = c.y;
        } else {
          cx[p] = __ldg(W + (int64_t)fm * K + pk[p]);
          cy[p] = 0.f;
        }
      }
      __syncthreads();
      const int fcount = min(kRefineChunk, F - f0);
      for (int fl = 0; fl < fcount; ++fl) {
        const double2 e0 = Es[fl][lane], e1 = Es[fl][lane + 32];
#pragma unroll
        for (int p = 0; p < kRefinePairs; ++p) {
          const double c_re = (double)__shfl_sync(0xffffffffu, cx[p], fl), c_im = (double)__shfl_sync(0xffffffffu, cy[p], fl);
          const double wd = (double)__shfl_sync(0xffffffffu, cx[p], 16 + fl);
          acc[p][0] = fma(c_re * e0.x - c_im * e0.y, wd, acc[p][0]);
          acc[p][1] = fma(c_re * e1.x - c_im * e1.y, wd, acc[p][1]);
        }
      }
    }
#pragma unroll
    for (int p = 0; p < kRefinePairs; ++p) {
      double bv = acc[p][0];
      int bi = lane;
      if (lane + 32 < D && argmax_better64(acc[p][1], lane + 32, bv, bi)) { bv = acc[p][1]; bi = lane + 32; }
      if (lane >= D) { bv = -INFINITY; bi = 1 << 30; }
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (argmax_better64(ov, oi, bv, bi)) { bv = ov; bi = oi; }
      }
      const int idx = base + warp * kRefinePairs + p;
      if (lane == 0 && idx < n) argmax[(int64_t)pk[p] * T + pt[p]] = bi;
    }
  }
}

__global__ void transpose_w_kernel(const float* __restrict__ W, int F, int K, float* __restrict__ WT, int64_t ldwt, float* __restrict__ colsum) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int f = f0 + i, k = k0 + threadIdx.x;
    tile[i][threadIdx.x] = (f < F && k < K) ? W[(int64_t)f * K + k] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int k = k0 + i, f = f0 + threadIdx.x;
    if (k < K && f < ldwt) WT[(int64_t)k * ldwt + f] = tile[threadIdx.x][i];
  }
  (void)colsum;
}

__global__ void abs_colsum_kernel(const float* __restrict__ W, int F, int K, float* __restrict__ colsum) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float s = 0.f;
  for (int f = 0; f < F; ++f) s += fabsf(W[(int64_t)f * K + k]);
  colsum[k] = s;
}

struct ArgmaxWorkspace {
  float *G, *WT, *colsum;
  int2* list;
  int* count;
  int64_t Fp;
  int capacity;
  bool ok;
};

int list_capacity(int K, int T) { return (int)std::min<int64_t>((int64_t)K * T, std::max<int64_t>(1 << 16, (int64_t)K * T / 8)); }

ArgmaxWorkspace carve_argmax(void* ws, size_t bytes, int F, int T, int D, int K) {
  WorkspaceCarver c(ws, bytes);
  ArgmaxWorkspace w;
  w.Fp = (F + 3) & ~3;
  w.capacity = list_capacity(K, T);
  w.G = c.take<float>((size_t)T * D * w.Fp);
  w.WT = c.take<float>((size_t)K * w.Fp);
  w.colsum = c.take<float>(K);
  w.list = c.take<int2>(w.capacity);
  w.count = c.take<int>(4);
  w.ok = c.ok();
  return w;
}

template <class Epi>
int launch_argmax_gemm(gccnmf_handle* h, const GemmArgs& args, const Epi& epi, void* stream) {
  using S = umma::GemmSmem<128, umma::kSplitBF16>;
  auto kernel = umma::gemm_tn_3xtf32_kernel<128, false, umma::kSplitBF16, umma::kLoaderWarps, Epi>;
  static DeviceFlags configured;     // per device: the attribute belongs to the device's copy of the kernel
  if (!configured(h)) {
    GCCNMF_CHECK_CUDA(h, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    configured(h) = true;
  }
  // x = atom tiles (fastest) so that the CTAs sharing one slab of G run together and it is read from HBM once
  dim3 grid(args.m_tiles, (args.N + 127) / 128, 1);
  GCCNMF_LAUNCH(h, kernel, grid, umma::kThreads, S::kTotal, stream, args, epi);
  return 0;
}

}  // namespace

bool gccnmf_tdoa_argmax_tc_supported(int F, int T, int D, int K) {
  return (D == 32 || D == 64) && K % 4 == 0 && K >= 32 && F >= 64 && (int64_t)T * D >= 128 && (int64_t)T * D < ((int64_t)1 << 31);
}

extern "C" {

int gccnmf_tdoa_argmax_refine_capacity(int K, int T) { return (K > 0 && T > 0) ? list_capacity(K, T) : 0; }

size_t gccnmf_tdoa_argmax_workspace_bytes(int F, int T, int D, int K) {
  if (F <= 0 || T <= 0 || D <= 0 || K <= 0) return 0;
  if (!gccnmf_tdoa_argmax_tc_supported(F, T, D, K)) return 256;
  const size_t Fp = (F + 3) & ~3;
  size_t n = 0;
  auto add = [&](size_t b) { n = align_up(n, 256) + b; };
  add((size_t)T * D * Fp * 4); add((size_t)K * Fp * 4); add((size_t)K * 4); add((size_t)list_capacity(K, T) * 8); add(16);
  return align_up(n, 256);
}

int gccnmf_tdoa_argmax(gccnmf_handle* h, const float* coherence, int F, int T, const double* E, int D, const float* W, int K,
                       int32_t* argmax, int32_t* overflow_flag, void* workspace, size_t workspace_bytes, void* stream) {
  GCCNMF_ENTER(h);
  GCCNMF_REQUIRE(h, F > 0 && T > 0 && D > 0 && K > 0 && coherence && E && W && argmax, "tdoa_argmax: bad arguments");
  if (h->force_simt_nmf || !gccnmf_tdoa_argmax_tc_supported(F, T, D, K)) {
    // exact float64 SIMT kernel: nothing to refine, and the caller must not read an unwritten counter
    if (overflow_flag) GCCNMF_CHECK_CUDA(h, cudaMemsetAsync(overflow_flag, 0, sizeof(int32_t), (cudaStream_t)stream));
    return gccnmf_tdoa_gccnmf(h, coherence, F, T, E, D, W, K, nullptr, argmax, stream);
  }
  ArgmaxWorkspace w = carve_argmax(workspace, workspace_bytes, F, T, D, K);
  if (!w.ok) return gccnmf_fail(h, GCCNMF_ERR_WORKSPACE, "tdoa_argmax workspace too small: need %zu bytes", gccnmf_tdoa_argmax_workspace_bytes(F, T, D, K));
  GCCNMF_CHECK_CUDA(h, cudaMemsetAsync(w.count, 0, 16, (cudaStream_t)stream));
  GCCNMF_LAUNCH(h, build_gcc_matrix_kernel, dim3((int)((w.Fp + 31) / 32), (T + 31) / 32), 256, 0, stream,
                reinterpret_cast<const float2*>(coherence), F, T, reinterpret_cast<const double2*>(E), D, w.G, w.Fp);
  GCCNMF_LAUNCH(h, transpose_w_kernel, dim3((K + 31) / 32, (int)((w.Fp + 31) / 32)), dim3(32, 8), 0, stream, W, F, K, w.WT, w.Fp, w.colsum);
  GCCNMF_LAUNCH(h, abs_colsum_kernel, (K + 127) / 128, 128, 0, stream, W, F, K, w.colsum);
  const int N = T * D;
  GemmArgs args{w.WT, w.G, K, N, F, w.Fp, w.Fp, (F + umma::kBK - 1) / umma::kBK, (K + umma::kBM - 1) / umma::kBM, nullptr, nullptr, 1};
  EpiArgmaxTDOA epi{argmax, w.colsum, w.list, w.count, w.capacity, K, T, D, N};
  if (int st = launch_argmax_gemm(h, args, epi, stream)) return st;
  if (h->argmax_refine_shared && D <= 64) {
    GCCNMF_LAUNCH(h, refine_argmax_shared_kernel, h->sm_count * 4, 256, 0, stream, w.list, w.count, w.capacity,
                  reinterpret_cast<const float2*>(coherence), F, T, reinterpret_cast<const double2*>(E), D, W, K, argmax);
  } else {
    GCCNMF_LAUNCH(h, refine_argmax_kernel, h->sm_count * 4, 256, 0, stream, w.list, w.count, w.capacity,
                  reinterpret_cast<const float2*>(coherence), F, T, reinterpret_cast<const double2*>(E), D, W, K, argmax);
  }
  // more near-ties than the list holds (never seen: the list holds 1/8 of all decisions): the caller must fall back
  if (overflow_flag) GCCNMF_CHECK_CUDA(h, cudaMemcpyAsync(overflow_flag, w.count, sizeof(int), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return GCCNMF_OK;
}

}  // extern "C"
