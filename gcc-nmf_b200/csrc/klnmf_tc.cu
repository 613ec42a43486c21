// Tensor-core (tcgen05 / TMEM) path of the KL-NMF contractions: 3xTF32 GEMM entry points.
#include "common.cuh"
#include "umma_gemm.cuh"

#include <algorithm>

namespace {

using umma::GemmArgs;

struct EpiStoreRowMajor {  // D[m][n0 .. n0+31] = v   (split z writes slab z)
  float* D; int64_t ldd; int M, N; int64_t slab;
  __device__ void operator()(int m, int n0, const float (&v)[32], int z) const {
    if (m >= M) return;
    float* row = D + (int64_t)z * slab + (int64_t)m * ldd;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (n0 + j < N) row[n0 + j] = v[j];
  }
};

template <int BN, int LW, class Epi>
int launch_gemm(gccnmf_handle* h, const GemmArgs& args, int m_tiles, int splits, const Epi& epi, void* stream) {
  using S = umma::GemmSmem<BN, LW>;
  auto kernel = umma::gemm_tn_3xtf32_kernel<BN, LW, Epi>;
  static bool configured = false;
  if (!configured) {
    GCCNMF_CHECK_CUDA(h, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    configured = true;
  }
  dim3 grid((args.N + BN - 1) / BN, m_tiles, splits);
  GCCNMF_LAUNCH(h, kernel, grid, (LW + 1) * 32, S::kTotal, stream, args, epi);
  return 0;
}

}  // namespace


// ------------------------------------------------------------------------------------------------
// KL-NMF driver on the tensor-core GEMM (reference order, gccNMFFunctions.py:75-81).
//
// Every contraction is a "TN" product with both operands k-contiguous, so each intermediate is kept
// in the orientation(s) its consumers contract over (the producing epilogue writes them):
//   W  (F, K)   ld K      A of G1/G3 (contract over atoms)        WT (K, Fp)  A of G2 (contract over f)
//   H  (K, T2)  ld T2p    B of G4   (contract over frames)        HT (T2, K)  B of G1/G3
//   V  (F, T2)  ld T2     epilogue of G3                          VT (T2, Fp) epilogue of G1
//   R  (F, T2p)           A of G4, written by G3                  RT (T2, Fp) B of G2, written by G1
// Fp, T2p = leading dimensions rounded up to 4 floats (16-byte rows); pad columns hold zeros.
//   G1: R^T = V^T / (W.H)          M = f, N = t, over atoms      transposed-write epilogue
//   G2: H  *= (W^T.R) / denom      M = atom, N = t, over f       writes H and H^T
//   G3: R   = V / (W.H)            M = f, N = t, over atoms      row-write epilogue
//   G4: partial[z] = R.H^T         M = f, N = atom, over frames, split over z
// The 513th frequency row (F = 4 x 128 + 1) is not worth a 128-row tile: rows past the last full
// tile (when fewer than kTailRowsMax) are computed by a warp-per-output SIMT dot kernel with the
// same epilogue functors.
namespace {

constexpr int kTailRowsMax = 8;

struct EpiRatioT {   // RT[n][m] = VT[n][m] / acc     (G1)
  const float* VT; float* RT; int64_t ld; int M, N;
  __device__ void elem(int m, int n, float acc, int) const { RT[(int64_t)n * ld + m] = VT[(int64_t)n * ld + m] / acc; }
  __device__ void operator()(int m, int n0, const float (&v)[32], int z) const {
    if (m >= M) return;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (n0 + j < N) elem(m, n0 + j, v[j], z);
  }
};
struct EpiRatioRow {  // R[m][n] = V[m][n] / acc       (G3)
  const float* V; float* R; int64_t ldv, ldr; int M, N;
  __device__ void elem(int m, int n, float acc, int) const { R[(int64_t)m * ldr + n] = V[(int64_t)m * ldv + n] / acc; }
  __device__ void operator()(int m, int n0, const float (&v)[32], int z) const {
    if (m >= M) return;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (n0 + j < N) elem(m, n0 + j, v[j], z);
  }
};
struct EpiUpdateHBoth {  // H[m][n] *= acc / denom[m]; HT[n][m] = same      (G2; m = atom, n = frame)
  float* H; float* HT; const float* colsumW; float alpha, eps; int64_t ldh, ldht; int M, N;
  __device__ void elem(int m, int n, float acc, int) const {
    const float denom = (colsumW[m] + alpha) + eps;
    float* hp = H + (int64_t)m * ldh + n;
    const float hv = *hp * (acc / denom);
    *hp = hv;
    HT[(int64_t)n * ldht + m] = hv;
  }
  __device__ void operator()(int m, int n0, const float (&v)[32], int z) const {
    if (m >= M) return;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (n0 + j < N) elem(m, n0 + j, v[j], z);
  }
};
struct EpiPartial {  // partial[z][m][n] = acc        (G4)
  float* P; int64_t ld; int M, N; int64_t slab;
  __device__ void elem(int m, int n, float acc, int z) const { P[(int64_t)z * slab + (int64_t)m * ld + n] = acc; }
  __device__ void operator()(int m, int n0, const float (&v)[32], int z) const {
    if (m >= M) return;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (n0 + j < N) elem(m, n0 + j, v[j], z);
  }
};

// Rows [row_begin, M) of D = A.B^T by one warp per output element (float32 FMA), k range of split z.
template <class Epi>
__global__ void __launch_bounds__(256)
gemm_tail_rows_kernel(GemmArgs args, int row_begin, Epi epi) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;
  const int m = row_begin + blockIdx.y;
  if (n >= args.N || m >= args.M) return;
  const int k_begin = blockIdx.z * args.kblocks_per_split * umma::kBK;
  const int k_end = min((args.Kc + 3) & ~3, k_begin + args.kblocks_per_split * umma::kBK);
  const float4* a = reinterpret_cast<const float4*>(args.A + (int64_t)m * args.lda);
  const float4* b = reinterpret_cast<const float4*>(args.B + (int64_t)n * args.ldb);
  float acc = 0.f;
  for (int k4 = k_begin / 4 + lane; k4 < k_end / 4; k4 += 32) {
    const float4 x = __ldg(a + k4), y = __ldg(b + k4);
    acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc); acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) epi.elem(m, n, acc, (int)blockIdx.z);
}

// dst (cols, ld_dst) = src (rows, ld_src)^T, zero-filling dst columns [rows, ld_dst).
__global__ void transpose_pad_kernel(const float* __restrict__ src, int rows, int cols, int64_t ld_src, float* __restrict__ dst, int64_t ld_dst) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? src[(int64_t)r * ld_src + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < ld_dst) dst[(int64_t)c * ld_dst + r] = tile[threadIdx.x][i];
  }
}

__global__ void tc_colsum_kernel(const float* W, int F, int K, float* colsum) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float s = 0.f;
  for (int f = 0; f < F; ++f) s += W[(int64_t)f * K + k];   // row order like numpy.sum(W, axis=0)
  colsum[k] = s;
}

__global__ void tc_rowsum_kernel(const float* H, int T2, int64_t ld, float* rowsum) {
  __shared__ float warp_sums[32];
  const float* row = H + (int64_t)blockIdx.x * ld;
  float s = 0.f;
  for (int t = threadIdx.x; t < T2; t += blockDim.x) s += row[t];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? warp_sums[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) rowsum[blockIdx.x] = s;
  }
}

// W *= (sum_z partial[z]) / rowsum(H) (:77); unit-L2 atoms (:79-80); also W^T, norms and colsum(W) for the next iteration.
constexpr int kTcApplyCols = 32, kTcApplyGroups = 32;
__global__ void __launch_bounds__(kTcApplyCols * kTcApplyGroups)
tc_apply_w_kernel(float* W, float* WT, int64_t ldwt, const float* partial, int splits, const float* rowsumH, int F, int K,
                  float* norms, float* colsum) {
  __shared__ float part[kTcApplyGroups][kTcApplyCols + 1];
  __shared__ float norm_s[kTcApplyCols];
  const int c = threadIdx.x % kTcApplyCols, g = threadIdx.x / kTcApplyCols;
  const int k = blockIdx.x * kTcApplyCols + c;
  const int rows_per_group = (F + kTcApplyGroups - 1) / kTcApplyGroups;
  const int f0 = g * rows_per_group, f1 = min(F, f0 + rows_per_group);
  const int64_t slab = (int64_t)F * K;
  float sumsq = 0.f;
  if (k < K) {
    const float rs = rowsumH[k];
    for (int f = f0; f < f1; ++f) {
      const int64_t i = (int64_t)f * K + k;
      float numer = partial[i];
      for (int z = 1; z < splits; ++z) numer += partial[(int64_t)z * slab + i];
      const float w = W[i] * (numer / rs);
      W[i] = w;
      sumsq += w * w;
    }
  }
  part[g][c] = sumsq;
  __syncthreads();
  if (g == 0) {
    float s = 0.f;
    for (int j = 0; j < kTcApplyGroups; ++j) s += part[j][c];
    const float nrm = sqrtf(s);
    norm_s[c] = nrm;
    if (k < K) norms[k] = nrm;
  }
  __syncthreads();
  float csum = 0.f;
  if (k < K) {
    const float nrm = norm_s[c];
    for (int f = f0; f < f1; ++f) {
      const int64_t i = (int64_t)f * K + k;
      const float w = W[i] / nrm;
      W[i] = w;
      WT[(int64_t)k * ldwt + f] = w;
      csum += w;
    }
  }
  part[g][c] = csum;
  __syncthreads();
  if (g == 0 && k < K) {
    float s = 0.f;
    for (int j = 0; j < kTcApplyGroups; ++j) s += part[j][c];
    colsum[k] = s;
  }
}

// H[k][t] *= norms[k] and HT[t][k] *= norms[k]   (:81)
__global__ void tc_scale_h_kernel(float* H, int64_t ldh, float* HT, int64_t ldht, const float* norms, int K, int T2) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)K * T2) return;
  const int k = (int)(i / T2), t = (int)(i - (int64_t)k * T2);
  H[(int64_t)k * ldh + t] *= norms[k];
  const int t2 = (int)(i / K), k2 = (int)(i - (int64_t)t2 * K);
  HT[(int64_t)t2 * ldht + k2] *= norms[k2];
}

struct TcWorkspace {
  float *HT, *WT, *VT, *R, *RT, *partial, *colsum, *rowsum, *norms;
  int64_t Fp, T2p;
  int splits;
  bool ok;
};

int tc_pick_splits(int tiles, int total_kblocks, int sm_count) {
  int s = 1;
  while (s < 8 && tiles * (s * 2) <= sm_count && total_kblocks / (s * 2) >= 8) s *= 2;
  return s;
}

TcWorkspace tc_carve(void* ws, size_t bytes, int F, int T2, int K, int sm_count) {
  WorkspaceCarver c(ws, bytes);
  TcWorkspace w;
  w.Fp = (F + 3) & ~3;
  w.T2p = (T2 + 3) & ~3;
  const int m_tiles_f = (F % umma::kBM != 0 && F % umma::kBM <= kTailRowsMax) ? F / umma::kBM : (F + umma::kBM - 1) / umma::kBM;
  w.splits = tc_pick_splits(std::max(1, m_tiles_f) * ((K + 127) / 128), (T2 + umma::kBK - 1) / umma::kBK, sm_count);
  w.HT = c.take<float>((size_t)T2 * K);
  w.WT = c.take<float>((size_t)K * w.Fp);
  w.VT = c.take<float>((size_t)T2 * w.Fp);
  w.R = c.take<float>((size_t)F * w.T2p);
  w.RT = c.take<float>((size_t)T2 * w.Fp);
  w.partial = c.take<float>((size_t)8 * F * K);
  w.colsum = c.take<float>(K);
  w.rowsum = c.take<float>(K);
  w.norms = c.take<float>(K);
  w.ok = c.ok();
  return w;
}

size_t tc_workspace_bytes(int F, int T2, int K) {
  const size_t Fp = (F + 3) & ~3, T2p = (T2 + 3) & ~3;
  size_t n = 0;
  auto add = [&](size_t count) { n = align_up(n, 256) + count * sizeof(float); };
  add((size_t)T2 * K); add((size_t)K * Fp); add((size_t)T2 * Fp); add((size_t)F * T2p); add((size_t)T2 * Fp);
  add((size_t)8 * F * K); add(K); add(K); add(K);
  return align_up(n, 256);
}

// Runs D = A.B^T over full 128-row tiles on the tensor cores and any short row tail on SIMT warps.
template <class Epi>
int tc_gemm(gccnmf_handle* h, GemmArgs args, int splits, const Epi& epi, void* stream) {
  const int total_kb = (args.Kc + umma::kBK - 1) / umma::kBK;
  args.kblocks_per_split = (total_kb + splits - 1) / splits;
  const int tail = args.M % umma::kBM;
  const bool simt_tail = tail != 0 && tail <= kTailRowsMax && args.M > umma::kBM;
  const int m_tiles = simt_tail ? args.M / umma::kBM : (args.M + umma::kBM - 1) / umma::kBM;
  GemmArgs main_args = args;
  if (simt_tail) main_args.M = m_tiles * umma::kBM;      // epilogue and loader stop at the last full tile
  const int tiles128 = m_tiles * ((args.N + 127) / 128);
  int st;
  if (splits == 1 && args.N >= 256 && tiles128 > h->sm_count + h->sm_count / 4)
    st = launch_gemm<256, 8>(h, main_args, m_tiles, splits, epi, stream);
  else
    st = launch_gemm<128, 4>(h, main_args, m_tiles, splits, epi, stream);
  if (st) return st;
  if (simt_tail) {
    auto kernel = gemm_tail_rows_kernel<Epi>;
    GCCNMF_LAUNCH(h, kernel, dim3((args.N + 7) / 8, tail, splits), 256, 0, stream, args, m_tiles * umma::kBM, epi);
  }
  return 0;
}

}  // namespace

// Whether the tensor-core path supports this problem (else the SIMT path in klnmf.cu is used).
bool gccnmf_klnmf_tc_supported(int F, int T2, int K) { return K % 4 == 0 && T2 % 4 == 0 && F >= 128 && T2 >= 128 && K >= 32; }
size_t gccnmf_klnmf_tc_workspace_bytes(int F, int T2, int K) { return tc_workspace_bytes(F, T2, K); }

// Transposes / pads the caller's V, W, H into the k-contiguous operand set.  H is used in place (ld T2).
int gccnmf_klnmf_tc_prepare(gccnmf_handle* h, const float* V, int F, int T2, const float* W, const float* H, int K,
                            void* workspace, size_t workspace_bytes, bool need_vt, bool need_wt, bool need_ht, void* stream) {
  TcWorkspace w = tc_carve(workspace, workspace_bytes, F, T2, K, h->sm_count);
  if (!w.ok) return gccnmf_fail(h, GCCNMF_ERR_WORKSPACE, "klnmf (tensor-core path) workspace too small: need %zu bytes", tc_workspace_bytes(F, T2, K));
  dim3 block(32, 8);
  if (need_vt) {
    GCCNMF_LAUNCH(h, transpose_pad_kernel, dim3((T2 + 31) / 32, (F + 31) / 32 + 1), block, 0, stream, V, F, T2, (int64_t)T2, w.VT, w.Fp);
    GCCNMF_CHECK_CUDA(h, cudaMemsetAsync(w.RT, 0, (size_t)T2 * w.Fp * sizeof(float), (cudaStream_t)stream));
  }
  if (need_wt) GCCNMF_LAUNCH(h, transpose_pad_kernel, dim3((K + 31) / 32, (F + 31) / 32 + 1), block, 0, stream, W, F, K, (int64_t)K, w.WT, w.Fp);
  if (need_ht) GCCNMF_LAUNCH(h, transpose_pad_kernel, dim3((T2 + 31) / 32, (K + 31) / 32), block, 0, stream, H, K, T2, (int64_t)T2, w.HT, (int64_t)K);
  return 0;
}

// :76  H *= (W^T (V / (W H))) / (colsum(W) + alpha + eps)   -- requires prepare(); keeps H and HT in sync
int gccnmf_klnmf_tc_update_H(gccnmf_handle* h, const float* V, int F, int T2, const float* W, float* H, int K, float alpha, float eps,
                             void* workspace, size_t workspace_bytes, bool have_colsum, void* stream) {
  TcWorkspace w = tc_carve(workspace, workspace_bytes, F, T2, K, h->sm_count);
  if (!w.ok) return gccnmf_fail(h, GCCNMF_ERR_WORKSPACE, "klnmf (tensor-core path) workspace too small");
  (void)V;
  {  // G1: RT = VT / (W.H)
    GemmArgs a{W, w.HT, F, T2, K, (int64_t)K, (int64_t)K, 0};
    EpiRatioT e{w.VT, w.RT, w.Fp, F, T2};
    if (int st = tc_gemm(h, a, 1, e, stream)) return st;
  }
  if (!have_colsum) GCCNMF_LAUNCH(h, tc_colsum_kernel, (K + 127) / 128, 128, 0, stream, W, F, K, w.colsum);
  {  // G2: H, HT *= (WT.RT^T) / denom
    GemmArgs a{w.WT, w.RT, K, T2, F, w.Fp, w.Fp, 0};
    EpiUpdateHBoth e{H, w.HT, w.colsum, alpha, eps, (int64_t)T2, (int64_t)K, K, T2};
    if (int st = tc_gemm(h, a, 1, e, stream)) return st;
  }
  return 0;
}

// :77 numerator: partial[z] = (V / (W H)) . H^T over the frame range of split z, and rowsum(H)
int gccnmf_klnmf_tc_partial_W(gccnmf_handle* h, const float* V, int F, int T2, const float* W, const float* H, int K,
                              void* workspace, size_t workspace_bytes, void* stream) {
  TcWorkspace w = tc_carve(workspace, workspace_bytes, F, T2, K, h->sm_count);
  if (!w.ok) return gccnmf_fail(h, GCCNMF_ERR_WORKSPACE, "klnmf (tensor-core path) workspace too small");
  {  // G3: R = V / (W.H)
    GemmArgs a{W, w.HT, F, T2, K, (int64_t)K, (int64_t)K, 0};
    EpiRatioRow e{V, w.R, (int64_t)T2, w.T2p, F, T2};
    if (int st = tc_gemm(h, a, 1, e, stream)) return st;
  }
  GCCNMF_LAUNCH(h, tc_rowsum_kernel, K, 256, 0, stream, H, T2, (int64_t)T2, w.rowsum);
  {  // G4: partial[z] = R.H^T
    GemmArgs a{w.R, H, F, K, T2, w.T2p, (int64_t)T2, 0};
    EpiPartial e{w.partial, (int64_t)K, F, K, (int64_t)F * K};
    if (int st = tc_gemm(h, a, w.splits, e, stream)) return st;
  }
  return 0;
}

// :77-:81 with the numerator taken from `numer` (F*K + K floats: all-reduced across ranks) when given,
// else from this rank's split partials.
int gccnmf_klnmf_tc_apply_W(gccnmf_handle* h, int F, int T2, float* W, float* H, int K, const float* numer,
                            void* workspace, size_t workspace_bytes, void* stream) {
  TcWorkspace w = tc_carve(workspace, workspace_bytes, F, T2, K, h->sm_count);
  if (!w.ok) return gccnmf_fail(h, GCCNMF_ERR_WORKSPACE, "klnmf (tensor-core path) workspace too small");
  const float* partial = numer ? numer : w.partial;
  const float* rowsum = numer ? numer + (int64_t)F * K : w.rowsum;
  GCCNMF_LAUNCH(h, tc_apply_w_kernel, (K + kTcApplyCols - 1) / kTcApplyCols, kTcApplyCols * kTcApplyGroups, 0, stream, W, w.WT, w.Fp,
                partial, numer ? 1 : w.splits, rowsum, F, K, w.norms, w.colsum);
  const int64_t n = (int64_t)K * T2;
  GCCNMF_LAUNCH(h, tc_scale_h_kernel, (unsigned)((n + 255) / 256), 256, 0, stream, H, (int64_t)T2, w.HT, (int64_t)K, w.norms, K, T2);
  return 0;
}

// Sums this rank's split partials and row sums into `numer` (F*K + K floats) for the all-reduce.
__global__ void tc_pack_numer_kernel(const float* partial, int splits, int64_t n, const float* rowsum, int K, float* numer) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float s = partial[i];
    for (int z = 1; z < splits; ++z) s += partial[(int64_t)z * n + i];
    numer[i] = s;
  } else if (i < n + K) {
    numer[i] = rowsum[i - n];
  }
}

int gccnmf_klnmf_tc_pack_numer(gccnmf_handle* h, int F, int T2, int K, float* numer, void* workspace, size_t workspace_bytes, void* stream) {
  TcWorkspace w = tc_carve(workspace, workspace_bytes, F, T2, K, h->sm_count);
  if (!w.ok) return gccnmf_fail(h, GCCNMF_ERR_WORKSPACE, "klnmf (tensor-core path) workspace too small");
  const int64_t n = (int64_t)F * K;
  GCCNMF_LAUNCH(h, tc_pack_numer_kernel, (unsigned)((n + K + 255) / 256), 256, 0, stream, w.partial, w.splits, n, w.rowsum, K, numer);
  return 0;
}

extern "C" {

// D (M, N) row-major (ldd) = A (M, Kc; lda) . B (N, Kc; ldb)^T with 3xTF32 error compensation.
// lda, ldb multiples of 4 with zero padding up to round_up(Kc, 4); tile_n in {128, 256}.
int gccnmf_gemm_tn_3xtf32(gccnmf_handle* h, const float* A, int64_t lda, const float* B, int64_t ldb, float* D,
                                     int64_t ldd, int M, int N, int Kc, int tile_n, void* stream) {
  if (!h) return GCCNMF_ERR_INVALID_ARGUMENT;
  GCCNMF_REQUIRE(h, A && B && D && M > 0 && N > 0 && Kc > 0, "gemm_tn_3xtf32: bad arguments");
  GCCNMF_REQUIRE(h, lda % 4 == 0 && ldb % 4 == 0 && lda >= ((Kc + 3) & ~3) && ldb >= ((Kc + 3) & ~3),
                 "gemm_tn_3xtf32: leading dimensions must be multiples of 4 covering round_up(Kc, 4)");
  GCCNMF_REQUIRE(h, (reinterpret_cast<uintptr_t>(A) % 16 == 0) && (reinterpret_cast<uintptr_t>(B) % 16 == 0),
                 "gemm_tn_3xtf32: operands must be 16-byte aligned");
  GemmArgs args{A, B, M, N, Kc, lda, ldb, (Kc + umma::kBK - 1) / umma::kBK};
  EpiStoreRowMajor epi{D, ldd, M, N, 0};
  const int m_tiles = (M + umma::kBM - 1) / umma::kBM;
  if (tile_n == 256) return launch_gemm<256, 8>(h, args, m_tiles, 1, epi, stream);
  if (tile_n == 128) return launch_gemm<128, 4>(h, args, m_tiles, 1, epi, stream);
  return gccnmf_fail(h, GCCNMF_ERR_UNSUPPORTED, "gemm_tn_3xtf32: tile_n must be 128 or 256");
}

}  // extern "C"
