// Tensor-core (tcgen05 / TMEM) path of KL-NMF (reference: gccNMF/gccNMFFunctions.py:69-83).
//
// Every contraction is a "TN" product with both operands k-contiguous (umma_gemm.cuh), so each
// matrix is kept in the orientation(s) its consumers contract over; the producing epilogue writes them:
//   W  (F, K)   ld K      A of G1/G3 (contract over atoms)        WT (K, Fp)  A of G2 (contract over f)
//   Hp (K, T2p)           B of G4   (contract over frames)        HT (T2, K)  B of G1/G3
//   V  (F, T2)  ld T2     epilogue of G3                          VT (T2, Fp) epilogue of G1
//   R  (F, T2p)           A of G4, written by G3                  RT (T2, Fp) B of G2, written by G1
// Fp, T2p = F, T2 rounded up to 4 floats (16-byte rows); pad columns hold zeros.  Hp is the working copy of
// the caller's H (any T2): it is read once at the start and written once at the end (gccnmf_klnmf_tc_finish).
//
// One iteration, reference order (:76-:81):
//   G1  RT = VT / (W.(n*H))          M = f, N = t, over atoms     n = pending atom norms (see below)
//   G2  H, HT = (n*H) * (WT.RT^T) / (colsum(W) + alpha + eps)      M = atom, N = t, over f; + row-sum partials
//   G3  R = V / (W.H)                M = f, N = t, over atoms
//   G4  partial[z] = R.H^T           M = f, N = atom, over frames, split over z
//   A   W *= sum_z partial / rowsum(H); unit-L2 atoms; WT, colsum(W), n = norms
// The rescaling H *= n (:81) is applied lazily with the reference's own float32 products: G1's loader
// multiplies H^T by n while splitting the operand and G2's epilogue multiplies the old H by n, so the
// 30 MB of H / H^T are not rewritten every iteration; one scale pass runs after the last iteration.
// F = 513 = 4 x 128 + 1: rows past the last full 128-row tile (at most kTailRowsMax) are computed by
// extra SIMT CTAs of the same launch (umma_gemm.cuh), on SMs the tile grid leaves idle.
#include <algorithm>

#include "common.cuh"
#include "umma_gemm.cuh"

namespace {

using umma::GemmArgs;
using umma::warp_transpose_32x32;

constexpr int kTailRowsMax = 8;
constexpr int kMaxSplits = 8;
constexpr int kNmfWorkerWarps = 8;    // 2 loader groups of 128 threads (16 warps / 4 groups measured no faster: the loop is L2 / shared-memory bound)

// ------------------------------------------------------------------------------------------------ epilogues
struct EpiStoreRowMajor {  // D[z][m][n] = acc   (test entry, G4 partials)
  UMMA_EPILOGUE_STATELESS
  float* __restrict__ D; int64_t ldd; int M, N; int64_t slab;
  __device__ void elem(int m, int n, float acc, int z) const { D[(int64_t)z * slab + (int64_t)m * ldd + n] = acc; }
  __device__ void tile(int m_base, int lane, int n0, float (&v)[32], int z, int, float* scratch, State&) const {
    warp_transpose_32x32(v, scratch, lane);             // v[i] = D[m_base + i][n0 + lane]
    const int n = n0 + lane;
    if (n >= N) return;
    float* out = D + (int64_t)z * slab + n;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (m_base + i < M) out[(int64_t)(m_base + i) * ldd] = v[i];
  }
};

struct EpiRatioT {   // RT[n][m] = VT[n][m] / acc     (G1; lanes run along m: coalesced as is)
  UMMA_EPILOGUE_STATELESS
  const float* __restrict__ VT; float* __restrict__ RT; int64_t ld; int M, N;
  __device__ void elem(int m, int n, float acc, int) const { RT[(int64_t)n * ld + m] = VT[(int64_t)n * ld + m] / acc; }
  __device__ void tile(int m_base, int lane, int n0, float (&v)[32], int, int, float*, State&) const {
    const int m = m_base + lane;
    if (m >= M) return;
    float vt[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) vt[j] = (n0 + j < N) ? __ldg(VT + (int64_t)(n0 + j) * ld + m) : 1.f;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (n0 + j < N) RT[(int64_t)(n0 + j) * ld + m] = vt[j] / v[j];
  }
};

struct EpiRatioRow {  // R[m][n] = V[m][n] / acc       (G3; transposed through shared memory for coalesced rows)
  UMMA_EPILOGUE_STATELESS
  const float* __restrict__ V; float* __restrict__ R; int64_t ldv, ldr; int M, N;
  __device__ void elem(int m, int n, float acc, int) const { R[(int64_t)m * ldr + n] = V[(int64_t)m * ldv + n] / acc; }
  __device__ void tile(int m_base, int lane, int n0, float (&v)[32], int, int, float* scratch, State&) const {
    warp_transpose_32x32(v, scratch, lane);
    const int n = n0 + lane;
    if (n >= N) return;
    float vv[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) vv[i] = (m_base + i < M) ? __ldg(V + (int64_t)(m_base + i) * ldv + n) : 1.f;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (m_base + i < M) R[(int64_t)(m_base + i) * ldr + n] = vv[i] / v[i];
  }
};

// G2 (m = atom, n = frame): new = (old * pending_norm[m]) * (acc / (colsum[m] + alpha + eps))  (:81 then :76).
// The old value is read from H^T (coalesced along m), H^T is rewritten in place, H is written through the
// shared-memory transpose, and the per-row partial sums of the new H go to rowsum_part[slot][m].
struct EpiUpdateHBoth {
  UMMA_EPILOGUE_STATELESS
  float* __restrict__ H; float* __restrict__ HT; const float* __restrict__ colsumW; const float* __restrict__ pending;
  float* __restrict__ rowsum_part; float alpha, eps; int64_t ldh, ldht; int M, N; int colsum_slots;
  __device__ float colsum(int m) const {   // colsum(W): sum of the W update's per-row-block partials
    float s = colsumW[m];
    for (int b = 1; b < colsum_slots; ++b) s += colsumW[(int64_t)b * M + m];
    return s;
  }
  __device__ float update(int m, int n, float acc) const {
    const float denom = (colsum(m) + alpha) + eps;
    float old = HT[(int64_t)n * ldht + m];
    if (pending) old = old * pending[m];
    return old * (acc / denom);
  }
  __device__ void elem(int m, int n, float acc, int) const {   // tail rows (not used for M = atoms, kept for completeness)
    const float hv = update(m, n, acc);
    HT[(int64_t)n * ldht + m] = hv;
    H[(int64_t)m * ldh + n] = hv;
  }
  __device__ void tile(int m_base, int lane, int n0, float (&v)[32], int, int slot, float* scratch, State&) const {
    const int m = m_base + lane;
    float rsum = 0.f;
    if (m < M) {
      const float denom = (colsum(m) + alpha) + eps;
      const float pn = pending ? pending[m] : 1.f;
      float old[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) old[j] = (n0 + j < N) ? HT[(int64_t)(n0 + j) * ldht + m] : 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float o = old[j];
        if (pending) o = o * pn;
        const float hv = (n0 + j < N) ? o * (v[j] / denom) : 0.f;
        if (n0 + j < N) HT[(int64_t)(n0 + j) * ldht + m] = hv;
        v[j] = hv;
        rsum += hv;
      }
      atomicAdd(rowsum_part + (int64_t)slot * M + m, rsum);   // one writer per (slot, m) per chunk: plain accumulation, deterministic order
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0.f;
    }
    warp_transpose_32x32(v, scratch, lane);
    const int n = n0 + lane;
    if (n >= N) return;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (m_base + i < M) H[(int64_t)(m_base + i) * ldh + n] = v[i];
  }
};

// ------------------------------------------------------------------------------------------------ small kernels
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

// W update, two fully parallel passes over 32 x 32 tiles (grid: atoms / 32 x rows / 32):
//   pass 1  W' = W * (sum_z partial[z]) / rowsum(H)  (:77), per-tile column sums of squares -> sumsq_part[row block][atom]
//   pass 2  norm = sqrt(sum of the row-block partials) (:79); W = W' / norm (:80); W^T through a shared-memory transpose;
//           per-tile column sums -> colsum_part[row block][atom] (the next H update's colsum(W)); norms[atom].
// rowsum(H) = sum of `rowsum_slots` partial vectors (G2's epilogue) or one all-reduced vector.
// Cross-rank sum read straight from the NVSwitch: `p` is the MULTICAST address of a symmetric buffer; the switch
// returns the float32 sum of the word at that offset over all GPUs of the multicast group (NVLS in-switch reduction).
__device__ __forceinline__ float multimem_sum_f32(const float* p) {
  float v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}

constexpr int kApplyTile = 32;
template <bool MULTIMEM>
__global__ void __launch_bounds__(kApplyTile * 8)
tc_apply_w1_kernel(float* __restrict__ W, const float* __restrict__ partial, int splits, const float* __restrict__ rowsum,
                   int rowsum_slots, int F, int K, float* __restrict__ sumsq_part) {
  __shared__ float rs_s[kApplyTile];
  __shared__ float part[8][kApplyTile + 1];
  const int c = threadIdx.x, g = threadIdx.y;
  const int k = blockIdx.x * kApplyTile + c;
  const int64_t slab = (int64_t)F * K;
  if (g == 0) {
    float rs = 0.f;
    if (k < K) {
      if (MULTIMEM) rs = multimem_sum_f32(rowsum + k);
      else
        for (int s = 0; s < rowsum_slots; ++s) rs += rowsum[(int64_t)s * K + k];
    }
    rs_s[c] = rs;
  }
  __syncthreads();
  float sumsq = 0.f;
  if (k < K) {
    const float rs = rs_s[c];
#pragma unroll
    for (int r = 0; r < kApplyTile / 8; ++r) {
      const int f = blockIdx.y * kApplyTile + g + 8 * r;
      if (f < F) {
        const int64_t i = (int64_t)f * K + k;
        float numer;
        if (MULTIMEM) {
          numer = multimem_sum_f32(partial + i);      // sum over ranks, reduced inside the NVSwitch
        } else {
          float p[kMaxSplits];                       // all split partials in flight at once, then summed in split order
#pragma unroll
          for (int z = 0; z < kMaxSplits; ++z) p[z] = z < splits ? __ldg(partial + (int64_t)z * slab + i) : 0.f;
          numer = p[0];
#pragma unroll
          for (int z = 1; z < kMaxSplits; ++z)
            if (z < splits) numer += p[z];
        }
        const float w = W[i] * (numer / rs);
        W[i] = w;
        sumsq += w * w;
      }
    }
  }
  part[g][c] = sumsq;
  __syncthreads();
  if (g == 0 && k < K) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += part[j][c];
    sumsq_part[(int64_t)blockIdx.y * K + k] = s;
  }
}

__global__ void __launch_bounds__(kApplyTile * 8)
tc_apply_w2_kernel(float* __restrict__ W, float* __restrict__ WT, int64_t ldwt, const float* __restrict__ sumsq_part, int row_blocks,
                   int F, int K, float* __restrict__ norms, float* __restrict__ colsum_part) {
  __shared__ float norm_s[kApplyTile];
  __shared__ float part[8][kApplyTile + 1];
  __shared__ float tile[kApplyTile][kApplyTile + 1];
  const int c = threadIdx.x, g = threadIdx.y;
  const int k = blockIdx.x * kApplyTile + c;
  if (g == 0) {
    float s = 0.f;
    if (k < K)
      for (int b = 0; b < row_blocks; ++b) s += sumsq_part[(int64_t)b * K + k];
    const float nrm = sqrtf(s);
    norm_s[c] = nrm;
    if (k < K && blockIdx.y == 0) norms[k] = nrm;
  }
  __syncthreads();
  const float nrm = norm_s[c];
  float csum = 0.f;
#pragma unroll
  for (int r = 0; r < kApplyTile / 8; ++r) {
    const int fl = g + 8 * r, f = blockIdx.y * kApplyTile + fl;
    float w = 0.f;
    if (k < K && f < F) {
      const int64_t i = (int64_t)f * K + k;
      w = W[i] / nrm;
      W[i] = w;
      csum += w;
    }
    tile[fl][c] = w;
  }
  part[g][c] = csum;
  __syncthreads();
  if (g == 0 && k < K) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += part[j][c];
    colsum_part[(int64_t)blockIdx.y * K + k] = s;
  }
  // transposed write: WT[atom = block.x*32 + row-of-thread][f = block.y*32 + c]
#pragma unroll
  for (int r = 0; r < kApplyTile / 8; ++r) {
    const int kl = g + 8 * r, kt = blockIdx.x * kApplyTile + kl, ft = blockIdx.y * kApplyTile + c;
    if (kt < K && ft < F) WT[(int64_t)kt * ldwt + ft] = tile[c][kl];
  }
}

// dst (rows, ld_dst) = src (rows, cols; ld_src), zero in the pad columns.
__global__ void tc_copy_pad_kernel(const float* __restrict__ src, int64_t ld_src, int cols, float* __restrict__ dst, int64_t ld_dst, int rows) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * ld_dst) return;
  const int r = (int)(i / ld_dst), c = (int)(i - (int64_t)r * ld_dst);
  dst[i] = c < cols ? src[(int64_t)r * ld_src + c] : 0.f;
}

// H (caller, ld T2) = Hp * norms (pending :81) or a plain copy.
__global__ void tc_finish_h_kernel(const float* __restrict__ Hp, int64_t ldp, const float* __restrict__ norms, float* __restrict__ H, int K, int T2) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)K * T2) return;
  const int k = (int)(i / T2), t = (int)(i - (int64_t)k * T2);
  const float v = Hp[(int64_t)k * ldp + t];
  H[i] = norms ? v * norms[k] : v;
}

// numer = [sum_z partial[z] (F*K) | sum_s rowsum_part[s] (K)] for the cross-rank all-reduce.
__global__ void tc_pack_numer_kernel(const float* partial, int splits, int64_t n, const float* rowsum, int rowsum_slots, int K, float* numer) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float s = partial[i];
    for (int z = 1; z < splits; ++z) s += partial[(int64_t)z * n + i];
    numer[i] = s;
  } else if (i < n + K) {
    float s = 0.f;
    for (int j = 0; j < rowsum_slots; ++j) s += rowsum[(int64_t)j * K + (i - n)];
    numer[i] = s;
  }
}

// ------------------------------------------------------------------------------------------------ launch helpers
template <int BN, bool SCALE_B, int SPLIT, class Epi>
int launch_gemm_split(gccnmf_handle* h, const GemmArgs& args, int tail_rows, int splits, const Epi& epi, void* stream) {
  using S = umma::GemmSmem<BN, SPLIT>;
  constexpr int LW = kNmfWorkerWarps;
  auto kernel = umma::gemm_tn_3xtf32_kernel<BN, SCALE_B, SPLIT, LW, Epi>;
  static DeviceFlags configured;     // per device: the attribute belongs to the device's copy of the kernel
  if (!configured(h)) {
    GCCNMF_CHECK_CUDA(h, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    configured(h) = true;
  }
  dim3 grid((args.N + BN - 1) / BN, args.m_tiles + (tail_rows > 0 ? 1 : 0), splits);
  GCCNMF_LAUNCH(h, kernel, grid, LW * 32 + 32, S::kTotal, stream, args, epi);
  return 0;
}

template <int BN, bool SCALE_B, class Epi>
int launch_gemm(gccnmf_handle* h, const GemmArgs& args, int tail_rows, int splits, const Epi& epi, void* stream) {
  if (h->nmf_split_bf16) return launch_gemm_split<BN, SCALE_B, umma::kSplitBF16>(h, args, tail_rows, splits, epi, stream);
  return launch_gemm_split<BN, SCALE_B, umma::kSplitTF32>(h, args, tail_rows, splits, epi, stream);
}

int tile_width(const gccnmf_handle* h, int m_tiles, int N, int splits) {
  const int tiles128 = m_tiles * ((N + 127) / 128);
  return (splits == 1 && N >= 256 && tiles128 > h->sm_count + h->sm_count / 4) ? 256 : 128;
}

// D = A.B^T over 128-row tiles on the tensor cores and a short row tail on SIMT CTAs of the same launch.
template <bool SCALE_B, class Epi>
int tc_gemm(gccnmf_handle* h, GemmArgs args, int splits, const Epi& epi, void* stream, int force_tile_n = 0) {
  const int total_kb = (args.Kc + umma::kBK - 1) / umma::kBK;
  args.kblocks_per_split = (total_kb + splits - 1) / splits;
  const int tail = args.M % umma::kBM;
  const bool simt_tail = tail != 0 && tail <= kTailRowsMax && args.M > umma::kBM;
  args.m_tiles = simt_tail ? args.M / umma::kBM : (args.M + umma::kBM - 1) / umma::kBM;
  const int bn = force_tile_n ? force_tile_n : tile_width(h, args.m_tiles, args.N, splits);
  if (bn == 256) return launch_gemm<256, SCALE_B>(h, args, simt_tail ? tail : 0, splits, epi, stream);
  return launch_gemm<128, SCALE_B>(h, args, simt_tail ? tail : 0, splits, epi, stream);
}

struct TcWorkspace {
  float *HT, *WT, *VT, *R, *RT, *Hp, *partial, *colsum, *sumsq_part, *rowsum_part, *norms;
  int row_blocks;
  int64_t Fp, T2p;
  int splits, rowsum_slots;
  bool ok;
};

int tc_pick_splits(int tiles, int total_kblocks, int sm_count) {
  int s = 1;
  while (s < kMaxSplits && tiles * (s * 2) <= sm_count && total_kblocks / (s * 2) >= 8) s *= 2;
  return s;
}

int m_tiles_of(int M) {
  const int tail = M % umma::kBM;
  return (tail != 0 && tail <= kTailRowsMax && M > umma::kBM) ? M / umma::kBM : (M + umma::kBM - 1) / umma::kBM;
}

TcWorkspace tc_carve(const gccnmf_handle* h, void* ws, size_t bytes, int F, int T2, int K) {
  WorkspaceCarver c(ws, bytes);
  TcWorkspace w;
  w.Fp = (F + 3) & ~3;
  w.T2p = (T2 + 3) & ~3;
  w.splits = tc_pick_splits(m_tiles_of(F) * ((K + 127) / 128), (T2 + umma::kBK - 1) / umma::kBK, h->sm_count);
  const int bn = tile_width(h, m_tiles_of(K), T2, 1);          // G2's tile width decides the number of row-sum slots
  w.rowsum_slots = ((T2 + bn - 1) / bn) * (kNmfWorkerWarps / 4);
  w.HT = c.take<float>((size_t)T2 * K);
  w.WT = c.take<float>((size_t)K * w.Fp);
  w.VT = c.take<float>((size_t)T2 * w.Fp);
  w.R = c.take<float>((size_t)F * w.T2p);
  w.RT = c.take<float>((size_t)T2 * w.Fp);
  w.Hp = c.take<float>((size_t)K * w.T2p);
  w.partial = c.take<float>((size_t)kMaxSplits * F * K);
  w.row_blocks = (F + kApplyTile - 1) / kApplyTile;
  w.colsum = c.take<float>((size_t)w.row_blocks * K);
  w.sumsq_part = c.take<float>((size_t)w.row_blocks * K);
  w.rowsum_part = c.take<float>((size_t)((kNmfWorkerWarps / 4) * ((T2 + 127) / 128)) * K);
  w.norms = c.take<float>(K);
  w.ok = c.ok();
  return w;
}

size_t tc_workspace_bytes(int F, int T2, int K) {
  const size_t Fp = (F + 3) & ~3, T2p = (T2 + 3) & ~3;
  size_t n = 0;
  auto add = [&](size_t count) { n = align_up(n, 256) + count * sizeof(float); };
  add((size_t)T2 * K); add((size_t)K * Fp); add((size_t)T2 * Fp); add((size_t)F * T2p); add((size_t)T2 * Fp); add((size_t)K * T2p);
  add((size_t)kMaxSplits * F * K); add((size_t)((F + 31) / 32) * K); add((size_t)((F + 31) / 32) * K);
  add((size_t)((kNmfWorkerWarps / 4) * ((T2 + 127) / 128)) * K); add(K);
  return align_up(n, 256);
}

#define TC_CARVE_OR_FAIL(w)                                                                                         \
  TcWorkspace w = tc_carve(h, workspace, workspace_bytes, F, T2, K);                                                \
  if (!w.ok) return gccnmf_fail(h, GCCNMF_ERR_WORKSPACE, "klnmf (tensor-core path) workspace too small: need %zu bytes", tc_workspace_bytes(F, T2, K))

}  // namespace

// Whether the tensor-core path supports this problem (else the SIMT path in klnmf.cu is used).
bool gccnmf_klnmf_tc_supported(int F, int T2, int K) { return K % 4 == 0 && F >= 128 && T2 >= 128 && K >= 32; }
size_t gccnmf_klnmf_tc_workspace_bytes(int F, int T2, int K) { return tc_workspace_bytes(F, T2, K); }

// Transposes / pads the caller's V, W, H into the k-contiguous operand set.
int gccnmf_klnmf_tc_prepare(gccnmf_handle* h, const float* V, int F, int T2, const float* W, const float* H, int K,
                            void* workspace, size_t workspace_bytes, bool need_vt, bool need_wt, bool need_ht, void* stream) {
  TC_CARVE_OR_FAIL(w);
  dim3 block(32, 8);
  if (need_vt) {
    GCCNMF_LAUNCH(h, transpose_pad_kernel, dim3((T2 + 31) / 32, (F + 31) / 32 + 1), block, 0, stream, V, F, T2, (int64_t)T2, w.VT, w.Fp);
    GCCNMF_CHECK_CUDA(h, cudaMemsetAsync(w.RT, 0, (size_t)T2 * w.Fp * sizeof(float), (cudaStream_t)stream));
  }
  if (need_wt) GCCNMF_LAUNCH(h, transpose_pad_kernel, dim3((K + 31) / 32, (F + 31) / 32 + 1), block, 0, stream, W, F, K, (int64_t)K, w.WT, w.Fp);
  if (need_ht) {
    GCCNMF_LAUNCH(h, transpose_pad_kernel, dim3((T2 + 31) / 32, (K + 31) / 32), block, 0, stream, H, K, T2, (int64_t)T2, w.HT, (int64_t)K);
    GCCNMF_CHECK_CUDA(h, cudaMemsetAsync(w.R, 0, (size_t)F * w.T2p * sizeof(float), (cudaStream_t)stream));
    const int64_t n = (int64_t)K * w.T2p;
    GCCNMF_LAUNCH(h, tc_copy_pad_kernel, (unsigned)((n + 255) / 256), 256, 0, stream, H, (int64_t)T2, T2, w.Hp, w.T2p, K);
  }
  return 0;
}

// :76 (preceded by the pending :81 when pending_norms): H = (n*H) * (W^T (V / (W (n*H)))) / (colsum(W) + alpha + eps).
// Requires prepare(); keeps H and HT in sync and leaves per-row partial sums of the new H in the workspace.
int gccnmf_klnmf_tc_update_H(gccnmf_handle* h, const float* V, int F, int T2, const float* W, float* H, int K, float alpha, float eps,
                             void* workspace, size_t workspace_bytes, int colsum_state, bool pending_norms, void* stream) {
  // colsum_state: 0 = compute colsum(W) now; 1 = reuse the one computed before (fixed dictionary); 2 = per-row-block partials left by the W update
  TC_CARVE_OR_FAIL(w);
  (void)V;
  const float* pending = pending_norms ? w.norms : nullptr;
  {  // G1: RT = VT / (W.(n*H))
    GemmArgs a{W, w.HT, F, T2, K, (int64_t)K, (int64_t)K, 0, 0, pending, nullptr, 0};
    EpiRatioT e{w.VT, w.RT, w.Fp, F, T2};
    const int st = pending ? tc_gemm<true>(h, a, 1, e, stream) : tc_gemm<false>(h, a, 1, e, stream);
    if (st) return st;
  }
  if (colsum_state == 0) GCCNMF_LAUNCH(h, tc_colsum_kernel, (K + 127) / 128, 128, 0, stream, W, F, K, w.colsum);
  GCCNMF_CHECK_CUDA(h, cudaMemsetAsync(w.rowsum_part, 0, (size_t)w.rowsum_slots * K * sizeof(float), (cudaStream_t)stream));
  {  // G2: H, HT = (n*H) * (WT.RT^T) / denom
    GemmArgs a{w.WT, w.RT, K, T2, F, w.Fp, w.Fp, 0, 0, nullptr, nullptr, 0};
    EpiUpdateHBoth e{w.Hp, w.HT, w.colsum, pending, w.rowsum_part, alpha, eps, w.T2p, (int64_t)K, K, T2, colsum_state == 2 ? w.row_blocks : 1};
    (void)H;
    const int bn = tile_width(h, m_tiles_of(K), T2, 1);
    if (int st = tc_gemm<false>(h, a, 1, e, stream, bn)) return st;
  }
  return 0;
}

// :77 numerator: partial[z] = (V / (W H)) . H^T over the frame range of split z (row sums of H come from update_H).
int gccnmf_klnmf_tc_partial_W(gccnmf_handle* h, const float* V, int F, int T2, const float* W, const float* H, int K,
                              void* workspace, size_t workspace_bytes, bool have_rowsum, void* stream) {
  TC_CARVE_OR_FAIL(w);
  {  // G3: R = V / (W.H)
    GemmArgs a{W, w.HT, F, T2, K, (int64_t)K, (int64_t)K, 0, 0, nullptr, nullptr, 0};
    EpiRatioRow e{V, w.R, (int64_t)T2, w.T2p, F, T2};
    if (int st = tc_gemm<false>(h, a, 1, e, stream)) return st;
  }
  if (!have_rowsum) {   // stateless building block: H may not be the one update_H just wrote
    GCCNMF_CHECK_CUDA(h, cudaMemsetAsync(w.rowsum_part, 0, (size_t)w.rowsum_slots * K * sizeof(float), (cudaStream_t)stream));
    GCCNMF_LAUNCH(h, tc_rowsum_kernel, K, 256, 0, stream, w.Hp, T2, w.T2p, w.rowsum_part);
  }
  {  // G4: partial[z] = R.H^T
    (void)H;
    GemmArgs a{w.R, w.Hp, F, K, T2, w.T2p, w.T2p, 0, 0, nullptr, nullptr, 0};
    EpiStoreRowMajor e{w.partial, (int64_t)K, F, K, (int64_t)F * K};
    if (int st = tc_gemm<false>(h, a, w.splits, e, stream)) return st;
  }
  return 0;
}

// :77-:80; the H rescale of :81 stays pending (applied by the next update_H, or by finish).  Numerator and row
// sums come from `numer` (F*K + K floats, all-reduced across ranks) when given, else from this rank's partials.
int gccnmf_klnmf_tc_apply_W(gccnmf_handle* h, int F, int T2, float* W, int K, const float* numer, bool numer_is_multicast,
                            void* workspace, size_t workspace_bytes, void* stream) {
  TC_CARVE_OR_FAIL(w);
  const float* partial = numer ? numer : w.partial;
  const float* rowsum = numer ? numer + (int64_t)F * K : w.rowsum_part;
  const dim3 grid((K + kApplyTile - 1) / kApplyTile, w.row_blocks), block(kApplyTile, 8);
  if (numer_is_multicast) {
    GCCNMF_LAUNCH(h, tc_apply_w1_kernel<true>, grid, block, 0, stream, W, partial, 1, rowsum, 1, F, K, w.sumsq_part);
  } else {
    GCCNMF_LAUNCH(h, tc_apply_w1_kernel<false>, grid, block, 0, stream, W, partial, numer ? 1 : w.splits, rowsum,
                  numer ? 1 : w.rowsum_slots, F, K, w.sumsq_part);
  }
  GCCNMF_LAUNCH(h, tc_apply_w2_kernel, grid, block, 0, stream, W, w.WT, w.Fp, w.sumsq_part, w.row_blocks, F, K, w.norms, w.colsum);
  return 0;
}

// Writes the caller's H from the working copy, applying a pending H *= norms (:81) when there is one.
int gccnmf_klnmf_tc_finish(gccnmf_handle* h, int F, int T2, float* H, int K, bool pending_norms, void* workspace,
                           size_t workspace_bytes, void* stream) {
  TC_CARVE_OR_FAIL(w);
  const int64_t n = (int64_t)K * T2;
  GCCNMF_LAUNCH(h, tc_finish_h_kernel, (unsigned)((n + 255) / 256), 256, 0, stream, w.Hp, w.T2p, pending_norms ? w.norms : nullptr, H, K, T2);
  return 0;
}

int gccnmf_klnmf_tc_pack_numer(gccnmf_handle* h, int F, int T2, int K, float* numer, void* workspace, size_t workspace_bytes, void* stream) {
  TC_CARVE_OR_FAIL(w);
  const int64_t n = (int64_t)F * K;
  GCCNMF_LAUNCH(h, tc_pack_numer_kernel, (unsigned)((n + K + 255) / 256), 256, 0, stream, w.partial, w.splits, n, w.rowsum_part,
                w.rowsum_slots, K, numer);
  return 0;
}

extern "C" {

// D (M, N) row-major (ldd) = A (M, Kc; lda) . B (N, Kc; ldb)^T with 3xTF32 error compensation.
// lda, ldb multiples of 4 with zero padding up to round_up(Kc, 4); tile_n in {128, 256}.
int gccnmf_gemm_tn_3xtf32_timed(gccnmf_handle* h, const float* A, int64_t lda, const float* B, int64_t ldb, float* D, int64_t ldd,
                                int M, int N, int Kc, int tile_n, unsigned long long* timing, void* stream);

int gccnmf_gemm_tn_3xtf32(gccnmf_handle* h, const float* A, int64_t lda, const float* B, int64_t ldb, float* D,
                          int64_t ldd, int M, int N, int Kc, int tile_n, void* stream) {
  return gccnmf_gemm_tn_3xtf32_timed(h, A, lda, B, ldb, D, ldd, M, N, Kc, tile_n, nullptr, stream);
}

// Diagnostics: same product; `timing` (device, 6 x number of CTAs uint64, or NULL) receives per-CTA clock64 stamps:
// [0] kernel start, [1] first stage full, [2] last MMA issued, [3] loaders finished, [4] accumulator complete, [5] epilogue end.
int gccnmf_gemm_tn_3xtf32_timed(gccnmf_handle* h, const float* A, int64_t lda, const float* B, int64_t ldb, float* D, int64_t ldd,
                                int M, int N, int Kc, int tile_n, unsigned long long* timing, void* stream) {
  GCCNMF_ENTER(h);
  GCCNMF_REQUIRE(h, A && B && D && M > 0 && N > 0 && Kc > 0, "gemm_tn_3xtf32: bad arguments");
  GCCNMF_REQUIRE(h, lda % 4 == 0 && ldb % 4 == 0 && lda >= ((Kc + 3) & ~3) && ldb >= ((Kc + 3) & ~3),
                 "gemm_tn_3xtf32: leading dimensions must be multiples of 4 covering round_up(Kc, 4)");
  GCCNMF_REQUIRE(h, (reinterpret_cast<uintptr_t>(A) % 16 == 0) && (reinterpret_cast<uintptr_t>(B) % 16 == 0),
                 "gemm_tn_3xtf32: operands must be 16-byte aligned");
  if (tile_n != 128 && tile_n != 256) return gccnmf_fail(h, GCCNMF_ERR_UNSUPPORTED, "gemm_tn_3xtf32: tile_n must be 128 or 256");
  GemmArgs args{A, B, M, N, Kc, lda, ldb, 0, 0, nullptr, timing, 0};
  EpiStoreRowMajor epi{D, ldd, M, N, 0};
  return tc_gemm<false>(h, args, 1, epi, stream, tile_n);
}

}  // extern "C"
