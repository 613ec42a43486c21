// The real-time block path as ONE stream-ordered unit with no host synchronisation inside, capturable in a CUDA graph:
//   reference  gccNMF/realtime/gccNMFProcessor.py:201-231 (GCCNMFProcessor.processFrames), :245-270 (the Theano graph it runs),
//              gccNMF/realtime/utils.py:99-116 (OverlapAddProcessor.processFrames: 8-block input / output rings around it),
//              notebooks/onlineSpeechEnhancement.ipynb:433-438 (per-frame coefficient inference, the numInferenceIterations > 0 branch).
//
// One audio block = five small kernels (+ two per inference iteration), every intermediate in the caller-owned state buffer:
//   A  rt_analysis    per frame: window . samples (from the input ring, or from the caller's windowed frames) -> one complex FFT for
//                     the stereo pair (left + i right, Hermitian split) -> X (2, F, nT) complex64, PHAT coherence, realGCC
//                     G[t][d][f] = Re(coh[f] E[f][d]) (float32, as the Theano float32 graph), gccPHAT[d][t] = nanmean_f G, |X|
//   I  rt_inf_ratio / rt_inf_update   (inference_iterations > 0)  H-only KL updates with the fixed dictionary, all 2 nT columns
//   B  rt_atoms       gccNMF[d][k] = sum_f G[t][d][f] W[f][k] (float32) -> argmax over TDOA per atom -> atom mask (boxcar / window, float64)
//   C  rt_filter      tfMask[f][t] = (W . mask) / rowsum(W)   (or (W . (H mask)) / (W . H) per channel with inference); Y = tfMask X
//   D  rt_synthesis   per frame: Hermitian rebuild, inverse FFT of the stereo pair, x synthesis window
//   E  rt_ola_emit    overlap-add of the nT frames into the output ring in frame order, emits block [-3B, -2B), pushes the block into
//                     the input ring; one extra CTA keeps the GCC-PHAT history ring and the sliding-window localisation
//                     (argmax of the nanmean over the last `localization_window` columns -> target TDOA index of the NEXT block).
// The rings are circular in place (the reference shifts 8 blocks of memory per call); positions derive from a device-side block
// counter, so a captured graph replays unchanged block after block.
#include <cmath>

#include "common.cuh"
#include "fft.cuh"

namespace {

constexpr int kRtMaxN = 2048;
constexpr int kRtMaxFrames = 8;          // frames per block the filter kernel keeps in registers
constexpr int kRtMaxD = 128;
constexpr int kRtAtomsPerCta = 16;
constexpr int kRtRingBlocks = 8;         // utils.py:85 numBlocksPerBuffer

struct RtDev {                           // device-resident parameters + loop-carried state (first bytes of the state buffer)
  float target, eps, beta, noise_floor;  // gccNMFProcessor.py:196-199 (Theano shared scalars)
  int mode, separation, localization, loc_window;
  int block_counter;                     // blocks completed (incremented by the synthesis kernel)
  int hist_index;                        // write position of the GCC-PHAT history ring (utils.py:45-59)
};

struct RtLayout {                        // carve of the caller-owned state buffer (pure function of the configuration)
  RtDev* dev;
  float *in_ring, *out_ring, *win_a, *win_s, *W, *WT, *recV, *colsumW, *G, *gccphat, *Vabs, *H, *H0, *R, *frames, *tw32;
  double *tw64, *hist, *hmask;
  float2 *ET, *X, *Y;
  int32_t* argmax;
  int F, Fp, L;
  size_t bytes;
  bool ok;
};

RtLayout rt_carve(const gccnmf_rt_config& c, void* state, size_t state_bytes) {
  RtLayout l{};
  const int N = c.window_size, nT = c.windows_per_block, K = c.num_atoms, D = c.num_tdoas;
  l.F = N / 2 + 1;
  l.Fp = (l.F + 3) & ~3;
  l.L = kRtRingBlocks * c.block_size;
  WorkspaceCarver w(state ? state : reinterpret_cast<void*>(256), state ? state_bytes : ~size_t(0) >> 1);
  l.dev = w.take<RtDev>(1);
  l.tw64 = w.take<double>(N);
  l.hist = w.take<double>((size_t)D * c.history_length);
  l.hmask = w.take<double>((size_t)K * nT);
  l.tw32 = w.take<float>(N);
  l.in_ring = w.take<float>((size_t)2 * l.L);
  l.out_ring = w.take<float>((size_t)2 * l.L);
  l.win_a = w.take<float>(N);
  l.win_s = w.take<float>(N);
  l.W = w.take<float>((size_t)l.F * K);
  l.WT = w.take<float>((size_t)K * l.Fp);
  l.recV = w.take<float>(l.F);
  l.colsumW = w.take<float>(K);
  l.G = w.take<float>((size_t)nT * D * l.Fp);
  l.gccphat = w.take<float>((size_t)D * nT);
  l.Vabs = w.take<float>((size_t)l.F * 2 * nT);
  l.H = w.take<float>((size_t)K * 2 * nT);
  l.H0 = w.take<float>((size_t)K * 2);
  l.R = w.take<float>((size_t)l.F * 2 * nT);
  l.frames = w.take<float>((size_t)2 * nT * N);
  l.ET = w.take<float2>((size_t)D * l.Fp);
  l.X = w.take<float2>((size_t)2 * l.F * nT);
  l.Y = w.take<float2>((size_t)2 * l.F * nT);
  l.argmax = w.take<int32_t>((size_t)K * nT);
  l.bytes = align_up(w.used, 256);
  l.ok = state != nullptr && w.ok();
  return l;
}

int rt_check(gccnmf_handle* h, const gccnmf_rt_config* c) {
  GCCNMF_REQUIRE(h, c != nullptr, "rt: NULL configuration");
  const int N = c->window_size;
  GCCNMF_REQUIRE(h, N >= 64 && N <= kRtMaxN && (N & (N - 1)) == 0, "rt: window_size must be a power of two in [64, %d] (got %d)", kRtMaxN, N);
  GCCNMF_REQUIRE(h, c->hop_size >= 1 && c->block_size >= 1, "rt: hop_size and block_size must be positive");
  GCCNMF_REQUIRE(h, c->windows_per_block >= 1 && c->windows_per_block <= kRtMaxFrames, "rt: windows_per_block must be in [1, %d] (got %d)", kRtMaxFrames,
                 c->windows_per_block);
  GCCNMF_REQUIRE(h, c->num_atoms >= 1 && c->num_tdoas >= 1, "rt: num_atoms and num_tdoas must be positive");
  if (c->num_tdoas > kRtMaxD) return gccnmf_fail(h, GCCNMF_ERR_UNSUPPORTED, "rt: num_tdoas %d > %d", c->num_tdoas, kRtMaxD);
  GCCNMF_REQUIRE(h, c->history_length >= 1 && c->inference_iterations >= 0, "rt: history_length must be positive, inference_iterations >= 0");
  // the windows of one block must lie inside the 8-block rings (utils.py:107)
  GCCNMF_REQUIRE(h, N + (c->windows_per_block - 1) * c->hop_size <= kRtRingBlocks * c->block_size && 3 * c->block_size <= kRtRingBlocks * c->block_size,
                 "rt: window_size + (windows_per_block - 1) hop_size exceeds the 8-block ring");
  return 0;
}

// ---------------------------------------------------------------------------------------------- init-time kernels
__global__ void rt_init_dictionary_kernel(const float* __restrict__ W, int F, int Fp, int K, float* __restrict__ WT, float* __restrict__ recV,
                                          float* __restrict__ colsumW) {
  // one thread per atom: column sum in row order (numpy.sum(W, axis=0)) + transposed copy; one thread per bin: row sum
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < K) {
    float s = 0.f;
    for (int f = 0; f < Fp; ++f) {
      const float w = f < F ? W[(int64_t)f * K + i] : 0.f;
      WT[(int64_t)i * Fp + f] = w;
      if (f < F) s += w;
    }
    colsumW[i] = s;
  }
  if (i < F) {
    double s = 0.0;
    for (int k = 0; k < K; ++k) s += (double)W[(int64_t)i * K + k];
    recV[i] = (float)s;                  // tensor.sum(W, axis=-1) in float32 (gccNMFProcessor.py:268)
  }
}

__global__ void rt_init_steering_kernel(const float2* __restrict__ E, int F, int Fp, int D, float2* __restrict__ ET) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D * Fp) return;
  const int d = i / Fp, f = i - d * Fp;
  ET[i] = f < F ? E[(int64_t)f * D + d] : float2{0.f, 0.f};
}

__global__ void rt_set_params_kernel(RtDev* dev, float target, float eps, float beta, float noise_floor, int mode, int separation, int localization,
                                     int loc_window, int set_target) {
  if (set_target) dev->target = target;
  dev->eps = eps; dev->beta = beta; dev->noise_floor = noise_floor;
  dev->mode = mode; dev->separation = separation; dev->localization = localization; dev->loc_window = loc_window;
}

// ---------------------------------------------------------------------------------------------- numerics shared with gcc.cu
// numpy / Theano complex64 arithmetic of  X0 * conj(X1) / |X0| / |X1|  (gccNMFProcessor.py:253; runGCCNMF.py:44)
__device__ __forceinline__ float2 rt_coherence(float2 a, float2 b) {
  float re = __fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));
  float im = __fsub_rn(__fmul_rn(a.y, b.x), __fmul_rn(a.x, b.y));
  const float ma = (float)sqrt((double)a.x * a.x + (double)a.y * a.y);
  const float mb = (float)sqrt((double)b.x * b.x + (double)b.y * b.y);
  const float ia = 1.0f / ma, ib = 1.0f / mb;
  re = __fmul_rn(re, ia); im = __fmul_rn(im, ia);
  re = __fmul_rn(re, ib); im = __fmul_rn(im, ib);
  return float2{re, im};
}
// numpy.argmax ordering: NaN is a maximum, the first occurrence wins
__device__ __forceinline__ bool rt_better(float v, int i, float bv, int bi) {
  const bool vn = v != v, bn = bv != bv;
  if (vn || bn) return vn && (!bn || i < bi);
  return v > bv || (v == bv && i < bi);
}
__device__ __forceinline__ bool rt_better64(double v, int i, double bv, int bi) {
  const bool vn = v != v, bn = bv != bv;
  if (vn || bn) return vn && (!bn || i < bi);
  return v > bv || (v == bv && i < bi);
}

// Logical ring position p of the reference's 8-block buffers (0 = oldest sample, L - 1 = newest) AFTER the shift of call number
// `call` (1-based) -> physical index: the sample stream position is call * B - L + p.
__device__ __forceinline__ int rt_ring_index(int p, int call, int B, int L) {
  long long a = (long long)call * B - L + p;
  a %= L;
  return (int)(a < 0 ? a + L : a);
}

// ---------------------------------------------------------------------------------------------- A: analysis (one CTA per frame)
__global__ void __launch_bounds__(kFftThreads)
rt_analysis_kernel(const RtDev* __restrict__ dev, const float* __restrict__ windowed,   // (2, N, nT) or NULL: read the ring + the new block
                   const float* __restrict__ in_block, const float* __restrict__ in_ring, int B, int L, int hop, int nT,
                   const float* __restrict__ win_a, const double2* __restrict__ tw, int N, int log2n, const float2* __restrict__ ET, int D, int Fp,
                   float2* __restrict__ X, float* __restrict__ G, float* __restrict__ gccphat, float* __restrict__ Vabs,
                   const float* __restrict__ H0, float* __restrict__ H, int K, int inference) {
  __shared__ double2 fft[kRtMaxN];
  __shared__ float2 coh[kRtMaxN / 2 + 1];
  const int t = blockIdx.x, F = N / 2 + 1;
  const int call = dev->block_counter + 1;
  const int w0 = L - N - (nT - 1 - t) * hop;         // utils.py:107 windowIndexes[t]
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    float l, r;
    if (windowed) {
      l = windowed[((int64_t)0 * N + i) * nT + t];
      r = windowed[((int64_t)1 * N + i) * nT + t];
    } else {
      const int p = w0 + i;
      if (p >= L - B) {                               // the block being pushed in by this call
        l = in_block[p - (L - B)];
        r = in_block[B + p - (L - B)];
      } else {
        const int q = rt_ring_index(p, call, B, L);
        l = in_ring[q];
        r = in_ring[L + q];
      }
    }
    const float w = win_a[i];
    fft[bitrev(i, log2n)] = double2{(double)__fmul_rn(l, w), (double)__fmul_rn(r, w)};   // float32 product (:202), double transform
  }
  __syncthreads();
  fft_inplace<double2, double>(fft, tw, N, log2n, false);
  for (int k = threadIdx.x; k < F; k += blockDim.x) {
    const double2 a = fft[k], b = fft[(N - k) & (N - 1)];
    const float2 xl = float2{(float)(0.5 * (a.x + b.x)), (float)(0.5 * (a.y - b.y))};
    const float2 xr = float2{(float)(0.5 * (a.y + b.y)), (float)(0.5 * (b.x - a.x))};
    X[((int64_t)0 * F + k) * nT + t] = xl;
    X[((int64_t)1 * F + k) * nT + t] = xr;
    coh[k] = rt_coherence(xl, xr);
    if (inference) {                                  // abs(stereoSTFTFrame).T (onlineSpeechEnhancement.ipynb:433): column 2 t + channel
      Vabs[(int64_t)k * (2 * nT) + 2 * t] = (float)sqrt((double)xl.x * xl.x + (double)xl.y * xl.y);
      Vabs[(int64_t)k * (2 * nT) + 2 * t + 1] = (float)sqrt((double)xr.x * xr.x + (double)xr.y * xr.y);
    }
  }
  if (inference)                                      // every frame starts from the same seeded H0 (the notebook re-seeds per call)
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
      H[(int64_t)k * (2 * nT) + 2 * t] = H0[2 * k];
      H[(int64_t)k * (2 * nT) + 2 * t + 1] = H0[2 * k + 1];
    }
  __syncthreads();
  // realGCC[f][t][d] = Re(coh[f] * E[f][d]) in complex64 arithmetic (:254), stored [t][d][f]
  float* Gt = G + (int64_t)t * D * Fp;
  for (int i = threadIdx.x; i < D * Fp; i += blockDim.x) {
    const int d = i / Fp, f = i - d * Fp;
    float v = 0.f;
    if (f < F) {
      const float2 c = coh[f], e = ET[i];
      v = __fsub_rn(__fmul_rn(c.x, e.x), __fmul_rn(c.y, e.y));
    }
    Gt[i] = v;
  }
  __syncthreads();
  // gccPHAT[d] = nanmean over f (:214)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int d = warp; d < D; d += kFftThreads / 32) {
    double s = 0.0;
    int n = 0;
    for (int f = lane; f < F; f += 32) {
      const float v = Gt[(int64_t)d * Fp + f];
      if (v == v) { s += (double)v; ++n; }
    }
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); n += __shfl_xor_sync(0xffffffffu, n, o); }
    if (lane == 0) gccphat[(int64_t)d * nT + t] = n > 0 ? (float)(s / (double)n) : __int_as_float(0x7fc00000);
  }
}

// ---------------------------------------------------------------------------------------------- I: coefficient inference
// R[f][j] = V[f][j] / sum_k W[f][k] H[k][j]   (gccNMFFunctions.py:76, V / dot(W, H)); one warp per bin, J = 2 nT columns
template <int J>
__global__ void __launch_bounds__(256)
rt_inf_ratio_kernel(const float* __restrict__ W, const float* __restrict__ H, const float* __restrict__ V, int F, int K, float* __restrict__ R) {
  const int f = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (f >= F) return;
  float acc[J];
#pragma unroll
  for (int j = 0; j < J; ++j) acc[j] = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float w = W[(int64_t)f * K + k];
#pragma unroll
    for (int j = 0; j < J; ++j) acc[j] = fmaf(w, H[(int64_t)k * J + j], acc[j]);
  }
#pragma unroll
  for (int j = 0; j < J; ++j) {
    float s = acc[j];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) R[(int64_t)f * J + j] = V[(int64_t)f * J + j] / s;
  }
}
// H[k][j] *= (sum_f W[f][k] R[f][j]) / (colsum(W)[k] + alpha + eps)   (:76); one warp per atom over the transposed dictionary
template <int J>
__global__ void __launch_bounds__(256)
rt_inf_update_kernel(const float* __restrict__ WT, int Fp, const float* __restrict__ R, int F, int K, const float* __restrict__ colsumW, float alpha,
                     float eps, float* __restrict__ H) {
  const int k = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (k >= K) return;
  float acc[J];
#pragma unroll
  for (int j = 0; j < J; ++j) acc[j] = 0.f;
  for (int f = lane; f < F; f += 32) {
    const float w = WT[(int64_t)k * Fp + f];
#pragma unroll
    for (int j = 0; j < J; ++j) acc[j] = fmaf(w, R[(int64_t)f * J + j], acc[j]);
  }
  const float denom = (colsumW[k] + alpha) + eps;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    float s = acc[j];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) H[(int64_t)k * J + j] = H[(int64_t)k * J + j] * (s / denom);
  }
}

// ---------------------------------------------------------------------------------------------- B: per-atom TDOA argmax + atom mask
// grid (K / 16, nT); thread = (TDOA group dg = tid / 16, atom = tid % 16); TDOAs d = dg + 16 j.
template <int DJ>        // TDOAs per thread: D <= 16 DJ
__global__ void __launch_bounds__(256)
rt_atoms_kernel(const RtDev* __restrict__ dev, const float* __restrict__ G, const float* __restrict__ W, int F, int Fp, int K, int D, int nT,
                int32_t* __restrict__ argmax, double* __restrict__ hmask) {
  __shared__ float Gs[16 * DJ][33];
  __shared__ float Ws[32][kRtAtomsPerCta];
  __shared__ float vals[16 * DJ][kRtAtomsPerCta + 1];
  const int t = blockIdx.y, k0 = blockIdx.x * kRtAtomsPerCta;
  const int atom = threadIdx.x & 15, dg = threadIdx.x >> 4;
  const float* Gt = G + (int64_t)t * D * Fp;
  float acc[DJ];
#pragma unroll
  for (int j = 0; j < DJ; ++j) acc[j] = 0.f;
  for (int f0 = 0; f0 < F; f0 += 32) {
    for (int i = threadIdx.x; i < 16 * DJ * 32; i += 256) {
      const int d = i >> 5, ff = i & 31;
      Gs[d][ff] = (d < D && f0 + ff < F) ? Gt[(int64_t)d * Fp + f0 + ff] : 0.f;
    }
    for (int i = threadIdx.x; i < 32 * kRtAtomsPerCta; i += 256) {
      const int ff = i >> 4, a = i & 15;
      Ws[ff][a] = (f0 + ff < F && k0 + a < K) ? W[(int64_t)(f0 + ff) * K + k0 + a] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int ff = 0; ff < 32; ++ff) {
      const float w = Ws[ff][atom];
#pragma unroll
      for (int j = 0; j < DJ; ++j) acc[j] = fmaf(Gs[dg + 16 * j][ff], w, acc[j]);     // tensor.dot(realGCC.T, W) in float32 (:259)
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < DJ; ++j) vals[dg + 16 * j][atom] = acc[j];
  __syncthreads();
  if (threadIdx.x < kRtAtomsPerCta && k0 + threadIdx.x < K) {
    const int a = threadIdx.x;
    float bv = vals[0][a];
    int bi = 0;
    for (int d = 1; d < D; ++d)
      if (rt_better(vals[d][a], d, bv, bi)) { bv = vals[d][a]; bi = d; }
    const int64_t o = (int64_t)(k0 + a) * nT + t;
    argmax[o] = bi;
    // int64 - float32 promotes to float64 in Theano and numpy alike: the mask arithmetic is float64 (:263, :265)
    const double dist = fabs((double)bi - (double)dev->target);
    double m;
    if (dev->mode == 0) m = dist < (double)dev->eps ? 1.0 : 0.0;
    else m = exp(-pow(dist / (double)dev->eps, (double)dev->beta)) / (double)(1.0f + dev->noise_floor) + (double)dev->noise_floor;
    hmask[o] = m;
  }
}

// ---------------------------------------------------------------------------------------------- C: time-frequency mask, one warp per bin
__global__ void __launch_bounds__(256)
rt_filter_kernel(const RtDev* __restrict__ dev, const float* __restrict__ W, const double* __restrict__ hmask, const float* __restrict__ recV,
                 const float* __restrict__ H, int inference, const float2* __restrict__ X, int F, int K, int nT, float2* __restrict__ Y) {
  const int f = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (f >= F) return;
  if (!dev->separation) {                            // :211 outputSpectrogram = complexMixtureSpectrogram.copy()
    for (int i = lane; i < 2 * nT; i += 32) {
      const int c = i / nT, t = i - c * nT;
      Y[((int64_t)c * F + f) * nT + t] = X[((int64_t)c * F + f) * nT + t];
    }
    return;
  }
  double num[2][kRtMaxFrames], den[2][kRtMaxFrames];
#pragma unroll
  for (int t = 0; t < kRtMaxFrames; ++t) { num[0][t] = num[1][t] = den[0][t] = den[1][t] = 0.0; }
  for (int k = lane; k < K; k += 32) {
    const double w = (double)W[(int64_t)f * K + k];
#pragma unroll
    for (int t = 0; t < kRtMaxFrames; ++t) {
      if (t < nT) {
        const double m = hmask[(int64_t)k * nT + t];
        if (inference) {                             // sourceEstimate = W . (H * mask), recV = W . H  (ipynb:435-437)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const double hv = (double)H[(int64_t)k * (2 * nT) + 2 * t + c];
            num[c][t] += w * (hv * m);
            den[c][t] += w * hv;
          }
        } else {
          num[0][t] += w * m;                        // tensor.dot(W, HMask) (:267)
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < kRtMaxFrames; ++t) {
    if (t < nT) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (c == 1 && !inference) break;
        double a = num[c][t], b = den[c][t];
        for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
        num[c][t] = a; den[c][t] = b;
      }
    }
  }
  if (lane < 2 * nT) {
    const int c = lane / nT, t = lane - c * nT;
    double tf = 0.0;
#pragma unroll
    for (int tt = 0; tt < kRtMaxFrames; ++tt)
      if (tt == t) tf = inference ? num[c & 1][tt] / den[c & 1][tt] : num[0][tt] / (double)recV[f];
    const float2 x = X[((int64_t)c * F + f) * nT + t];
    Y[((int64_t)c * F + f) * nT + t] = float2{(float)(tf * (double)x.x), (float)(tf * (double)x.y)};   // inputMask * spectrogram (:209)
  }
}

// ---------------------------------------------------------------------------------------------- D: synthesis (one CTA per frame)
__global__ void __launch_bounds__(kFftThreads)
rt_synthesis_kernel(RtDev* __restrict__ dev, const float2* __restrict__ Y, const float2* __restrict__ tw, int N, int log2n, int nT,
                    const float* __restrict__ win_s, float* __restrict__ frames, float* __restrict__ out_windowed, int advance) {
  __shared__ float2 fft[kRtMaxN];
  const int t = blockIdx.x, F = N / 2 + 1;
  for (int k = threadIdx.x; k < F; k += blockDim.x) {
    float2 a = Y[((int64_t)0 * F + k) * nT + t], b = Y[((int64_t)1 * F + k) * nT + t];
    if (k == 0 || k == N / 2) { a.y = 0.f; b.y = 0.f; }          // numpy.fft.irfft ignores them
    fft[bitrev(k, log2n)] = float2{a.x - b.y, a.y + b.x};
    if (k != 0 && k != N / 2) fft[bitrev(N - k, log2n)] = float2{a.x + b.y, b.x - a.y};
  }
  __syncthreads();
  fft_inplace<float2, float>(fft, tw, N, log2n, true);
  const float inv_n = 1.0f / (float)N;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float2 z = fft[i];
    const float w = win_s[i];
    const float l = z.x * inv_n * w, r = z.y * inv_n * w;        // irfft(...) * synthesisWindowFunction (:231)
    frames[((int64_t)0 * nT + t) * N + i] = l;
    frames[((int64_t)1 * nT + t) * N + i] = r;
    if (out_windowed) {
      out_windowed[((int64_t)0 * N + i) * nT + t] = l;
      out_windowed[((int64_t)1 * N + i) * nT + t] = r;
    }
  }
  if (advance && t == 0 && threadIdx.x == 0) atomicAdd(&dev->block_counter, 1);     // the ring kernel that follows uses counter (= this call's number)
}

// ---------------------------------------------------------------------------------------------- localisation (one CTA, D <= 128 threads used)
__device__ void rt_localize(RtDev* dev, const float* __restrict__ gccphat, int D, int nT, double* __restrict__ hist, int hist_len) {
  __shared__ double mean_s[kRtMaxD];
  const int d = threadIdx.x;
  int idx = dev->hist_index;
  // gccPHATHistory.set(nanmean(realGCC, axis=0).T)   (:214 -> utils.py:45-59)
  if (d < D)
    for (int t = 0; t < nT; ++t) hist[(int64_t)d * hist_len + (idx + t) % hist_len] = (double)gccphat[(int64_t)d * nT + t];
  idx = (idx + nT) % hist_len;
  __syncthreads();
  if (d < D) {
    // nanmean over the last loc_window columns of the unravelled history (:221-222)
    const int w = min(max(dev->loc_window, 1), hist_len);
    double s = 0.0;
    int n = 0;
    for (int j = 0; j < w; ++j) {
      const double v = hist[(int64_t)d * hist_len + (idx - 1 - j + 2 * hist_len) % hist_len];
      if (v == v) { s += v; ++n; }
    }
    mean_s[d] = n > 0 ? s / (double)n : __longlong_as_double(0x7ff8000000000000LL);
  }
  __syncthreads();
  if (d == 0) {
    if (dev->localization) {
      double bv = mean_s[0];
      int bi = 0;
      for (int j = 1; j < D; ++j)
        if (rt_better64(mean_s[j], j, bv, bi)) { bv = mean_s[j]; bi = j; }
      dev->target = (float)bi;                       // targetTDOAIndex.set_value(tdoaIndex)
    }
    dev->hist_index = idx;
  }
}

__global__ void __launch_bounds__(128)
rt_localize_kernel(RtDev* dev, const float* __restrict__ gccphat, int D, int nT, double* __restrict__ hist, int hist_len) {
  rt_localize(dev, gccphat, D, nT, hist, hist_len);
}

// ---------------------------------------------------------------------------------------------- E: overlap-add ring, block emit, input push
// CTAs [0, gridDim.x - 1): one thread per logical ring position of the range that changes or is emitted; last CTA: localisation.
__global__ void __launch_bounds__(128)
rt_ola_emit_kernel(RtDev* dev, const float* __restrict__ frames, int N, int hop, int nT, int B, int L, int p_first, float* __restrict__ out_ring,
                   float* __restrict__ out_block, const float* __restrict__ in_block, float* __restrict__ in_ring, const float* __restrict__ gccphat,
                   int D, double* __restrict__ hist, int hist_len) {
  if (blockIdx.x == gridDim.x - 1) {
    rt_localize(dev, gccphat, D, nT, hist, hist_len);
    return;
  }
  const int call = dev->block_counter;               // already advanced by the synthesis kernel
  const int p = p_first + blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= L) return;
  const int q = rt_ring_index(p, call, B, L);
  const int w_first = L - N - (nT - 1) * hop;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    float acc = p >= L - B ? 0.f : out_ring[c * L + q];          // outputBuffer[:, -B:] = 0 (utils.py:105)
    if (p >= w_first) {
      for (int i = 0; i < nT; ++i) {                              // frame order, float32 buffer += float64 frame (utils.py:113-114)
        const int r = p - (w_first + i * hop);
        if (r >= 0 && r < N) acc = (float)((double)acc + (double)frames[((int64_t)c * nT + i) * N + r]);
      }
    }
    out_ring[c * L + q] = acc;
    if (p >= L - 3 * B && p < L - 2 * B) out_block[c * B + (p - (L - 3 * B))] = acc;   // outputFrames = outputBuffer[:, -3B:-2B] (:115)
    if (p >= L - B) in_ring[c * L + q] = in_block[c * B + (p - (L - B))];              // inputBuffer[:, -B:] = inputFrames (:102)
  }
}

int ilog2_of(int n) {
  int l = 0;
  while ((1 << l) < n) ++l;
  return l;
}

#define RT_CARVE_OR_FAIL(l)                                                                                                        \
  if (int st__ = rt_check(h, cfg)) return st__;                                                                                    \
  RtLayout l = rt_carve(*cfg, state, state_bytes);                                                                                 \
  if (!l.ok) return gccnmf_fail(h, GCCNMF_ERR_WORKSPACE, "rt state buffer missing or too small: need %zu bytes", l.bytes)

// Kernels A .. D (+ inference) of one block; `windowed` != NULL: frames given by the caller (processFrames), else cut from the ring.
int rt_enqueue_core(gccnmf_handle* h, const gccnmf_rt_config* cfg, const RtLayout& l, const float* windowed, const float* in_block, float* out_windowed,
                    const double* forced_mask, int advance, void* stream) {
  const int N = cfg->window_size, nT = cfg->windows_per_block, K = cfg->num_atoms, D = cfg->num_tdoas, F = l.F, log2n = ilog2_of(N);
  const int inf = cfg->inference_iterations;
  GCCNMF_LAUNCH(h, rt_analysis_kernel, nT, kFftThreads, 0, stream, l.dev, windowed, in_block, l.in_ring, cfg->block_size, l.L, cfg->hop_size, nT, l.win_a,
                reinterpret_cast<const double2*>(l.tw64), N, log2n, l.ET, D, l.Fp, l.X, l.G, l.gccphat, l.Vabs, l.H0, l.H, K, inf > 0 ? 1 : 0);
  for (int it = 0; it < inf; ++it) {
#define RT_INF_CASE(J)                                                                                                             \
    case J:                                                                                                                        \
      GCCNMF_LAUNCH(h, rt_inf_ratio_kernel<2 * J>, (F + 7) / 8, 256, 0, stream, l.W, l.H, l.Vabs, F, K, l.R);                        \
      GCCNMF_LAUNCH(h, rt_inf_update_kernel<2 * J>, (K + 7) / 8, 256, 0, stream, l.WT, l.Fp, l.R, F, K, l.colsumW, cfg->sparsity_alpha, cfg->epsilon, l.H); \
      break;
    switch (nT) {
      RT_INF_CASE(1) RT_INF_CASE(2) RT_INF_CASE(3) RT_INF_CASE(4) RT_INF_CASE(5) RT_INF_CASE(6) RT_INF_CASE(7) RT_INF_CASE(8)
    }
#undef RT_INF_CASE
  }
  const dim3 grid_b((K + kRtAtomsPerCta - 1) / kRtAtomsPerCta, nT);
  if (forced_mask) {     // teacher forcing / externally decided masks: the filter uses the caller's (K, nT) float64 atom mask
    GCCNMF_CHECK_CUDA(h, cudaMemcpyAsync(l.hmask, forced_mask, (size_t)K * nT * sizeof(double), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  } else if (D <= 32) GCCNMF_LAUNCH(h, rt_atoms_kernel<2>, grid_b, 256, 0, stream, l.dev, l.G, l.W, F, l.Fp, K, D, nT, l.argmax, l.hmask);
  else if (D <= 64) GCCNMF_LAUNCH(h, rt_atoms_kernel<4>, grid_b, 256, 0, stream, l.dev, l.G, l.W, F, l.Fp, K, D, nT, l.argmax, l.hmask);
  else GCCNMF_LAUNCH(h, rt_atoms_kernel<8>, grid_b, 256, 0, stream, l.dev, l.G, l.W, F, l.Fp, K, D, nT, l.argmax, l.hmask);
  GCCNMF_LAUNCH(h, rt_filter_kernel, (F + 7) / 8, 256, 0, stream, l.dev, l.W, l.hmask, l.recV, l.H, inf > 0 ? 1 : 0, l.X, F, K, nT, l.Y);
  GCCNMF_LAUNCH(h, rt_synthesis_kernel, nT, kFftThreads, 0, stream, l.dev, l.Y, reinterpret_cast<const float2*>(l.tw32), N, log2n, nT, l.win_s, l.frames,
                out_windowed, advance);
  return 0;
}

}  // namespace

extern "C" {

size_t gccnmf_rt_state_bytes(const gccnmf_rt_config* cfg) {
  if (!cfg || cfg->window_size < 2 || cfg->block_size < 1 || cfg->windows_per_block < 1 || cfg->num_atoms < 1 || cfg->num_tdoas < 1 ||
      cfg->history_length < 1)
    return 0;
  return rt_carve(*cfg, nullptr, 0).bytes;
}

// W (F, K) f32, E (F, D) complex64 (expJOmegaTau, gccNMFProcessor.py:248), windows (N) f32, H0 (K, 2) f32 or NULL (all device pointers).
int gccnmf_rt_init(gccnmf_handle* h, const gccnmf_rt_config* cfg, const float* W, const float* E, const float* analysis_window,
                   const float* synthesis_window, const float* H0, void* state, size_t state_bytes, void* stream) {
  GCCNMF_ENTER(h);
  RT_CARVE_OR_FAIL(l);
  GCCNMF_REQUIRE(h, W && E && analysis_window && synthesis_window, "rt_init: NULL pointer");
  GCCNMF_REQUIRE(h, cfg->inference_iterations == 0 || H0 != nullptr, "rt_init: coefficient inference needs the initial H0 (K, 2)");
  cudaStream_t s = (cudaStream_t)stream;
  const int N = cfg->window_size, K = cfg->num_atoms, D = cfg->num_tdoas;
  GCCNMF_CHECK_CUDA(h, cudaMemsetAsync(state, 0, l.bytes, s));                 // rings, history (initValue = 0), counters
  const double* tw64 = nullptr;
  const float* tw32 = nullptr;
  if (int st = gccnmf_get_twiddles(h, N, &tw64, &tw32)) return st;
  GCCNMF_CHECK_CUDA(h, cudaMemcpyAsync(l.tw64, tw64, (size_t)N * sizeof(double), cudaMemcpyDeviceToDevice, s));
  GCCNMF_CHECK_CUDA(h, cudaMemcpyAsync(l.tw32, tw32, (size_t)N * sizeof(float), cudaMemcpyDeviceToDevice, s));
  GCCNMF_CHECK_CUDA(h, cudaMemcpyAsync(l.win_a, analysis_window, (size_t)N * sizeof(float), cudaMemcpyDeviceToDevice, s));
  GCCNMF_CHECK_CUDA(h, cudaMemcpyAsync(l.win_s, synthesis_window, (size_t)N * sizeof(float), cudaMemcpyDeviceToDevice, s));
  GCCNMF_CHECK_CUDA(h, cudaMemcpyAsync(l.W, W, (size_t)l.F * K * sizeof(float), cudaMemcpyDeviceToDevice, s));
  if (H0) GCCNMF_CHECK_CUDA(h, cudaMemcpyAsync(l.H0, H0, (size_t)K * 2 * sizeof(float), cudaMemcpyDeviceToDevice, s));
  const int n = K > l.F ? K : l.F;
  GCCNMF_LAUNCH(h, rt_init_dictionary_kernel, (n + 127) / 128, 128, 0, stream, l.W, l.F, l.Fp, K, l.WT, l.recV, l.colsumW);
  GCCNMF_LAUNCH(h, rt_init_steering_kernel, (D * l.Fp + 255) / 256, 256, 0, stream, reinterpret_cast<const float2*>(E), l.F, l.Fp, D, l.ET);
  // defaults of gccNMFProcessor.py:190-199
  GCCNMF_LAUNCH(h, rt_set_params_kernel, 1, 1, 0, stream, l.dev, 10.0f, 2.0f, 1.0f, 0.0f, 1, 1, 0, 6, 1);
  return GCCNMF_OK;
}

// setTargetTDOARange (:272-276) + the settable attributes (:136-151).  set_target = 0 leaves the target TDOA index alone (it is
// loop-carried device state when localisation is on).
int gccnmf_rt_set_params(gccnmf_handle* h, const gccnmf_rt_config* cfg, void* state, size_t state_bytes, float target_index, int set_target,
                         float epsilon, float beta, float noise_floor, int mode, int separation_enabled, int localization_enabled,
                         int localization_window, void* stream) {
  GCCNMF_ENTER(h);
  RT_CARVE_OR_FAIL(l);
  GCCNMF_REQUIRE(h, mode == 0 || mode == 1, "rt_set_params: mode must be 0 (boxcar) or 1 (window)");
  GCCNMF_LAUNCH(h, rt_set_params_kernel, 1, 1, 0, stream, l.dev, target_index, epsilon, beta, noise_floor, mode, separation_enabled ? 1 : 0,
                localization_enabled ? 1 : 0, localization_window, set_target ? 1 : 0);
  return GCCNMF_OK;
}

// GCCNMFProcessor.processFrames (:201-231): windowed (2, N, nT) f32 -> out (2, N, nT) f32, both on the device.
int gccnmf_rt_process_frames(gccnmf_handle* h, const gccnmf_rt_config* cfg, void* state, size_t state_bytes, const float* windowed, float* out,
                             const double* forced_atom_mask, void* stream) {
  GCCNMF_ENTER(h);
  RT_CARVE_OR_FAIL(l);
  GCCNMF_REQUIRE(h, windowed && out, "rt_process_frames: NULL pointer");
  if (int st = rt_enqueue_core(h, cfg, l, windowed, nullptr, out, forced_atom_mask, 0, stream)) return st;
  GCCNMF_LAUNCH(h, rt_localize_kernel, 1, 128, 0, stream, l.dev, l.gccphat, cfg->num_tdoas, cfg->windows_per_block, l.hist, cfg->history_length);
  return GCCNMF_OK;
}

// OverlapAddProcessor.processFrames(GCCNMFProcessor.processFrames) (utils.py:99-116 around gccNMFProcessor.py:201-231):
// in_block (2, B) f32 -> out_block (2, B) f32 (the block emitted is the one pushed in two calls earlier).
int gccnmf_rt_process_block(gccnmf_handle* h, const gccnmf_rt_config* cfg, void* state, size_t state_bytes, const float* in_block, float* out_block,
                            const double* forced_atom_mask, void* stream) {
  GCCNMF_ENTER(h);
  RT_CARVE_OR_FAIL(l);
  GCCNMF_REQUIRE(h, in_block && out_block, "rt_process_block: NULL pointer");
  if (int st = rt_enqueue_core(h, cfg, l, nullptr, in_block, nullptr, forced_atom_mask, 1, stream)) return st;
  const int N = cfg->window_size, nT = cfg->windows_per_block, B = cfg->block_size;
  const int w_first = l.L - N - (nT - 1) * cfg->hop_size;
  const int p_first = w_first < l.L - 3 * B ? w_first : l.L - 3 * B;
  const int ctas = (l.L - p_first + 127) / 128;
  GCCNMF_LAUNCH(h, rt_ola_emit_kernel, ctas + 1, 128, 0, stream, l.dev, l.frames, N, cfg->hop_size, nT, B, l.L, p_first, l.out_ring, out_block, in_block,
                l.in_ring, l.gccphat, cfg->num_tdoas, l.hist, cfg->history_length);
  return GCCNMF_OK;
}

// One block as a CUDA graph: [H2D of in_host ->] the kernels of gccnmf_rt_process_block [-> D2H to out_host].  in_block / out_block
// are device staging buffers (2, B); in_host / out_host pinned host buffers or NULL.  *graph_exec is a cudaGraphExec_t.
int gccnmf_rt_graph_create(gccnmf_handle* h, const gccnmf_rt_config* cfg, void* state, size_t state_bytes, float* in_block, float* out_block,
                           const float* in_host, float* out_host, void** graph_exec, void* stream) {
  GCCNMF_ENTER(h);
  GCCNMF_REQUIRE(h, graph_exec != nullptr && stream != nullptr, "rt_graph_create: needs a non-default stream and an output slot");
  *graph_exec = nullptr;
  if (int st = rt_check(h, cfg)) return st;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t block_bytes = (size_t)2 * cfg->block_size * sizeof(float);
  GCCNMF_CHECK_CUDA(h, cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
  int st = GCCNMF_OK;
  if (in_host && cudaMemcpyAsync(in_block, in_host, block_bytes, cudaMemcpyHostToDevice, s) != cudaSuccess) st = GCCNMF_ERR_CUDA;
  if (st == GCCNMF_OK) st = gccnmf_rt_process_block(h, cfg, state, state_bytes, in_block, out_block, nullptr, stream);
  if (st == GCCNMF_OK && out_host && cudaMemcpyAsync(out_host, out_block, block_bytes, cudaMemcpyDeviceToHost, s) != cudaSuccess) st = GCCNMF_ERR_CUDA;
  cudaGraph_t graph = nullptr;
  const cudaError_t end = cudaStreamEndCapture(s, &graph);
  if (st != GCCNMF_OK || end != cudaSuccess) {
    if (graph) cudaGraphDestroy(graph);
    if (st == GCCNMF_OK) return gccnmf_fail(h, GCCNMF_ERR_CUDA, "rt_graph_create: stream capture failed: %s", cudaGetErrorString(end));
    return st;
  }
  cudaGraphExec_t exec = nullptr;
  const cudaError_t inst = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (inst != cudaSuccess) return gccnmf_fail(h, GCCNMF_ERR_CUDA, "rt_graph_create: cudaGraphInstantiate failed: %s", cudaGetErrorString(inst));
  *graph_exec = exec;
  return GCCNMF_OK;
}

int gccnmf_rt_graph_launch(gccnmf_handle* h, void* graph_exec, void* stream) {
  GCCNMF_ENTER(h);
  GCCNMF_REQUIRE(h, graph_exec != nullptr, "rt_graph_launch: NULL graph");
  GCCNMF_CHECK_CUDA(h, cudaGraphLaunch((cudaGraphExec_t)graph_exec, (cudaStream_t)stream));
  h->launches++;
  return GCCNMF_OK;
}

int gccnmf_rt_graph_destroy(gccnmf_handle* h, void* graph_exec) {
  GCCNMF_ENTER(h);
  if (graph_exec) GCCNMF_CHECK_CUDA(h, cudaGraphExecDestroy((cudaGraphExec_t)graph_exec));
  return GCCNMF_OK;
}

// Copies one piece of the block state to dst (device or pinned host memory, cudaMemcpyDefault), stream-ordered:
//   0 gccPHAT (D, nT) f32   1 target TDOA index (1) f32   2 atom mask (K, nT) f64   3 input spectrogram X (2, F, nT) c64
//   4 output spectrogram (2, F, nT) c64   5 TDOA argmax per atom (K, nT) i32   6 inferred coefficients H (K, 2 nT) f32
//   7 GCC-PHAT history ring (D, history_length) f64 followed by nothing (its write index is item 8)   8 history write index (1) i32
int gccnmf_rt_export(gccnmf_handle* h, const gccnmf_rt_config* cfg, void* state, size_t state_bytes, int what, void* dst, void* stream) {
  GCCNMF_ENTER(h);
  RT_CARVE_OR_FAIL(l);
  GCCNMF_REQUIRE(h, dst != nullptr, "rt_export: NULL destination");
  const size_t nT = cfg->windows_per_block, K = cfg->num_atoms, D = cfg->num_tdoas, F = l.F;
  const void* src = nullptr;
  size_t bytes = 0;
  switch (what) {
    case 0: src = l.gccphat; bytes = D * nT * sizeof(float); break;
    case 1: src = &l.dev->target; bytes = sizeof(float); break;
    case 2: src = l.hmask; bytes = K * nT * sizeof(double); break;
    case 3: src = l.X; bytes = 2 * F * nT * sizeof(float2); break;
    case 4: src = l.Y; bytes = 2 * F * nT * sizeof(float2); break;
    case 5: src = l.argmax; bytes = K * nT * sizeof(int32_t); break;
    case 6: src = l.H; bytes = K * 2 * nT * sizeof(float); break;
    case 7: src = l.hist; bytes = D * (size_t)cfg->history_length * sizeof(double); break;
    case 8: src = &l.dev->hist_index; bytes = sizeof(int); break;
    default: return gccnmf_fail(h, GCCNMF_ERR_INVALID_ARGUMENT, "rt_export: unknown item %d", what);
  }
  GCCNMF_CHECK_CUDA(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
  return GCCNMF_OK;
}

}  // extern "C"
