// GCC-PHAT localisation and GCC-NMF masking kernels.
//   phat_angspec       runGCCNMF.py:44 + gccNMFFunctions.py:85-92      (a3, a4)
//   tdoa_gccnmf        gccNMFFunctions.py:118-135; offlineSpeechEnhancement.ipynb:444-450; online :422  (a6, a10, a11)
//   coeff_mask         gccNMFFunctions.py:137-143                        (a7)
//   argmax_mask        offlineSpeechEnhancement.ipynb:466-472            (a10 mask)
//   masked_recon_phase gccNMFFunctions.py:145-151                        (a8)
// Contractions that the reference evaluates in complex128 / float64 are accumulated in float64 here
// so that the integer decisions taken on them (peak picking, argmax over TDOA) are the reference's.
#include "common.cuh"
#include "gemm_simt.cuh"

namespace {

// ------------------------------------------------------------------ a3 + a4: coherence + angular spectrogram
constexpr int kAngT = 16;     // frames per CTA
constexpr int kAngWarps = 4;  // warps per CTA: each takes a quarter of the staged bins (partial sums added in warp order at the end)
constexpr int kAngMaxD = 128;

// numpy's complex64 arithmetic for  X0 * conj(X1) / |X0| / |X1|  (runGCCNMF.py:44): float32 products,
// magnitude correctly rounded (hypotf), division by a real done as multiplication by the float32
// reciprocal (numpy's complex division with a zero imaginary divisor).
__device__ __forceinline__ float2 phat_coherence(float2 a, float2 b) {
  float re = __fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));
  float im = __fsub_rn(__fmul_rn(a.y, b.x), __fmul_rn(a.x, b.y));
  const float ma = (float)sqrt((double)a.x * a.x + (double)a.y * a.y);
  const float mb = (float)sqrt((double)b.x * b.x + (double)b.y * b.y);
  const float ia = 1.0f / ma, ib = 1.0f / mb;
  re = __fmul_rn(re, ia); im = __fmul_rn(im, ia);
  re = __fmul_rn(re, ib); im = __fmul_rn(im, ib);
  return float2{re, im};
}

// CTA = 16 frames x all TDOAs, 4 warps.  A lane owns the TDOAs lane, lane + 32, ... (DPL of them) for ALL 16 frames of the tile:
// per staged bin it reads its DPL steering values once and then one broadcast coherence value per frame, i.e. 2 DPL float64 FMAs
// per shared-memory wavefront -- the previous layout (lane = frame, one broadcast E value per FMA pair) spent 6 wavefronts per 4 FMA
// instructions.  The four warps split the bins of every staged chunk; their partial sums are added in warp order at the end
// (deterministic).  The global loads of chunk i + 1 (spectrogram pair, steering rows) are issued into registers before chunk i is
// consumed: with one CTA of 4 warps per SM the un-overlapped load latency of 33-65 chunks was most of the kernel's time.
// Every CTA also writes the coherence of its frames.
template <int DPL, int BF>
__global__ void __launch_bounds__(kAngWarps * 32)
phat_angspec_kernel(const float2* __restrict__ X, int F, int T, int x_is_coherence, const double2* __restrict__ E, int D,
                    float2* __restrict__ coherence, double* __restrict__ angular, double* __restrict__ tile_sums) {
  constexpr int kThreads = kAngWarps * 32;
  constexpr int kCPerThread = BF * kAngT / kThreads;           // coherence values a thread stages per chunk
  constexpr int kEPerThread = BF * 32 * DPL / kThreads;         // steering values a thread stages per chunk (D <= 32 DPL)
  static_assert(BF * kAngT % kThreads == 0 && BF % kAngWarps == 0, "chunk shape");
  __shared__ double2 Cs[BF][kAngT];
  __shared__ double2 Es[(BF > kAngT / 2 ? BF : kAngT / 2)][32 * DPL];   // reused as the reduction buffer red[kAngT][32 DPL] (doubles) at the end
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int t0 = blockIdx.x * kAngT;
  const bool accumulate = angular != nullptr || tile_sums != nullptr;
  double acc[kAngT][DPL];
#pragma unroll
  for (int tt = 0; tt < kAngT; ++tt)
#pragma unroll
    for (int j = 0; j < DPL; ++j) acc[tt][j] = 0.0;

  float2 xa[kCPerThread], xb[kCPerThread];
  double2 er[kEPerThread];
  auto fetch = [&](int f0) {
#pragma unroll
    for (int q = 0; q < kCPerThread; ++q) {
      const int e = threadIdx.x + q * kThreads, f = f0 + e / kAngT, t = t0 + e % kAngT;
      xa[q] = xb[q] = float2{0.f, 0.f};
      if (f < F && t < T) {
        xa[q] = X[(int64_t)f * T + t];
        if (!x_is_coherence) xb[q] = X[((int64_t)F + f) * T + t];
      }
    }
    if (accumulate) {
#pragma unroll
      for (int q = 0; q < kEPerThread; ++q) {
        const int e = threadIdx.x + q * kThreads, ff = e / (32 * DPL), d = e % (32 * DPL);
        er[q] = (f0 + ff < F && d < D) ? E[(int64_t)(f0 + ff) * D + d] : double2{0.0, 0.0};
      }
    }
  };
  auto stash = [&](int f0) {
#pragma unroll
    for (int q = 0; q < kCPerThread; ++q) {
      const int e = threadIdx.x + q * kThreads, ff = e / kAngT, tt = e % kAngT;
      const int f = f0 + ff, t = t0 + tt;
      double2 c = double2{0.0, 0.0};
      if (f < F && t < T) {
        const float2 coh = x_is_coherence ? xa[q] : phat_coherence(xa[q], xb[q]);
        if (coherence) coherence[(int64_t)f * T + t] = coh;
        c = double2{(double)coh.x, (double)coh.y};
      }
      Cs[ff][tt] = c;
    }
    if (accumulate) {
#pragma unroll
      for (int q = 0; q < kEPerThread; ++q) {
        const int e = threadIdx.x + q * kThreads;
        Es[e / (32 * DPL)][e % (32 * DPL)] = er[q];
      }
    }
  };

  fetch(0);
  for (int f0 = 0; f0 < F; f0 += BF) {
    stash(f0);
    __syncthreads();
    if (f0 + BF < F) fetch(f0 + BF);
    if (accumulate) {
#pragma unroll
      for (int i = 0; i < BF / kAngWarps; ++i) {
        const int ff = w + kAngWarps * i;
        double2 e[DPL];
#pragma unroll
        for (int j = 0; j < DPL; ++j) e[j] = Es[ff][lane + 32 * j];
#pragma unroll
        for (int tt = 0; tt < kAngT; ++tt) {
          const double2 c = Cs[ff][tt];
#pragma unroll
          for (int j = 0; j < DPL; ++j) acc[tt][j] += c.x * e[j].x - c.y * e[j].y;   // Re(C * E)
        }
      }
    }
    __syncthreads();
  }
  if (!accumulate) return;
  double (*red)[32 * DPL] = reinterpret_cast<double (*)[32 * DPL]>(&Es[0][0]);
  for (int r = 0; r < kAngWarps; ++r) {
    if (w == r) {
#pragma unroll
      for (int tt = 0; tt < kAngT; ++tt)
#pragma unroll
        for (int j = 0; j < DPL; ++j) {
          const int d = lane + 32 * j;
          red[tt][d] = r == 0 ? acc[tt][j] : red[tt][d] + acc[tt][j];
        }
    }
    __syncthreads();
  }
  const int t_valid = min(kAngT, T - t0);
  if (angular)
    for (int e = threadIdx.x; e < kAngT * D; e += blockDim.x) {
      const int d = e / kAngT, tt = e % kAngT;
      if (tt < t_valid) angular[(int64_t)d * T + t0 + tt] = red[tt][d];
    }
  if (tile_sums)
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
      double sum = 0.0;
      for (int tt = 0; tt < t_valid; ++tt) sum += red[tt][d];
      tile_sums[(int64_t)blockIdx.x * D + d] = sum;
    }
}

__global__ void mean_tiles_kernel(const double* __restrict__ tile_sums, int tiles, int D, int T, double* __restrict__ mean) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  double s = 0.0;
  for (int i = 0; i < tiles; ++i) s += tile_sums[(int64_t)i * D + d];
  mean[d] = s / (double)T;
}

// ------------------------------------------------------------------ a6 / a10 / a11: GCC-NMF per TDOA (float64 GEMM)
constexpr int GM = 128, GN = 128, GK = 8, GTM = 8, GTN = 8;
constexpr int kGccThreads = (GM / GTM) * (GN / GTN);

struct LoadWAtoms {  // A(m = atom, k = f) = W[f][atom]
  static constexpr bool kContigK = false;
  const float* W; int K, F;
  __device__ double operator()(int m, int f) const { return (m < K && f < F) ? (double)__ldg(W + (int64_t)f * K + m) : 0.0; }
};
struct LoadRealGCC {  // B(n = t * D + d, k = f) = Re(coherence[f][t] * E[f][d])
  static constexpr bool kContigK = false;
  const float2* coh; const double2* E; int F, T, D, N;
  __device__ double operator()(int n, int f) const {
    if (n >= N || f >= F) return 0.0;
    const int t = n / D, d = n - t * D;
    const float2 c = __ldg(coh + (int64_t)f * T + t);
    const double2 e = __ldg(E + (int64_t)f * D + d);
    return (double)c.x * e.x - (double)c.y * e.y;
  }
};

// numpy.argmax ordering: NaN is a maximum, first occurrence wins.
__device__ __forceinline__ bool argmax_better(double v, int i, double bv, int bi) {
  const bool vn = isnan(v), bn = isnan(bv);
  if (vn || bn) return vn && (!bn || i < bi);
  return v > bv || (v == bv && i < bi);
}

template <bool ARGMAX>
__global__ void __launch_bounds__(kGccThreads)
tdoa_gccnmf_kernel(int K, int N, int F, LoadWAtoms aload, LoadRealGCC bload, int T, int D, float* __restrict__ values,
                   int32_t* __restrict__ argmax) {
  double acc[GTM][GTN];
  const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
  gemm_simt_mainloop<double, GM, GN, GK, GTM, GTN>(acc, m0, n0, F, aload, bload);
  constexpr int TX = GN / GTN;
  if (values) {
#pragma unroll
    for (int i = 0; i < GTM; ++i) {
      const int m = gemm_row<GM, GTM, TX>(m0, i);
      if (m >= K) continue;
#pragma unroll
      for (int j = 0; j < GTN; ++j) {
        const int n = gemm_col<GN, GTN, TX>(n0, j);
        if (n >= N) continue;
        const int t = n / D, d = n - t * D;
        values[((int64_t)d * K + m) * T + t] = (float)acc[i][j];
      }
    }
  }
  if (ARGMAX) {
    // D is a power of two in [4, 128] and divides GN, so n0 % D == 0 and every 4-column half-row of the
    // register tile lies inside one frame.  D <= 64: a frame spans D/4 consecutive lanes of one half;
    // D == 128: a frame is the whole tile row (both halves of all 16 lanes).
    const int tx = threadIdx.x % TX;
    const int lanes = min(D / 4, TX);
#pragma unroll
    for (int i = 0; i < GTM; ++i) {
      const int m = gemm_row<GM, GTM, TX>(m0, i);
      double bv[2];
      int bi[2];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int d0 = (half * (GN / 2) + tx * (GTN / 2)) % D;  // TDOA index of this half-row's first column
        bv[half] = acc[i][half * (GTN / 2)];
        bi[half] = d0;
#pragma unroll
        for (int j = 1; j < GTN / 2; ++j) {
          const double v = acc[i][half * (GTN / 2) + j];
          if (argmax_better(v, d0 + j, bv[half], bi[half])) { bv[half] = v; bi[half] = d0 + j; }
        }
      }
      if (D == GN) {  // one frame per tile row: fold half 1 into half 0 before the lane reduction
        if (argmax_better(bv[1], bi[1], bv[0], bi[0])) { bv[0] = bv[1]; bi[0] = bi[1]; }
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        if (D == GN && half == 1) break;
        for (int o = 1; o < lanes; o <<= 1) {
          const double ov = __shfl_xor_sync(0xffffffffu, bv[half], o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi[half], o);
          if (argmax_better(ov, oi, bv[half], bi[half])) { bv[half] = ov; bi[half] = oi; }
        }
        const int n_first = n0 + half * (GN / 2) + tx * (GTN / 2);
        if ((tx % lanes) == 0 && m < K && n_first < N) argmax[(int64_t)m * T + n_first / D] = bi[half];
      }
    }
  }
}

// ------------------------------------------------------------------ a7: masks
__global__ void coeff_mask_kernel(const float* __restrict__ G, int S, int64_t KT, float* __restrict__ masks, int32_t* all_nan_flag) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= KT) return;
  int best = -1;
  float bv = 0.f;
  for (int s = 0; s < S; ++s) {  // numpy.nanargmax: ignore NaN, first maximum wins
    const float v = G[(int64_t)s * KT + i];
    if (isnan(v)) continue;
    if (best < 0 || v > bv) { best = s; bv = v; }
  }
  if (best < 0 && all_nan_flag) *all_nan_flag = 1;
  for (int s = 0; s < S; ++s) masks[(int64_t)s * KT + i] = (s == best) ? 1.f : 0.f;
}

__global__ void argmax_mask_kernel(const int32_t* __restrict__ argmax, int64_t KT, const uint8_t* __restrict__ lut, int D, float* __restrict__ mask) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= KT) return;
  const int a = argmax[i];
  mask[i] = (a >= 0 && a < D && lut[a]) ? 1.f : 0.f;
}

// ------------------------------------------------------------------ a8: masked reconstruction with mixture phase
constexpr int RM = 128, RN = 128, RK = 16, RTM = 8, RTN = 8;
constexpr int kReconThreads = (RM / RTM) * (RN / RTN);

struct LoadWRows {  // A(m = f, k = atom) = W[f][atom]
  static constexpr bool kContigK = true;
  const float* W; int F, K;
  __device__ float operator()(int m, int k) const { return (m < F && k < K) ? __ldg(W + (int64_t)m * K + k) : 0.f; }
};
struct LoadMaskedH {  // B(n = t, k = atom) = H[atom][c*T + t] * mask[atom][t]   (gccNMFFunctions.py:150)
  static constexpr bool kContigK = false;
  const float* H; const float* mask; int K, T; int64_t ldh;
  __device__ float operator()(int n, int k) const {
    return (n < T && k < K) ? __ldg(H + (int64_t)k * ldh + n) * __ldg(mask + (int64_t)k * T + n) : 0.f;
  }
};

__global__ void __launch_bounds__(kReconThreads, 2)
masked_recon_kernel(const float* __restrict__ masks, const float2* __restrict__ X, const float* __restrict__ W,
                    const float* __restrict__ H, int F, int T, int K, float2* __restrict__ out) {
  const int s = blockIdx.z / 2, c = blockIdx.z % 2;
  float acc[RTM][RTN];
  const int m0 = blockIdx.y * RM, n0 = blockIdx.x * RN;
  LoadWRows a{W, F, K};
  LoadMaskedH b{H + (int64_t)c * T, masks + (int64_t)s * K * T, K, T, (int64_t)2 * T};
  gemm_simt_mainloop<float, RM, RN, RK, RTM, RTN>(acc, m0, n0, K, a, b);
  const float2* Xc = X + (int64_t)c * F * T;
  float2* o = out + ((int64_t)s * 2 + c) * F * T;
#pragma unroll
  for (int i = 0; i < RTM; ++i) {
    const int m = gemm_row<RM, RTM, RN / RTN>(m0, i);
    if (m >= F) continue;
#pragma unroll
    for (int j = 0; j < RTN; ++j) {
      const int n = gemm_col<RN, RTN, RN / RTN>(n0, j);
      if (n >= T) continue;
      // exp(1j * angle(X)) (gccNMFFunctions.py:151): unit phasor of the mixture bin; angle(0) = 0.
      const float2 x = Xc[(int64_t)m * T + n];
      const double mag = sqrt((double)x.x * x.x + (double)x.y * x.y);
      float pr = 1.f, pi = 0.f;
      if (mag > 0.0) { pr = (float)((double)x.x / mag); pi = (float)((double)x.y / mag); }
      else if (mag != mag) { pr = pi = __int_as_float(0x7fc00000); }
      o[(int64_t)m * T + n] = float2{acc[i][j] * pr, acc[i][j] * pi};
    }
  }
}

// ------------------------------------------------------------------ a11 / a13: online localisation, atom masks, Wiener-like filter
// acc[tau] = max over frames <= t of A[tau, t'] (onlineSpeechEnhancement.ipynb:416); one thread per TDOA scans time.
__global__ void cummax_time_kernel(const double* __restrict__ A, int D, int T, double* __restrict__ acc) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  double m = -INFINITY;
  for (int t = 0; t < T; ++t) {
    const double v = A[(int64_t)d * T + t];
    if (v > m || v != v) m = v;          // numpy.max: NaN propagates (and then sticks: comparisons with NaN are false)
    acc[(int64_t)d * T + t] = m;
  }
}
// target[t] = argmax over TDOA of acc[:, t] (:417), numpy.argmax semantics.
__global__ void argmax_tdoa_kernel(const double* __restrict__ acc, int D, int T, int32_t* __restrict__ target) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  double bv = acc[t];
  int bi = 0;
  for (int d = 1; d < D; ++d) {
    const double v = acc[(int64_t)d * T + t];
    if (argmax_better(v, d, bv, bi)) { bv = v; bi = d; }
  }
  target[t] = bi;
}

// mode 0 (boxcar): |argmax - target| < eps -> 1 else 0      (onlineSpeechEnhancement.ipynb:423-425; gccNMFProcessor.py:263)
// mode 1 (window): exp(-(|argmax - target| / eps)^beta) / (1 + floor) + floor     (gccNMFProcessor.py:265)
__global__ void atom_mask_kernel(const int32_t* __restrict__ argmax, int K, int T, const int32_t* __restrict__ target, int target_stride,
                                 float target_scalar, float eps, int mode, float beta, float noise_floor, float* __restrict__ mask) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)K * T) return;
  const int t = (int)(i % T);
  const float mu = target ? (float)target[(int64_t)t * target_stride] : target_scalar;
  const float dist = fabsf((float)argmax[i] - mu);
  float m;
  if (mode == 0) m = dist < eps ? 1.f : 0.f;
  else m = expf(-powf(dist / eps, beta)) / (1.f + noise_floor) + noise_floor;
  mask[i] = m;
}

__global__ void rowsum_w_kernel(const float* __restrict__ W, int F, int K, float* __restrict__ rowsum) {
  const int f = blockIdx.x;
  __shared__ float ws[32];
  float s = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) s += W[(int64_t)f * K + k];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? ws[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) rowsum[f] = s;
  }
}

struct LoadMaskT {  // B(n = t, k = atom) = mask[atom][t]
  static constexpr bool kContigK = false;
  const float* mask; int K, T;
  __device__ float operator()(int n, int k) const { return (n < T && k < K) ? __ldg(mask + (int64_t)k * T + n) : 0.f; }
};

// Y[c] = ((W . mask) / rowsum(W)) * X[c]     (onlineSpeechEnhancement.ipynb:429-431,440; gccNMFProcessor.py:267-269,209)
__global__ void __launch_bounds__(kReconThreads, 2)
wiener_apply_kernel(const float* __restrict__ mask, const float* __restrict__ W, const float* __restrict__ rowsumW,
                    const float2* __restrict__ X, int F, int T, int K, float2* __restrict__ Y, float* __restrict__ wiener) {
  float acc[RTM][RTN];
  const int m0 = blockIdx.y * RM, n0 = blockIdx.x * RN;
  LoadWRows a{W, F, K};
  LoadMaskT b{mask, K, T};
  gemm_simt_mainloop<float, RM, RN, RK, RTM, RTN>(acc, m0, n0, K, a, b);
#pragma unroll
  for (int i = 0; i < RTM; ++i) {
    const int m = gemm_row<RM, RTM, RN / RTN>(m0, i);
    if (m >= F) continue;
    const float rs = rowsumW[m];
#pragma unroll
    for (int j = 0; j < RTN; ++j) {
      const int n = gemm_col<RN, RTN, RN / RTN>(n0, j);
      if (n >= T) continue;
      const float w = acc[i][j] / rs;
      if (wiener) wiener[(int64_t)m * T + n] = w;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float2 x = X[((int64_t)c * F + m) * T + n];
        Y[((int64_t)c * F + m) * T + n] = float2{w * x.x, w * x.y};
      }
    }
  }
}

struct LoadHMaybeMasked {  // B(n = t, k = atom) = H[atom][c*T + t] (* mask[atom][t] when mask != NULL)
  static constexpr bool kContigK = false;
  const float* H; const float* mask; int K, T; int64_t ldh;
  __device__ float operator()(int n, int k) const {
    if (n >= T || k >= K) return 0.f;
    const float hv = __ldg(H + (int64_t)k * ldh + n);
    return mask ? hv * __ldg(mask + (int64_t)k * T + n) : hv;
  }
};

// Wiener-like filter with inferred coefficients (onlineSpeechEnhancement.ipynb:435-440), per channel c = blockIdx.z:
//   wiener[c] = (W . (H_c * mask)) / (W . H_c);   Y[c] = wiener[c] * X[c]          H (K, 2T), column c T + t
__global__ void __launch_bounds__(kReconThreads, 2)
wiener_apply_h_kernel(const float* __restrict__ mask, const float* __restrict__ W, const float* __restrict__ H, const float2* __restrict__ X,
                      int F, int T, int K, float2* __restrict__ Y, float* __restrict__ wiener) {
  const int c = blockIdx.z;
  float num[RTM][RTN], den[RTM][RTN];
  const int m0 = blockIdx.y * RM, n0 = blockIdx.x * RN;
  LoadWRows a{W, F, K};
  LoadHMaybeMasked masked{H + (int64_t)c * T, mask, K, T, (int64_t)2 * T}, plain{H + (int64_t)c * T, nullptr, K, T, (int64_t)2 * T};
  gemm_simt_mainloop<float, RM, RN, RK, RTM, RTN>(num, m0, n0, K, a, masked);
  gemm_simt_mainloop<float, RM, RN, RK, RTM, RTN>(den, m0, n0, K, a, plain);
#pragma unroll
  for (int i = 0; i < RTM; ++i) {
    const int m = gemm_row<RM, RTM, RN / RTN>(m0, i);
    if (m >= F) continue;
#pragma unroll
    for (int j = 0; j < RTN; ++j) {
      const int n = gemm_col<RN, RTN, RN / RTN>(n0, j);
      if (n >= T) continue;
      const float w = num[i][j] / den[i][j];
      const int64_t o = ((int64_t)c * F + m) * T + n;
      if (wiener) wiener[o] = w;
      const float2 x = X[o];
      Y[o] = float2{w * x.x, w * x.y};
    }
  }
}

bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

}  // namespace

// gcc_tc.cu
bool gccnmf_masked_recon_tc_supported(int S, int F, int T, int K);
extern "C" int gccnmf_masked_recon_planes(gccnmf_handle* h, const float* masks, const float* X, const float* W, const float* H, int S, int F, int T,
                                          int K, float* out, void* workspace, size_t workspace_bytes, void* stream);

extern "C" {

size_t gccnmf_phat_angspec_workspace_bytes(int F, int T, int D) {
  (void)F;
  if (T <= 0 || D <= 0) return 0;
  return align_up((size_t)((T + kAngT - 1) / kAngT) * D * sizeof(double), 256);
}

int gccnmf_phat_angspec(gccnmf_handle* h, const float* X, int F, int T, int x_is_coherence, const double* expJOmegaTau, int D, float* coherence,
                        double* angular, double* mean_angular, void* workspace, size_t workspace_bytes, void* stream) {
  GCCNMF_ENTER(h);
  GCCNMF_REQUIRE(h, F > 0 && T > 0, "phat_angspec: F and T must be positive");
  GCCNMF_REQUIRE(h, X != nullptr, "phat_angspec: NULL spectrogram");
  const bool need_ang = angular || mean_angular;
  if (need_ang) {
    GCCNMF_REQUIRE(h, expJOmegaTau != nullptr && D > 0, "phat_angspec: TDOA table required");
    if (D > kAngMaxD) return gccnmf_fail(h, GCCNMF_ERR_UNSUPPORTED, "phat_angspec: numTDOAs %d > %d", D, kAngMaxD);
  }
  double* tile_sums = nullptr;
  const int tiles = (T + kAngT - 1) / kAngT;
  if (mean_angular) {
    if (!workspace || workspace_bytes < gccnmf_phat_angspec_workspace_bytes(F, T, D))
      return gccnmf_fail(h, GCCNMF_ERR_WORKSPACE, "phat_angspec workspace too small");
    tile_sums = static_cast<double*>(workspace);
  }
  const int Dk = need_ang ? D : 0;
  const float2* Xc = reinterpret_cast<const float2*>(X);
  const double2* Ec = reinterpret_cast<const double2*>(expJOmegaTau);
  float2* Cc = reinterpret_cast<float2*>(coherence);
  if (Dk <= 32) GCCNMF_LAUNCH(h, (phat_angspec_kernel<1, 32>), tiles, kAngWarps * 32, 0, stream, Xc, F, T, x_is_coherence, Ec, Dk, Cc, angular, tile_sums);
  else if (Dk <= 64) GCCNMF_LAUNCH(h, (phat_angspec_kernel<2, 16>), tiles, kAngWarps * 32, 0, stream, Xc, F, T, x_is_coherence, Ec, Dk, Cc, angular, tile_sums);
  else GCCNMF_LAUNCH(h, (phat_angspec_kernel<4, 8>), tiles, kAngWarps * 32, 0, stream, Xc, F, T, x_is_coherence, Ec, Dk, Cc, angular, tile_sums);
  if (mean_angular) GCCNMF_LAUNCH(h, mean_tiles_kernel, (D + 63) / 64, 64, 0, stream, tile_sums, tiles, D, T, mean_angular);
  return GCCNMF_OK;
}

int gccnmf_tdoa_gccnmf(gccnmf_handle* h, const float* coherence, int F, int T, const double* E, int D, const float* W, int K,
                       float* values, int32_t* argmax, void* stream) {
  GCCNMF_ENTER(h);
  GCCNMF_REQUIRE(h, F > 0 && T > 0 && D > 0 && K > 0, "tdoa_gccnmf: dimensions must be positive");
  GCCNMF_REQUIRE(h, coherence && E && W, "tdoa_gccnmf: NULL pointer");
  GCCNMF_REQUIRE(h, (int64_t)T * D < (int64_t)1 << 31, "tdoa_gccnmf: T * D overflows int32");
  if (!values && !argmax) return GCCNMF_OK;
  if (argmax && !(is_pow2(D) && D >= 4 && D <= GN))
    return gccnmf_fail(h, GCCNMF_ERR_UNSUPPORTED, "tdoa_gccnmf: fused argmax needs numTDOAs a power of two in [4, %d] (got %d)", GN, D);
  const int N = T * D;
  LoadWAtoms a{W, K, F};
  LoadRealGCC b{reinterpret_cast<const float2*>(coherence), reinterpret_cast<const double2*>(E), F, T, D, N};
  dim3 grid((N + GN - 1) / GN, (K + GM - 1) / GM);
  if (argmax) {
    auto k = tdoa_gccnmf_kernel<true>;
    GCCNMF_LAUNCH(h, k, grid, kGccThreads, 0, stream, K, N, F, a, b, T, D, values, argmax);
  } else {
    auto k = tdoa_gccnmf_kernel<false>;
    GCCNMF_LAUNCH(h, k, grid, kGccThreads, 0, stream, K, N, F, a, b, T, D, values, argmax);
  }
  return GCCNMF_OK;
}

int gccnmf_coeff_mask(gccnmf_handle* h, const float* gccnmfs, int S, int K, int T, float* masks, int32_t* all_nan_flag, void* stream) {
  GCCNMF_ENTER(h);
  GCCNMF_REQUIRE(h, S > 0 && K > 0 && T > 0 && gccnmfs && masks, "coeff_mask: bad arguments");
  const int64_t KT = (int64_t)K * T;
  GCCNMF_LAUNCH(h, coeff_mask_kernel, (unsigned)((KT + 255) / 256), 256, 0, stream, gccnmfs, S, KT, masks, all_nan_flag);
  return GCCNMF_OK;
}

int gccnmf_argmax_mask(gccnmf_handle* h, const int32_t* argmax, int K, int T, const uint8_t* lut, int D, float* mask, void* stream) {
  GCCNMF_ENTER(h);
  GCCNMF_REQUIRE(h, K > 0 && T > 0 && D > 0 && argmax && lut && mask, "argmax_mask: bad arguments");
  const int64_t KT = (int64_t)K * T;
  GCCNMF_LAUNCH(h, argmax_mask_kernel, (unsigned)((KT + 255) / 256), 256, 0, stream, argmax, KT, lut, D, mask);
  return GCCNMF_OK;
}

int gccnmf_masked_recon_phase(gccnmf_handle* h, const float* masks, const float* X, const float* W, const float* H, int S, int F,
                              int T, int K, float* out, void* workspace, size_t workspace_bytes, void* stream) {
  GCCNMF_ENTER(h);
  GCCNMF_REQUIRE(h, S > 0 && F > 0 && T > 0 && K > 0 && masks && X && W && H && out, "masked_recon_phase: bad arguments");
  // tensor cores (plane GEMM over masked-H planes, gcc_tc.cu) when the caller provides the workspace and the shape is covered;
  // else the float32 SIMT GEMM below
  if (workspace && !h->force_simt_nmf && gccnmf_masked_recon_tc_supported(S, F, T, K))
    return gccnmf_masked_recon_planes(h, masks, X, W, H, S, F, T, K, out, workspace, workspace_bytes, stream);
  dim3 grid((T + RN - 1) / RN, (F + RM - 1) / RM, S * 2);
  GCCNMF_LAUNCH(h, masked_recon_kernel, grid, kReconThreads, 0, stream, masks, reinterpret_cast<const float2*>(X), W, H, F, T, K,
                reinterpret_cast<float2*>(out));
  return GCCNMF_OK;
}

int gccnmf_online_targets(gccnmf_handle* h, const double* angular, int D, int T, double* accumulated_max, int32_t* targets, void* stream) {
  GCCNMF_ENTER(h);
  GCCNMF_REQUIRE(h, angular && accumulated_max && targets && D > 0 && T > 0, "online_targets: bad arguments");
  GCCNMF_LAUNCH(h, cummax_time_kernel, (D + 63) / 64, 64, 0, stream, angular, D, T, accumulated_max);
  GCCNMF_LAUNCH(h, argmax_tdoa_kernel, (T + 127) / 128, 128, 0, stream, accumulated_max, D, T, targets);
  return GCCNMF_OK;
}

int gccnmf_atom_mask(gccnmf_handle* h, const int32_t* argmax, int K, int T, const int32_t* targets, float target_scalar, float epsilon,
                     int mode, float beta, float noise_floor, float* mask, void* stream) {
  GCCNMF_ENTER(h);
  GCCNMF_REQUIRE(h, argmax && mask && K > 0 && T > 0 && (mode == 0 || mode == 1), "atom_mask: bad arguments");
  const int64_t n = (int64_t)K * T;
  GCCNMF_LAUNCH(h, atom_mask_kernel, (unsigned)((n + 255) / 256), 256, 0, stream, argmax, K, T, targets, 1, target_scalar, epsilon, mode,
                beta, noise_floor, mask);
  return GCCNMF_OK;
}

size_t gccnmf_wiener_apply_workspace_bytes(int F) { return F > 0 ? align_up((size_t)F * sizeof(float), 256) : 0; }

int gccnmf_wiener_apply(gccnmf_handle* h, const float* mask, const float* W, const float* X, int F, int T, int K, float* Y,
                        float* wiener, void* workspace, size_t workspace_bytes, void* stream) {
  GCCNMF_ENTER(h);
  GCCNMF_REQUIRE(h, mask && W && X && Y && F > 0 && T > 0 && K > 0, "wiener_apply: bad arguments");
  if (!workspace || workspace_bytes < gccnmf_wiener_apply_workspace_bytes(F)) return gccnmf_fail(h, GCCNMF_ERR_WORKSPACE, "wiener_apply workspace too small");
  float* rowsum = static_cast<float*>(workspace);
  GCCNMF_LAUNCH(h, rowsum_w_kernel, F, 128, 0, stream, W, F, K, rowsum);
  dim3 grid((T + RN - 1) / RN, (F + RM - 1) / RM);
  GCCNMF_LAUNCH(h, wiener_apply_kernel, grid, kReconThreads, 0, stream, mask, W, rowsum, reinterpret_cast<const float2*>(X), F, T, K,
                reinterpret_cast<float2*>(Y), wiener);
  return GCCNMF_OK;
}

// Y (2, F, T) c64 = wiener * X with wiener (2, F, T) f32 = (W . (H_c * mask)) / (W . H_c) per channel (ipynb:435-440):
// the numInferenceIterations > 0 branch; H (K, 2T) holds channel c in columns [c T, (c + 1) T).
int gccnmf_wiener_apply_h(gccnmf_handle* h, const float* mask, const float* W, const float* H, const float* X, int F, int T, int K, float* Y,
                          float* wiener, void* stream) {
  GCCNMF_ENTER(h);
  GCCNMF_REQUIRE(h, mask && W && H && X && Y && F > 0 && T > 0 && K > 0, "wiener_apply_h: bad arguments");
  dim3 grid((T + RN - 1) / RN, (F + RM - 1) / RM, 2);
  GCCNMF_LAUNCH(h, wiener_apply_h_kernel, grid, kReconThreads, 0, stream, mask, W, H, reinterpret_cast<const float2*>(X), F, T, K,
                reinterpret_cast<float2*>(Y), wiener);
  return GCCNMF_OK;
}

}  // extern "C"
