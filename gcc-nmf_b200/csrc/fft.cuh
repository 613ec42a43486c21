// Shared-memory FFT building blocks used by the STFT / iSTFT kernels (stft.cu) and the real-time block path (rt.cu).
#pragma once
#include <cuda_runtime.h>

constexpr int kFftThreads = 256;

__device__ __forceinline__ unsigned bitrev(unsigned x, int bits) { return __brev(x) >> (32 - bits); }

// In-place radix-2 decimation-in-time butterflies on bit-reversed input held in shared memory.
// tw[j] = exp(-2 pi i j / n) for j < n/2; `inverse` conjugates the twiddles.
template <typename T2, typename T>
__device__ __forceinline__ void fft_inplace(T2* x, const T2* __restrict__ tw, int n, int log2n, bool inverse) {
  for (int s = 1; s <= log2n; ++s) {
    const int half = 1 << (s - 1);
    const int tw_stride = n >> s;
    for (int j = threadIdx.x; j < n / 2; j += blockDim.x) {
      const int pos = j & (half - 1);
      const int i0 = ((j >> (s - 1)) << s) + pos;
      const int i1 = i0 + half;
      T2 w = tw[pos * tw_stride];
      if (inverse) w.y = -w.y;
      const T2 a = x[i0], b = x[i1];
      const T tr = w.x * b.x - w.y * b.y;
      const T ti = w.x * b.y + w.y * b.x;
      x[i0] = T2{a.x + tr, a.y + ti};
      x[i1] = T2{a.x - tr, a.y - ti};
    }
    __syncthreads();
  }
}

