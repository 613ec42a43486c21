// Shared host-side plumbing for the gccnmf_b200 C ABI: handle, error capture, launch counting.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/gccnmf_b200.h"

struct gccnmf_handle {
  int device = 0;
  int sm_count = 148;
  int64_t launches = 0;
  bool force_simt_nmf = false;   // GCCNMF_NMF_PATH=simt: float32 SIMT contractions instead of the tcgen05 plane GEMM
  bool nmf_pdl = true;           // programmatic dependent launch between the kernels of a KL-NMF iteration
  int wh_tile = 0;               // diagnostics: tile width of the W.H contractions (0 = planned)
  bool argmax_refine_shared = true;   // exact float64 refinement of near-tie argmax decisions: E staged in shared memory (0 = one warp per pair from L2)
  int gemm_streaming = 0;        // st.global.cs for the k-split partials of the W-update numerator (diagnostics)
  bool argmax_persistent = true; // all-TDOA argmax GEMM as one persistent CTA per SM with double-buffered TMEM accumulators (0: one CTA per tile)
  int gemm_preload = 1;          // plane GEMM epilogues that fetch their global operands during the main loop: bit 0 ratio (W.H), bit 1 H update
  int gemm_pair = -1;            // plane GEMM on cta_group::2 CTA pairs (256-row MMAs): -1 where the call site prefers it, 0 never, 1 wherever possible
  int wh_split2 = 0;             // W.H contractions as plain 128 x 208 tiles split in two k-halves summed inside (1, 1, 2) clusters (0: dual-N 104-column tiles)
  bool w_cluster_reduce = true;  // W-update numerator: k-splits summed inside (1, 1, splits) clusters through distributed shared memory (0: k-split slabs)
  // set by gccnmf_klnmf_tma_step_pull around the contractions of a sharded iteration (pull exchange): where the row sums of G and the
  // W-update numerator go (this rank's symmetric buffer) and whom the numerator contraction signals when its last CTA is done
  float* xchg_rowsum = nullptr;
  float* xchg_numer = nullptr;
  unsigned* xchg_done = nullptr;
  unsigned* xchg_counters[8] = {nullptr};
  int xchg_world = 0;
  int pull_force_pack = 0;       // pull exchange: always go through the pack kernel (diagnostics)
  int mc_light_signal = 1;       // sharded runs: arrival signal of the numerator pack as device-scope fence + relaxed multimem.red (0: MEMBAR.SYS + release)
  int l2_persist = 0;            // KL-NMF loop: launch-attribute L2 access-policy window (persisting) over G^T: 1 = float32 master, 2 = master + planes
  const void* l2_window_base = nullptr;   // set by the KL-NMF loop while it runs
  size_t l2_window_bytes = 0;
  bool l2_limit_set = false;
  int gemm_cluster = -1;         // diagnostics: force the plane GEMM cluster shape 10 CN + CM (11 = no cluster); -1 = automatic
  unsigned long long* debug_timing = nullptr;   // diagnostics (gccnmf_debug_timing): CTA stamps of every plane GEMM
  size_t debug_timing_cursor = 0;
  struct gccnmf_tmap_cache* tmaps = nullptr;   // TMA tensor maps, keyed by (buffer, shape, box)
  std::string last_error;
  // twiddle tables e^{-2 pi i j / n}, j < n/2, float64 and float32, cached per FFT size
  static constexpr int kMaxPlans = 8;
  int plan_n[kMaxPlans] = {0};
  double* plan_tw64[kMaxPlans] = {nullptr};
  float* plan_tw32[kMaxPlans] = {nullptr};
};

inline int gccnmf_fail(gccnmf_handle* h, int status, const char* fmt, ...) __attribute__((format(printf, 3, 4)));
#include <cstdarg>
inline int gccnmf_fail(gccnmf_handle* h, int status, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->last_error = buf;
  return status;
}

#define GCCNMF_CHECK_CUDA(h, expr)                                                              \
  do {                                                                                          \
    cudaError_t err__ = (expr);                                                                 \
    if (err__ != cudaSuccess)                                                                   \
      return gccnmf_fail((h), GCCNMF_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                  \
                         cudaGetErrorString(err__), __FILE__, __LINE__);                        \
  } while (0)

// Every kernel launch goes through this so that launch errors are captured and counted.
#define GCCNMF_LAUNCH(h, kernel, grid, block, smem, stream, ...)                                \
  do {                                                                                          \
    kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__);                   \
    cudaError_t err__ = cudaGetLastError();                                                     \
    if (err__ != cudaSuccess)                                                                   \
      return gccnmf_fail((h), GCCNMF_ERR_CUDA, "launch of %s failed: %s (%s:%d)", #kernel,      \
                         cudaGetErrorString(err__), __FILE__, __LINE__);                        \
    (h)->launches++;                                                                            \
  } while (0)

#define GCCNMF_REQUIRE(h, cond, ...)                                                            \
  do {                                                                                          \
    if (!(cond)) return gccnmf_fail((h), GCCNMF_ERR_INVALID_ARGUMENT, __VA_ARGS__);             \
  } while (0)

// Every ABI entry that takes a handle starts here: NULL check + make the handle's device current (a process may hold
// handles on several devices; kernel attributes and occupancy caches below are kept per device index).
constexpr int kGccnmfMaxDevices = 64;
inline int gccnmf_enter(gccnmf_handle* h) {
  if (!h) return GCCNMF_ERR_INVALID_ARGUMENT;
  const cudaError_t err = cudaSetDevice(h->device);
  if (err != cudaSuccess) return gccnmf_fail(h, GCCNMF_ERR_CUDA, "cudaSetDevice(%d) failed: %s", h->device, cudaGetErrorString(err));
  return GCCNMF_OK;
}
#define GCCNMF_ENTER(h)                                                                         \
  do {                                                                                          \
    if (int st__ = gccnmf_enter(h)) return st__;                                                \
  } while (0)
// Per-device once-flag for cudaFuncSetAttribute and friends (function-local `static DeviceFlags configured;`).
struct DeviceFlags {
  bool done[kGccnmfMaxDevices] = {false};
  bool& operator()(const gccnmf_handle* h) { return done[h->device % kGccnmfMaxDevices]; }
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Carves aligned sub-buffers out of a caller-owned workspace.
struct WorkspaceCarver {
  char* base;
  size_t size, used = 0;
  WorkspaceCarver(void* p, size_t n) : base(static_cast<char*>(p)), size(n) {}
  template <typename T>
  T* take(size_t count) {
    used = align_up(used, 256);
    T* p = reinterpret_cast<T*>(base + used);
    used += count * sizeof(T);
    return p;
  }
  bool ok() const { return base != nullptr && used <= size; }
};

void gccnmf_tmap_cache_free(gccnmf_handle* h);
int gccnmf_get_twiddles(gccnmf_handle* h, int n, const double** tw64, const float** tw32);
