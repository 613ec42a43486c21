/*
 * gccnmf_b200 -- C ABI of the B200-native GCC-NMF separation hot path.
 *
 * The reference (seanwood/gcc-nmf) has no FFI: its boundary is the Python call surface of
 * gccNMF/gccNMFFunctions.py, gccNMF/librosaSTFT.py and gccNMF/realtime/gccNMFProcessor.py.
 * Each entry point below is the device-side replacement of one of those Python functions
 * (cited per function as file:line relative to the reference root) and is what a ctypes
 * binding in the reference would call.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *  - the caller owns all buffers (including workspaces, sized by the *_workspace_bytes helpers);
 *    the library borrows them for the duration of the call and never frees them;
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy default
 *    stream) and returns 0 on success or a negative gccnmf_status; no C++ exception crosses the ABI;
 *  - array layouts are the reference's numpy C-order layouts: spectrograms (channel, F, T) with T
 *    contiguous, W (F, K), H (K, 2T), masks (S, K, T), signals (S, 2, n);
 *  - complex64 is interleaved (re, im) float pairs, complex128 interleaved double pairs;
 *  - a handle is bound to one device and is not thread-safe (use one per host thread / stream).
 */
#ifndef GCCNMF_B200_H_
#define GCCNMF_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define GCCNMF_API __attribute__((visibility("default")))
#else
#define GCCNMF_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define GCCNMF_ABI_VERSION 2

typedef struct gccnmf_handle gccnmf_handle;

typedef enum gccnmf_status {
  GCCNMF_OK = 0,
  GCCNMF_ERR_INVALID_ARGUMENT = -1, /* librosaSTFT.ParameterError / ValueError in the reference   */
  GCCNMF_ERR_CUDA = -2,             /* a CUDA runtime call failed; text in gccnmf_last_error       */
  GCCNMF_ERR_WORKSPACE = -3,        /* workspace pointer NULL or too small                          */
  GCCNMF_ERR_UNSUPPORTED = -4,      /* shape outside what the kernels were built for               */
  GCCNMF_ERR_NO_DEVICE = -5         /* no CUDA device: there is NO CPU fallback                     */
} gccnmf_status;

/* ---- handle ------------------------------------------------------------------------------- */
GCCNMF_API int gccnmf_abi_version(void);
/* Binds to `device`; fails with GCCNMF_ERR_NO_DEVICE when no GPU is visible. */
GCCNMF_API int gccnmf_create(gccnmf_handle** out, int device);
GCCNMF_API int gccnmf_destroy(gccnmf_handle* h);
/* Text of the last error recorded on this handle (never NULL; h may be NULL for create errors). */
GCCNMF_API const char* gccnmf_last_error(const gccnmf_handle* h);
GCCNMF_API const char* gccnmf_status_string(int status);
/* Count of kernels this handle has launched since creation (bench.py's `gpu_launches`). */
GCCNMF_API int64_t gccnmf_launch_count(const gccnmf_handle* h);
/* Options (A/B switches between sm_100a code paths of this library; all ranks of a sharded run must use the same ones so
 * that W stays bit-identical).  Defaults in brackets.
 *   "force_simt_nmf" [0]        KL-NMF contractions on the float32 SIMT kernels even where the tcgen05 plane GEMM applies
 *   "nmf_pdl" [1]               programmatic dependent launch between the kernels of a KL-NMF iteration
 *   "gemm_cluster" [-1]         plane GEMM cluster shape 10 CN + CM (11, 12, 21, 22) instead of the automatic choice
 *   "argmax_refine_shared" [1]  float64 refinement of near-tie argmax decisions with E staged in shared memory
 *   "argmax_persistent" [1]     all-TDOA argmax GEMM as one persistent CTA per SM (double-buffered TMEM); 0 = one CTA per tile
 *   "wh_tile" [0]               tile width of the W.H contractions (104 / 112 / 128 / 256) instead of the planned one
 *   "gemm_pair" [-1]            plane GEMM on cta_group::2 CTA pairs: -1 where a call site prefers it, 0 never, 1 wherever possible
 *                               (bit-identical results either way; measured no faster)
 *   "l2_persist" [0]            KL-NMF loop: persisting L2 access-policy window over G^T (1 = float32 master, 2 = master + planes)
 *   "wh_split2" [0]             W.H contractions as plain 128 x 208 tiles with the contraction split in two halves that a
 *                               (1, 1, 2) cluster sums through distributed shared memory (measured: parity with the default)
 *   "w_cluster_reduce" [1]      W-update numerator: k-splits summed inside (1, 1, splits) clusters where every cluster of the
 *                               launch is resident at once (else k-split slabs)
 *   "mc_light_signal" [1]       sharded runs: arrival signal as device-scope fence + relaxed red (0: MEMBAR.SYS + release)
 *   "pull_force_pack" [0]       pull exchange: always through the pack kernel (diagnostics)
 *   "gemm_preload" [1]          bit 0: the W.H ratio epilogue fetches V during the main loop
 *   "gemm_streaming" [0]        st.global.cs / ld.global.cs for the k-split partials of the W-update numerator */
GCCNMF_API int gccnmf_set_option(gccnmf_handle* h, const char* name, int value);

/* ---- a1: STFT  (gccNMF/librosaSTFT.py:20-181 via gccNMFFunctions.py:61-67) ------------------ */
/* Frame count 1 + (num_samples - n_fft) / hop (librosaSTFT.py:425); <1 -> GCCNMF_ERR_INVALID_ARGUMENT. */
GCCNMF_API int gccnmf_stft_num_frames(int64_t num_samples, int n_fft, int hop);
/*
 * samples (channels, num_samples) f32 with row stride `sample_stride`; window (n_fft) f64 on device;
 * X (channels, F, T) c64.  Computation is float64 like the reference (window f64 * frame, double FFT)
 * and rounded once to complex64.  conjugate != 0 reproduces librosaSTFT.py:179 (offline path);
 * conjugate == 0 is numpy.fft.rfft (online / real-time path, onlineSpeechEnhancement.ipynb:410).
 * V (F, channels*T) f32 = |X| with the channels concatenated in time (runGCCNMF.py:40); may be NULL.
 * channels must be 1 or 2; n_fft a power of two in [32, 4096].
 */
GCCNMF_API int gccnmf_stft(gccnmf_handle* h, const float* samples, int64_t sample_stride, int channels,
                int64_t num_samples, const double* window, int n_fft, int hop, int conjugate,
                float* X, float* V, void* stream);

/* ---- a9: iSTFT + overlap-add  (librosaSTFT.py:183-286 via gccNMFFunctions.py:153-163) ------- */
/* Output length per signal: n_fft + hop (T-1) - (center ? n_fft : 0). */
GCCNMF_API int64_t gccnmf_istft_length(int n_fft, int hop, int T, int center);
GCCNMF_API size_t gccnmf_istft_workspace_bytes(int batch, int n_fft, int T);
/*
 * spec (batch, F, T) c64 -> y (batch, length) f32.  Per frame: Hermitian rebuild from conj(col)
 * (librosaSTFT.py:278), single-precision inverse FFT (:279), real part times window (f64),
 * sequential float32 overlap-add in frame order (:281), optional centre trim (:283-284), times
 * `gain` (gccNMFFunctions.py:155,163).  conjugate == 0 skips the conj (numpy.fft.irfft convention).
 */
GCCNMF_API int gccnmf_istft_ola(gccnmf_handle* h, const float* spec, int batch, int n_fft, int hop, int T,
                     const double* window, float gain, int center, int conjugate, float* y,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ---- a2: KL-NMF  (gccNMFFunctions.py:69-83) -------------------------------------------------- */
GCCNMF_API size_t gccnmf_klnmf_workspace_bytes(int F, int T2, int K);
/* 1 when the contractions of this shape run on tcgen05 (TMA-fed plane GEMM over bf16 hi/lo operand planes: K % 8 == 0, K >= 32,
 * F, T2 >= 128), 0 when they run on the float32 SIMT kernels (small or odd shapes, or GCCNMF_NMF_PATH=simt). */
GCCNMF_API int gccnmf_klnmf_uses_tensor_cores(const gccnmf_handle* h, int F, int T2, int K);
/*
 * V (F, T2) f32 non-negative; W (F, K) and H (K, T2) f32 hold the initial values on entry (the
 * seeded numpy draw of gccNMFFunctions.py:70-73 is made on the host) and the result on return.
 * Runs `iterations` passes of lines :76-81 (H update, W update on the recomputed W.H, unit-L2 atoms).
 * update_W == 0 runs only the H update (:76) with a fixed dictionary (the undefined
 * inferCoefficientsKLNMF of onlineSpeechEnhancement.ipynb:433).
 */
GCCNMF_API int gccnmf_klnmf(gccnmf_handle* h, const float* V, int F, int T2, float* W, float* H, int K,
                 int iterations, float sparsity_alpha, float epsilon, int update_W,
                 void* workspace, size_t workspace_bytes, void* stream);
/*
 * Frame-sharded dictionary learning (multi-GPU).  A rank holds the columns V_s
 * (F, T2s), H_s (K, T2s) of its frames and a replica of W.  The loop of gccNMFFunctions.py:75-81 becomes
 *
 *   gccnmf_klnmf_begin(...)                       once (operand re-layout for the tensor-core path)
 *   for it in range(iterations):
 *       gccnmf_klnmf_step_numer(..., it, numer)   H_s update (:76); numer = [ (V_s/(W H_s)).H_s^T (F*K) | rowsum(H_s) (K) ]
 *       all-reduce(sum) of numer over the ranks   ONE collective of F*K + K floats per iteration
 *       gccnmf_klnmf_step_apply(..., numer)       W *= numer / rowsum (:77), unit-L2 atoms (:79-80), H_s *= norms (:81)
 *   gccnmf_klnmf_end(..., iterations)             materialises the last (lazily applied) H_s rescale
 *
 * All state lives in W, H and the caller-owned workspace (gccnmf_klnmf_workspace_bytes), which must not be
 * touched between begin and end.  Every rank computes a bit-identical W from the same all-reduced numer.
 */
GCCNMF_API int gccnmf_klnmf_begin(gccnmf_handle* h, const float* V, int F, int T2, const float* W, const float* H,
                       int K, void* workspace, size_t workspace_bytes, void* stream);
GCCNMF_API int gccnmf_klnmf_step_numer(gccnmf_handle* h, const float* V, int F, int T2, const float* W, float* H,
                            int K, float sparsity_alpha, float epsilon, int iteration, float* numer,
                            void* workspace, size_t workspace_bytes, void* stream);
GCCNMF_API int gccnmf_klnmf_step_apply(gccnmf_handle* h, int F, int T2, float* W, float* H, int K,
                            const float* numer, void* workspace, size_t workspace_bytes, void* stream);
/*
 * Fused collective variant of step_apply for NVSwitch systems: every rank's step_numer wrote its partial into the
 * SAME offset of a symmetric buffer bound to an NVLink multicast object; `numer_multicast` is the multicast address
 * of that buffer.  After a cross-rank barrier the W update reads each numerator / row-sum word with
 * multimem.ld_reduce.add.f32 -- the sum over ranks is formed inside the switch (NVLS) while the kernel streams it,
 * so there is no separate all-reduce kernel and no second pass over the 2 MB.  Every rank receives the same
 * switch-reduced value, hence a bit-identical W.  Tensor-core path only (GCCNMF_ERR_UNSUPPORTED otherwise).
 */
GCCNMF_API int gccnmf_klnmf_step_apply_multimem(gccnmf_handle* h, int F, int T2, float* W, float* H, int K,
                                     const float* numer_multicast, void* workspace, size_t workspace_bytes,
                                     void* stream);
/* One iteration of a frame-sharded run with the cross-rank exchange fused into the kernels: the numerator pack signals "this rank
 * is complete" to every rank with one multimem.red on an arrival counter, the W-update kernel waits on its own copy of the counter
 * and reads the sum over ranks with multimem.ld_reduce (formed inside the NVSwitch) -- no host-launched barrier, no all-reduce.
 * numer_local / counter_local: this rank's addresses inside its symmetric buffer ((F*K + K) floats + one uint32 per buffer);
 * numer_multicast / counter_multicast: the NVLink multicast addresses of the same offsets; arrivals_expected = world size x the
 * number of times this buffer has been used so far (including this one).  Double-buffer by iteration parity. */
GCCNMF_API int gccnmf_klnmf_step_multimem(gccnmf_handle* h, const float* V, int F, int T2, float* W, float* H, int K,
                               float sparsity_alpha, float epsilon, int iteration, float* numer_local,
                               const float* numer_multicast, const uint32_t* counter_local, uint32_t* counter_multicast,
                               uint32_t arrivals_expected, void* workspace, size_t workspace_bytes, void* stream);
/* The same iteration with a TWO-SHOT exchange: after the pack every rank sums only its 1 / world slice of the numerator with
 * multimem.ld_reduce and multicasts the sum into the `reduced` buffer of every rank (multimem.st); the W update reads its local
 * `reduced` copy once the second arrival counter is complete.  Link traffic per GPU and iteration: one numerator out, one in, for any
 * world size (the one-shot form above makes every GPU serve `world` numerators).  numer_* / reduced_*: two symmetric buffers of
 * (F*K + K) floats (local and multicast addresses); counters_*: two consecutive uint32 in symmetric memory (pack arrivals, slice
 * arrivals), zero before the first iteration; arrivals_expected = world x (iteration + 1).  No double buffering needed: a rank
 * re-packs only after its own W update, which waits for every rank's slice. */
GCCNMF_API int gccnmf_klnmf_step_multimem2(gccnmf_handle* h, const float* V, int F, int T2, float* W, float* H, int K,
                                float sparsity_alpha, float epsilon, int iteration, int rank, int world, float* numer_local,
                                const float* numer_multicast, const float* reduced_local, float* reduced_multicast,
                                const uint32_t* counters_local, uint32_t* counters_multicast, uint32_t arrivals_expected,
                                void* workspace, size_t workspace_bytes, void* stream);
/* PULL exchange (the default of the sharded pipeline): nothing is pushed over the links and nothing is reduced in the switch.  The
 * numerator contraction writes this rank's (F, K) partial straight into its symmetric buffer and its last CTA adds 1 to every rank's
 * arrival counter (device-scope fence + relaxed red: the published data is local, peers fetch it through this GPU's L2); then
 *   two_shot = 2: the exchange happens INSIDE the W update, tile by tile: the CTA that owns a 32 x 128 tile of U sums this rank's
 *                 k-split slabs for it, publishes the tile in the symmetric buffer, flags it on every rank, waits for the same tile
 *                 of the other ranks and reads them with plain peer loads -- five launches per iteration, like the single-GPU loop;
 *   two_shot = 0: every rank's W update reads all ranks' partials with plain peer loads, added in rank order;
 *   two_shot = 1: each rank first sums its 1 / world slice that way into its own buffer, signals, and the W updates fetch each word
 *                 from its owner -- one numerator in each direction per GPU for any world size.
 * Row sums of G are read from every rank's slots.  bases: HOST array of `world` device pointers, every rank's buffer as mapped in
 * this process (gccnmf_klnmf_pull_buffer_floats(F, layout_T2, K) floats each, zero before the first iteration; layout_T2 = the
 * largest 2T over the ranks; epoch = iterations earlier runs executed on this buffer: the arrival counters keep counting).  Needs the
 * cluster-reduced numerator contraction for direct = 1 (see gccnmf_klnmf_pull_supported). */
GCCNMF_API int64_t gccnmf_klnmf_pull_buffer_floats(int F, int layout_T2, int K);
/* Bit mask.  0: gccnmf_klnmf_step_pull does not cover this shard shape on this device.  Bit 0: it does, through the pack kernel
 * (k-split slabs and row-sum slots summed into the symmetric buffer, then the signal).  Bit 1: also in the direct form (the numerator
 * contraction sums its k-splits inside clusters, writes the buffer itself and signals from its last CTA): `direct` of step_pull must be
 * the same on every rank, 1 only if every rank has this bit.  Bit 2: form 2 (exchange inside the W update) is available: every tile
 * CTA of the W update is resident at once. */
GCCNMF_API int gccnmf_klnmf_pull_supported(gccnmf_handle* h, int F, int T2, int K);
GCCNMF_API int gccnmf_klnmf_step_pull(gccnmf_handle* h, const float* V, int F, int T2, float* W, float* H, int K,
                           float sparsity_alpha, float epsilon, int iteration, int64_t epoch, int rank, int world,
                           void* const* bases, int layout_T2, int two_shot, int direct, void* workspace, size_t workspace_bytes,
                           void* stream);
GCCNMF_API int gccnmf_klnmf_end(gccnmf_handle* h, int F, int T2, float* W, float* H, int K, int iterations_done,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ---- a3 + a4: PHAT coherence and angular spectrogram  (runGCCNMF.py:44, gccNMFFunctions.py:85-92) */
/*
 * X (2, F, T) c64; expJOmegaTau (F, D) c128 = exp(-2 pi i f tau) built on the host in float64
 * exactly as gccNMFFunctions.py:89.  coherence (F, T) c64 = X0 conj(X1) / |X0| / |X1| (unguarded,
 * 0/0 -> NaN); angular (D, T) f64 = sum_f Re(coherence * E); mean_angular (D) f64 = mean over T
 * (runGCCNMF.py:46).  coherence, angular and mean_angular may each be NULL.
 * x_is_coherence != 0: X is an already-normalised (F, T) coherence (the argument
 * gccNMFFunctions.getAngularSpectrogram takes, :85) and is used as is.
 */
GCCNMF_API size_t gccnmf_phat_angspec_workspace_bytes(int F, int T, int D);
GCCNMF_API int gccnmf_phat_angspec(gccnmf_handle* h, const float* X, int F, int T, int x_is_coherence,
                        const double* expJOmegaTau, int D, float* coherence, double* angular, double* mean_angular,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ---- a6 / a10 / a11: GCC-NMF per TDOA  (gccNMFFunctions.py:118-135; offlineSpeechEnhancement.ipynb:444-450) */
/*
 * gccnmf[d, k, t] = sum_f Re(coherence[f, t] * E[f, d]) * W[f, k] accumulated in float64
 * (the reference contracts in complex128 / float64).  E (F, D) c128 holds either the selected
 * target TDOA columns (a6: D = number of targets) or all hypothesis TDOAs (a10/a11).
 * values (D, K, T) f32 (a6's output dtype, gccNMFFunctions.py:131) and/or argmax (K, T) int32 over d
 * with numpy.argmax semantics (first maximum; NaN counts as maximum).  Either may be NULL.
 */
GCCNMF_API int gccnmf_tdoa_gccnmf(gccnmf_handle* h, const float* coherence, int F, int T, const double* E,
                       int D, const float* W, int K, float* values, int32_t* argmax, void* stream);

/*
 * a10 / a11 fast path: argmax over ALL hypothesis TDOAs without materialising (K, D, T) float64.
 * For D a power of two in [8, 128], K % 8 == 0, K >= 64 the contraction runs as a tcgen05 GEMM over bf16 hi/lo planes of
 * Re(C E) (persistent kernel, two TMEM accumulators) whose epilogue keeps the best / second-best value per (atom, frame) and the
 * candidates within the margin; every decision whose margin is inside the GEMM's worst-case
 * error is recomputed exactly in float64, so the result equals gccnmf_tdoa_gccnmf's argmax (the
 * reference's numpy.argmax on float64) on every input.  Other shapes run the float64 kernel directly.
 * *overflow_flag (device int32, may be NULL) receives the number of refined decisions; if it exceeds
 * gccnmf_tdoa_argmax_refine_capacity(K, T) the caller must recompute with gccnmf_tdoa_gccnmf.
 */
GCCNMF_API size_t gccnmf_tdoa_argmax_workspace_bytes(int F, int T, int D, int K);
GCCNMF_API int gccnmf_tdoa_argmax_refine_capacity(int K, int T);
GCCNMF_API int gccnmf_tdoa_argmax(gccnmf_handle* h, const float* coherence, int F, int T, const double* E, int D,
                       const float* W, int K, int32_t* argmax, int32_t* overflow_flag, void* workspace,
                       size_t workspace_bytes, void* stream);

/* ---- a7: coefficient masks  (gccNMFFunctions.py:137-143; offlineSpeechEnhancement.ipynb:466-472) */
/* nanargmax over S -> one-hot (S, K, T) f32.  *all_nan_flag (device int32, may be NULL) is set to 1
 * when some (k, t) is NaN for every target (numpy.nanargmax raises ValueError there). */
GCCNMF_API int gccnmf_coeff_mask(gccnmf_handle* h, const float* gccnmfs, int S, int K, int T, float* masks,
                      int32_t* all_nan_flag, void* stream);
/* mask[k, t] = lut[argmax[k, t]] with lut (D) u8 built on the host in float64 from
 * |tdoa[argmax] - tdoa[target]| < window (ipynb:468-471). */
GCCNMF_API int gccnmf_argmax_mask(gccnmf_handle* h, const int32_t* argmax, int K, int T, const uint8_t* lut,
                       int D, float* mask, void* stream);

/* ---- a11 / a13: online localisation, atom masks, Wiener-like filter ------------------------------
 * (notebooks/onlineSpeechEnhancement.ipynb:406-447, lowLatencySpeechEnhancement.ipynb:511-584,
 *  realtime/gccNMFProcessor.py:201-231,259-269).  The notebooks' frame loop carries one piece of state, the
 * accumulated maximum of the GCC-PHAT angular spectrum; it is a prefix maximum over time, so all frames are
 * processed in one batch. */
/* accumulated_max (D, T) f64 = running max over frames of angular (D, T); targets (T) i32 = its argmax over TDOA (:416-417). */
GCCNMF_API int gccnmf_online_targets(gccnmf_handle* h, const double* angular, int D, int T, double* accumulated_max,
                          int32_t* targets, void* stream);
/* mask (K, T) f32 from argmax (K, T) and the target TDOA index (per frame: targets (T) i32; or targets == NULL and
 * one float target_scalar, the Theano shared scalar of gccNMFProcessor.py:196).
 * mode 0 boxcar |argmax - target| < epsilon (ipynb:423-425; gccNMFProcessor.py:263);
 * mode 1 window exp(-(|argmax - target| / epsilon)^beta) / (1 + noise_floor) + noise_floor (gccNMFProcessor.py:265). */
GCCNMF_API int gccnmf_atom_mask(gccnmf_handle* h, const int32_t* argmax, int K, int T, const int32_t* targets,
                     float target_scalar, float epsilon, int mode, float beta, float noise_floor, float* mask,
                     void* stream);
/* Y (2, F, T) c64 = wiener * X with wiener (F, T) f32 = (W . mask) / rowsum(W) (ipynb:429-431,440;
 * gccNMFProcessor.py:267-269,209); wiener may be NULL when only Y is wanted. */
GCCNMF_API size_t gccnmf_wiener_apply_workspace_bytes(int F);
GCCNMF_API int gccnmf_wiener_apply(gccnmf_handle* h, const float* mask, const float* W, const float* X, int F, int T,
                        int K, float* Y, float* wiener, void* workspace, size_t workspace_bytes, void* stream);

/* ---- a8: masked reconstruction with mixture phase  (gccNMFFunctions.py:145-151) -------------- */
/* out[s, c] = (W . (H[:, c*T:(c+1)*T] * masks[s])) * exp(i angle(X[c])) ; out (S, 2, F, T) c64.
 * With a workspace of gccnmf_masked_recon_workspace_bytes the S x 2 products run on the tensor cores (the plane GEMM of the
 * KL-NMF loop over masked-H planes: 3 bf16 products per product, ~3e-6 relative); with workspace == NULL, or a shape the
 * tensor-core path does not cover, on the float32 SIMT kernel. */
GCCNMF_API size_t gccnmf_masked_recon_workspace_bytes(int S, int F, int T, int K);
GCCNMF_API int gccnmf_masked_recon_phase(gccnmf_handle* h, const float* masks, const float* X, const float* W,
                              const float* H, int S, int F, int T, int K, float* out, void* workspace,
                              size_t workspace_bytes, void* stream);

/* The numInferenceIterations > 0 branch of the notebooks' frame loop (onlineSpeechEnhancement.ipynb:433-440): with
 * H (K, 2T) f32 the inferred coefficients (channel c in columns [c T, (c + 1) T)),
 * wiener (2, F, T) f32 = (W . (H_c * mask)) / (W . H_c) and Y (2, F, T) c64 = wiener * X; wiener may be NULL. */
GCCNMF_API int gccnmf_wiener_apply_h(gccnmf_handle* h, const float* mask, const float* W, const float* H, const float* X, int F,
                          int T, int K, float* Y, float* wiener, void* stream);

/* ---- the whole offline path as one call ------------------------------------------------------------------
 * runGCCNMF.py:36-52 (num_targets >= 1: separation into num_targets sources) or the enhancement flow of
 * notebooks/offlineSpeechEnhancement.ipynb cells 12-41 (num_targets == 0: one target, all-TDOA argmax mask), every stage
 * enqueued on `stream` without a host synchronisation in between: the target TDOAs are picked by a device kernel with the
 * semantics of scipy.signal.argrelmax + the numSources largest peaks (gccNMFFunctions.py:94-116) and stay on the device. */
typedef struct gccnmf_pipeline_config {
  int window_size, hop_size;      /* N (= n_fft), hop                                                               */
  int num_tdoas, num_atoms;       /* D, K                                                                           */
  int num_iterations;             /* KL-NMF iterations                                                              */
  int num_targets;                /* S >= 1: separation; 0: enhancement (one target)                               */
  float sparsity_alpha, epsilon;  /* gccNMFFunctions.py:69                                                          */
  float target_window_seconds;    /* enhancement: |tdoa - tdoa[target]| < this (ipynb:468-471)                      */
} gccnmf_pipeline_config;
GCCNMF_API size_t gccnmf_pipeline_workspace_bytes(const gccnmf_pipeline_config* cfg, int64_t num_samples);
/* targets (num_targets) i32 ascending; *status |= 1 when the spectrum has fewer strict local maxima than num_targets. */
GCCNMF_API int gccnmf_pick_targets(gccnmf_handle* h, const double* mean_angular, int D, int num_targets, int32_t* targets,
                        int32_t* status, void* stream);
/* samples (2, n) f32; window (N) f64; E (F, D) c128; tdoas (D) f64; W (F, K), H (K, 2T) f32 in: seeded initial values
 * (gccNMFFunctions.py:70-73), out: learnt; signals (S, 2, gccnmf_istft_length(N, hop, T, 1)) f32; target_indexes (S) i32 and
 * status (1) i32 are device outputs: bit 0 too few peaks (the reference aborts), bit 1 all-NaN mask column (numpy raises),
 * bit 2 more near-tie argmax decisions than the float64 refinement list holds (re-run gccnmf_tdoa_gccnmf). */
GCCNMF_API int gccnmf_separate(gccnmf_handle* h, const gccnmf_pipeline_config* cfg, const float* samples, int64_t num_samples,
                    const double* window, const double* E, const double* tdoas, float* W, float* H, float* signals,
                    int32_t* target_indexes, int32_t* status, void* workspace, size_t workspace_bytes, void* stream);

/* ---- a13 + f-2: the real-time block path as one stream-ordered unit ----------------------------------
 * GCCNMFProcessor.processFrames (realtime/gccNMFProcessor.py:201-231 and the Theano graph of :245-270) with the
 * OverlapAddProcessor rings around it (realtime/utils.py:72-116) and, optionally, the per-frame coefficient inference of
 * notebooks/onlineSpeechEnhancement.ipynb:433-438.  Everything between the input block and the output block runs on
 * the device without a host synchronisation: five kernels per block (+ two per inference iteration), capturable in a
 * CUDA graph; the 8-block rings, the GCC-PHAT history, the sliding-window localisation and the target TDOA index are
 * device-resident state inside the caller-owned `state` buffer (sized by gccnmf_rt_state_bytes, filled by gccnmf_rt_init). */
typedef struct gccnmf_rt_config {
  int window_size;           /* N: power of two in [64, 2048] (config.py:63 windowSize)                              */
  int hop_size;              /* config.py:64                                                                          */
  int block_size;            /* samples per channel per audio block (config.py:65); the rings hold 8 blocks (utils.py:85) */
  int windows_per_block;     /* numTimePerChunk = blockSize // hopSize (config.py:112), at most 8                      */
  int num_atoms;             /* K                                                                                     */
  int num_tdoas;             /* D <= 128                                                                              */
  int history_length;        /* columns of the gccPHATHistory ring (runRealtimeGCCNMF.py:58)                          */
  int inference_iterations;  /* 0 = the reference's real-time class (numHUpdates is never used there); > 0: H-only KL updates */
  float sparsity_alpha;      /* of the inference updates (gccNMFFunctions.py:76)                                       */
  float epsilon;
} gccnmf_rt_config;
GCCNMF_API size_t gccnmf_rt_state_bytes(const gccnmf_rt_config* cfg);
/* W (F, K) f32; E (F, D) complex64 = expJOmegaTau (:248); windows (N) f32 (:186); H0 (K, 2) f32 seeded initial coefficients
 * (gccNMFFunctions.py:73 shape (K, 2)) or NULL when inference_iterations == 0.  Zeroes the rings and the history. */
GCCNMF_API int gccnmf_rt_init(gccnmf_handle* h, const gccnmf_rt_config* cfg, const float* W, const float* E,
                   const float* analysis_window, const float* synthesis_window, const float* H0, void* state,
                   size_t state_bytes, void* stream);
/* setTargetTDOARange (:272-276) and the attributes GCCNMFProcess sets (:136-151).  mode 0 boxcar / 1 window (:262-265).
 * set_target = 0 keeps the device-resident target TDOA index (it is loop-carried when localisation is enabled). */
GCCNMF_API int gccnmf_rt_set_params(gccnmf_handle* h, const gccnmf_rt_config* cfg, void* state, size_t state_bytes,
                         float target_index, int set_target, float epsilon, float beta, float noise_floor, int mode,
                         int separation_enabled, int localization_enabled, int localization_window, void* stream);
/* processFrames: windowed (2, N, nT) f32 -> out (2, N, nT) f32 (device buffers); updates history / localisation.
 * forced_atom_mask: NULL, or a (K, nT) f64 atom mask that replaces the one derived from the per-atom TDOA argmax
 * (externally decided masks; teacher-forced parity tests). */
GCCNMF_API int gccnmf_rt_process_frames(gccnmf_handle* h, const gccnmf_rt_config* cfg, void* state, size_t state_bytes,
                             const float* windowed, float* out, const double* forced_atom_mask, void* stream);
/* One audio block through rings + processFrames (utils.py:99-116): in_block (2, B) f32 -> out_block (2, B) f32. */
GCCNMF_API int gccnmf_rt_process_block(gccnmf_handle* h, const gccnmf_rt_config* cfg, void* state, size_t state_bytes,
                            const float* in_block, float* out_block, const double* forced_atom_mask, void* stream);
/* The same block as an instantiated CUDA graph ([H2D of in_host ->] kernels [-> D2H to out_host]); in_host / out_host are
 * pinned HOST buffers or NULL; `stream` must be a capturable (non-default) stream.  *graph_exec receives a cudaGraphExec_t. */
GCCNMF_API int gccnmf_rt_graph_create(gccnmf_handle* h, const gccnmf_rt_config* cfg, void* state, size_t state_bytes,
                           float* in_block, float* out_block, const float* in_host, float* out_host, void** graph_exec,
                           void* stream);
GCCNMF_API int gccnmf_rt_graph_launch(gccnmf_handle* h, void* graph_exec, void* stream);
GCCNMF_API int gccnmf_rt_graph_destroy(gccnmf_handle* h, void* graph_exec);
/* Stream-ordered copy of one item of the block state to dst (device or pinned host memory): 0 gccPHAT (D, nT) f32,
 * 1 target TDOA index f32, 2 atom mask (K, nT) f64, 3 / 4 input / output spectrogram (2, F, nT) c64, 5 per-atom TDOA
 * argmax (K, nT) i32, 6 inferred H (K, 2 nT) f32, 7 GCC-PHAT history ring (D, history_length) f64, 8 its write index i32. */
GCCNMF_API int gccnmf_rt_export(gccnmf_handle* h, const gccnmf_rt_config* cfg, void* state, size_t state_bytes, int what,
                     void* dst, void* stream);

/*
 * The building block the KL-NMF loop runs on (klnmf_tma.cu): the same 3-product contraction, TMA-fed, over operands
 * that are pre-split into bf16 hi/lo planes and kept in ONE orientation each; an operand contracted over its
 * non-contiguous dimension is consumed MN-major.  Test / diagnostics entry: splits the float32 operands into planes
 * in the workspace, then DT (N, M) row-major = (A . B^T)^T.
 *   a_mn_major = 0: A is (M, Kc) row-major;  1: A is (Kc, M) row-major.   b_mn_major likewise with N.
 *   tile_n in {128, 176, 208, 256};  splits > 1: `splits` partial slabs DT[z] over k ranges (N * M floats each).
 *   timing: device uint64[8 x CTAs] stamps (layout: gccnmf_debug_timing) or NULL.
 */
/* Tile plan of the KL-NMF loop on a device with sm_count SMs (host logic only; callable without a GPU):
 * out[0] tile width of the W.H contractions, out[1] of the H update, out[2] / out[3] tile width / k-splits of the W-update
 * numerator, out[4] n tiles of the H update, out[5..7] CTAs of the three launches.  < 0: shape not covered by this path. */
GCCNMF_API int gccnmf_klnmf_tile_plan(int sm_count, int F, int T2, int K, int* out);
GCCNMF_API size_t gccnmf_gemm_planes_workspace_bytes(int M, int N, int Kc);
/* Diagnostics: while `stamps` (device uint64) is non-NULL every plane GEMM launched through the handle appends 8 values
 * per CTA at a running offset: [0] / [7] %globaltimer ns at CTA start / end, [1..6] clock64 at start, first stage full,
 * last MMA issued, producer done, accumulator complete, epilogue end.  Returns the offset reached; reset != 0 rewinds. */
GCCNMF_API int64_t gccnmf_debug_timing(gccnmf_handle* h, unsigned long long* stamps, int reset);
GCCNMF_API int gccnmf_gemm_planes(gccnmf_handle* h, const float* A, int a_mn_major, const float* B, int b_mn_major, float* DT,
                       int M, int N, int Kc, int tile_n, int splits, void* workspace, size_t workspace_bytes,
                       unsigned long long* timing, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GCCNMF_B200_H_ */
