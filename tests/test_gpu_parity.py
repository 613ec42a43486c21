"""GPU parity tests: every C-ABI stage against the golden fixtures produced by the unmodified
reference (tests/golden, oracle/make_golden.py) and against the CPU oracle on seeded inputs.

Stage tests are TEACHER-FORCED (each stage gets the reference's inputs for that stage) because the
binary masks make end-to-end tolerances meaningless (hard part 2); integer
decisions (TDOA indexes, argmax, masks) must be bit-exact, floating-point stages must agree within
the stated tolerance (north star: 1e-4 relative float32; most stages are held much tighter).
Everything goes through the C ABI (ctypes) -- never through the oracle.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import gccnmf_oracle as orc  # noqa: E402  (the checker)


def relerr(a, b):
    dt = np.complex128 if (np.iscomplexobj(a) or np.iscomplexobj(b)) else np.float64
    a, b = np.asarray(a, dt), np.asarray(b, dt)
    return np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300)


@pytest.fixture(scope='module')
def fn():
    import gcc_nmf_b200.gccNMFFunctions as fn
    return fn


@pytest.fixture(scope='module')
def h():
    from gcc_nmf_b200._lib import default_handle
    return default_handle()


# ------------------------------------------------------------------------------------ a1 STFT
def test_stft_matches_reference(golden, fn):
    g = golden('separation_mini')
    sr, N, hop = [int(v) for v in g['params'][:3]]
    X = fn.computeComplexMixtureSpectrogram(g['samples'], N, hop, np.hanning)
    assert X.shape == g['X'].shape and X.dtype == np.complex64
    # float64 FFT rounded once to complex64: identical up to rare last-bit differences
    scale = np.abs(g['X']).max()
    assert np.abs(X - g['X']).max() <= 2e-7 * scale
    assert np.mean(X == g['X']) > 0.98


def test_stft_librosa_dropin_and_errors(golden):
    from gcc_nmf_b200.librosaSTFT import stft, istft, ParameterError
    g = golden('separation_mini')
    sr, N, hop = [int(v) for v in g['params'][:3]]
    X0 = stft(g['samples'][0].copy(), N, hop, N, np.hanning, center=False)
    assert np.abs(X0 - g['X'][0]).max() <= 2e-7 * np.abs(g['X']).max()
    with pytest.raises(ParameterError):
        stft(np.zeros(100, np.float32), 256, 64, center=False)          # buffer too short (librosaSTFT.py:427)
    with pytest.raises(ParameterError):
        stft(np.zeros(1000, np.float32), 256, 0, center=False)          # invalid hop (:416)
    with pytest.raises(ParameterError):
        stft(np.full(1000, np.nan, np.float32), 256, 64, center=False)  # not finite (:486)
    x = np.random.default_rng(0).standard_normal(256).astype(np.float32)
    X1 = stft(x, 256, 64, None, np.hanning, center=False)               # exactly one frame
    assert X1.shape == (129, 1)
    assert relerr(X1, orc.stft(x, 256, 64)) < 1e-6
    assert istft(X1, 64, 256, np.hanning).shape == (0,)                 # centre trim eats the only frame


@pytest.mark.parametrize('n_fft,hop', [(32, 8), (64, 64), (512, 128), (2048, 512), (4096, 1024)])
def test_stft_sizes_against_oracle(fn, n_fft, hop):
    rng = np.random.default_rng(n_fft)
    x = (0.1 * rng.standard_normal((2, 3 * n_fft + 17))).astype(np.float32)
    X = fn.computeComplexMixtureSpectrogram(x, n_fft, hop, np.hanning)
    Xo = orc.computeComplexMixtureSpectrogram(x, n_fft, hop)
    assert X.shape == Xo.shape
    assert np.abs(X - Xo).max() <= 3e-7 * np.abs(Xo).max()


# ------------------------------------------------------------------------------------ a2 KL-NMF
def test_klnmf_single_iteration_teacher_forced(golden, h):
    g = golden('separation_mini')
    W, H = h.to_device(g['W0'].copy()), h.to_device(g['H0'].copy())
    h.klnmf(h.to_device(g['V']), W, H, 1, 0.0, 1e-16)
    assert relerr(W.cpu().numpy(), g['W1']) < 2e-6
    assert relerr(H.cpu().numpy(), g['H1']) < 2e-6
    np.testing.assert_allclose(W.cpu().numpy(), g['W1'], rtol=2e-5, atol=1e-9)
    np.testing.assert_allclose(H.cpu().numpy(), g['H1'], rtol=2e-5, atol=1e-9)


def test_klnmf_full_run_and_sparsity(golden, fn):
    g = golden('separation_mini')
    K, I = int(g['params'][5]), int(g['params'][6])
    W, H = fn.performKLNMF(g['V'], K, I, 0)
    assert W.dtype == np.float32 and H.dtype == np.float32
    assert relerr(W, g['W']) < 1e-4 and relerr(H, g['H']) < 1e-4      # north-star tolerance after 25 iterations
    W3, H3 = fn.performKLNMF(g['V'], K, 3, 0.5)
    assert relerr(W3, g['W3_alpha']) < 1e-5 and relerr(H3, g['H3_alpha']) < 1e-5


def test_klnmf_building_blocks_equal_fused(golden, h):
    g = golden('separation_mini')
    import torch
    V = h.to_device(g['V'])
    W1, H1 = h.to_device(g['W0'].copy()), h.to_device(g['H0'].copy())
    W2, H2 = W1.clone(), H1.clone()
    h.klnmf(V, W1, H1, 2, 0.0, 1e-16)
    F, K = W2.shape
    numer = torch.empty(F * K + K, dtype=torch.float32, device=V.device)
    h.klnmf_begin(V, W2, H2)
    for it in range(2):
        h.klnmf_step_numer(V, W2, H2, it, numer)
        h.klnmf_step_apply(W2, H2, numer)
    h.klnmf_end(W2, H2, 2)
    assert torch.equal(W1, W2) and torch.equal(H1, H2)


def test_klnmf_float64_pretraining_input(golden, fn):
    g = golden('pretraining_mini')
    W, H = fn.performKLNMF(g['trainV'], 12, 20, 0, 1e-16, 0)
    # the reference keeps float64 intermediates for float64 V; this path is float32 throughout
    assert relerr(W, g['W']) < 1e-4 and relerr(H, g['H']) < 1e-4


def test_infer_coefficients_matches_oracle(golden, fn):
    g = golden('separation_mini')
    Hgpu = fn.inferCoefficientsKLNMF(g['V'][:, :40], g['W'], 7, 0, 1e-16, 0)
    Horc = orc.inferCoefficientsKLNMF(g['V'][:, :40], g['W'], 7, 0, 1e-16, 0)
    assert relerr(Hgpu, Horc) < 1e-5


# ------------------------------------------------------------------------------------ a3, a4, a5
def test_coherence_and_angular_spectrogram(golden, fn):
    g = golden('separation_mini')
    sr, N, hop, D = [int(v) for v in g['params'][:4]]
    coh = fn.getSpectralCoherence(g['X'])
    assert coh.dtype == np.complex64
    assert np.abs(coh - g['coherence']).max() < 5e-7        # numpy complex64 op order, few-ulp agreement
    f = np.linspace(0, sr / 2.0, N // 2 + 1)
    A = fn.getAngularSpectrogram(g['coherence'], f, float(g['micSep']), D)   # teacher-forced coherence
    assert A.dtype == np.float64 and A.shape == g['angularSpectrogram'].shape
    np.testing.assert_allclose(A, g['angularSpectrogram'], rtol=0, atol=1e-10)
    idx = fn.estimateTargetTDOAIndexesFromAngularSpectrum(A.mean(axis=-1), float(g['micSep']), D, int(g['params'][4]))
    assert list(idx) == list(g['targetTDOAIndexes'])


def test_fused_phat_angspec_mean(golden, h, fn):
    g = golden('separation_mini')
    sr, N, hop, D = [int(v) for v in g['params'][:4]]
    f = np.linspace(0, sr / 2.0, N // 2 + 1)
    E = fn.getExpJOmegaTau(f, fn.getTDOAsInSeconds(float(g['micSep']), D))
    coh, ang, mean = h.phat_angspec(h.to_device(g['X']), h.to_device(np.ascontiguousarray(E)))
    np.testing.assert_allclose(ang.cpu().numpy(), g['angularSpectrogram'], rtol=0, atol=2e-4)  # coherence is float32-rounded upstream
    np.testing.assert_allclose(mean.cpu().numpy(), ang.cpu().numpy().mean(axis=-1), rtol=1e-12, atol=1e-12)


def test_nan_bins_propagate_like_numpy(h, fn):
    # a zero bin gives 0/0 = NaN coherence (runGCCNMF.py:44 is unguarded); argmax treats NaN as the maximum
    rng = np.random.default_rng(3)
    X = (rng.standard_normal((2, 33, 8)) + 1j * rng.standard_normal((2, 33, 8))).astype(np.complex64)
    X[0, 5, 2] = 0
    coh = fn.getSpectralCoherence(X)
    with np.errstate(all='ignore'):
        ref = X[0] * X[1].conj() / np.abs(X[0]) / np.abs(X[1])
    assert np.isnan(coh[5, 2]) and np.isnan(ref[5, 2])
    assert np.array_equal(np.isnan(coh), np.isnan(ref))


# ------------------------------------------------------------------------------------ a6, a7, a8, a9
def test_target_gccnmf_masks_recon_istft_teacher_forced(golden, fn):
    g = golden('separation_mini')
    sr, N, hop, D, S, K, I = [int(v) for v in g['params']]
    f = np.linspace(0, sr / 2.0, N // 2 + 1)
    stereoH = np.array(np.hsplit(g['H'], 2))
    G = fn.getTargetTDOAGCCNMFs(g['coherence'], float(g['micSep']), D, f, list(g['targetTDOAIndexes']), g['W'], stereoH)
    assert G.shape == g['targetTDOAGCCNMFs'].shape and G.dtype == np.float32
    np.testing.assert_allclose(G, g['targetTDOAGCCNMFs'], rtol=1e-5, atol=1e-6)
    M = fn.getTargetCoefficientMasks(g['targetTDOAGCCNMFs'], S)
    assert np.array_equal(M, g['targetCoefficientMasks'])            # bit-exact decisions
    assert np.array_equal(fn.getTargetCoefficientMasks(G, S), g['targetCoefficientMasks'])
    Sp = fn.getTargetSpectrogramEstimates(g['targetCoefficientMasks'], g['X'], g['W'], stereoH)
    assert Sp.shape == g['targetSpectrogramEstimates'].shape and Sp.dtype == np.complex64
    assert relerr(Sp, g['targetSpectrogramEstimates']) < 2e-6
    y = fn.getTargetSignalEstimates(g['targetSpectrogramEstimates'], N, hop, np.hanning)
    assert y.shape == g['targetSignalEstimates'].shape and y.dtype == np.float32
    assert np.abs(y - g['targetSignalEstimates']).max() < 2e-6 * np.abs(g['targetSignalEstimates']).max() + 1e-9


def test_coeff_mask_nan_semantics(fn):
    G = np.random.default_rng(1).standard_normal((3, 4, 5)).astype(np.float32)
    G[0, 1, 1] = np.nan                 # ignored by nanargmax
    G[:, 2, 2] = G[1, 2, 2]             # tie -> first index
    M = fn.getTargetCoefficientMasks(G, 3)
    assert np.array_equal(M, orc.getTargetCoefficientMasks(G, 3))
    G[:, 3, 3] = np.nan                 # all-NaN slice raises like numpy.nanargmax
    with pytest.raises(ValueError):
        fn.getTargetCoefficientMasks(G, 3)


# ------------------------------------------------------------------------------------ a10 enhancement mask
def test_enhancement_argmax_bit_exact(golden, fn):
    g = golden('enhancement_mini')
    sr, N, hop, D, S, K, I = [int(v) for v in g['params']]
    f = fn.getFrequenciesInHz(sr, N // 2 + 1)
    argmax = fn.getGCCNMFArgMaxTDOA(g['coherence'], f, float(g['micSep']), D, g['W'])
    assert np.array_equal(argmax, g['argMaxGCCNMF'])
    tdoas = fn.getTDOAsInSeconds(float(g['micSep']), D)
    lut = fn.getTargetTDOALookup(tdoas, int(g['targetTDOAIndexes'][0]), (tdoas[-1] - tdoas[0]) * 0.05)
    assert np.array_equal(lut[argmax][None], g['targetCoefficientMasks'])


@pytest.mark.parametrize('D', [4, 8, 32, 64, 128])
def test_argmax_over_tdoa_all_supported_sizes(h, fn, D):
    rng = np.random.default_rng(D)
    F, T, K = 65, 21, 70
    coh = np.exp(1j * rng.uniform(-np.pi, np.pi, (F, T))).astype(np.complex64)
    W = rng.random((F, K)).astype(np.float32)
    f = fn.getFrequenciesInHz(16000, F)
    E = fn.getExpJOmegaTau(f, fn.getTDOAsInSeconds(0.2, D))
    values, argmax = h.tdoa_gccnmf(h.to_device(coh), h.to_device(np.ascontiguousarray(E)), h.to_device(W),
                                   want_values=True, want_argmax=True)
    ref = orc.getGCCNMFAllTDOAs(coh, E, W)                          # (K, D, T) float64
    assert np.array_equal(argmax.cpu().numpy(), np.argmax(ref, axis=1))
    np.testing.assert_allclose(values.cpu().numpy(), ref.transpose(1, 0, 2).astype(np.float32), rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------------------------ end to end
def test_pipeline_separation_end_to_end(golden):
    from gcc_nmf_b200.pipeline import GCCNMFPipeline
    g = golden('separation_mini')
    sr, N, hop, D, S, K, I = [int(v) for v in g['params']]
    pipe = GCCNMFPipeline(sr, N, hop, D, float(g['micSep']), K, I)
    r = pipe.separate(pipe.h.to_device(g['samples']), S)
    assert r['targetTDOAIndexes'] == [int(i) for i in g['targetTDOAIndexes']]
    masks = r['targetCoefficientMasks'].cpu().numpy()
    agreement = np.mean(masks == g['targetCoefficientMasks'])
    assert agreement > 0.995, agreement          # free-running float32 NMF may flip a handful of near-ties
    y = r['targetSignalEstimates'].cpu().numpy()
    assert y.shape == g['targetSignalEstimates'].shape
    if agreement == 1.0:
        assert relerr(y, g['targetSignalEstimates']) < 1e-4


def test_pipeline_enhancement_end_to_end(golden):
    from gcc_nmf_b200.pipeline import GCCNMFPipeline
    g = golden('enhancement_mini')
    sr, N, hop, D, S, K, I = [int(v) for v in g['params']]
    pipe = GCCNMFPipeline(sr, N, hop, D, float(g['micSep']), K, I)
    out = pipe.enhance_host(g['samples']).numpy()
    r = pipe.enhance(pipe.h.to_device(g['samples']))
    assert r['targetTDOAIndexes'] == [int(i) for i in g['targetTDOAIndexes']]
    assert relerr(r['W'].cpu().numpy(), g['W']) < 1e-4 and relerr(r['H'].cpu().numpy(), g['H']) < 1e-4
    agreement = np.mean(r['targetCoefficientMasks'].cpu().numpy() == g['targetCoefficientMasks'])
    assert agreement > 0.995, agreement
    assert out.shape == g['targetSignalEstimates'].shape
    assert np.array_equal(out, r['targetSignalEstimates'].cpu().numpy())   # deterministic run to run
    if agreement == 1.0:
        assert relerr(out, g['targetSignalEstimates']) < 1e-4


# ------------------------------------------------------------------------------------ full-size properties (C2 shape)
def test_full_size_properties_config2(h, fn):
    """BASELINE.json configs[1] shape: 30 s @ 16 kHz, N=1024, hop=256, K=1024, D=64."""
    import torch
    from gcc_nmf_b200.synth import synthetic_stereo
    N, hop, K, D = 1024, 256, 1024, 64
    x = synthetic_stereo(30.0)
    xd = h.to_device(x)
    window = h.to_device(np.hanning(N))
    X, V = h.stft(xd, window, N, hop, conjugate=True, want_V=True)
    assert tuple(X.shape) == (2, 513, 1872) and tuple(V.shape) == (513, 3744)
    # STFT: Parseval-type check per frame against the windowed samples (size-independent property)
    t = 777
    frame = np.hanning(N) * x[0, t * hop:t * hop + N].astype(np.float64)
    Xt = X[0, :, t].cpu().numpy().astype(np.complex128)
    energy = (np.abs(Xt[0]) ** 2 + np.abs(Xt[-1]) ** 2 + 2 * np.sum(np.abs(Xt[1:-1]) ** 2)) / N
    assert abs(energy - np.sum(frame ** 2)) < 1e-5 * np.sum(frame ** 2)
    # STFT -> iSTFT: y = gain * x * sum_i w^2(n - i hop) on the centre-trimmed support
    y = h.istft_ola(X, window, N, hop, gain=np.float32(0.5), center=True, conjugate=True).cpu().numpy()
    T = X.shape[2]
    w2 = np.zeros(N + hop * (T - 1))
    for i in range(T):
        w2[i * hop:i * hop + N] += np.hanning(N) ** 2
    expect = 0.5 * (x[:, :len(w2)].astype(np.float64) * w2)[:, N // 2:-(N // 2)]
    assert y.shape == expect.shape
    assert np.abs(y - expect).max() < 5e-6
    # KL-NMF: 3 iterations against the oracle at full size, unit-L2 atoms, KL objective decreasing
    W0, H0 = fn._seededInit(513, 3744, K, 1e-16, 0)
    Vh = V.cpu().numpy()
    W, H = h.to_device(W0), h.to_device(H0)

    def kl(Wt, Ht):
        P = Wt @ Ht
        return float(torch.sum(V * torch.log(V / P) - V + P))
    kls = [kl(W, H)]
    for _ in range(3):
        h.klnmf(V, W, H, 1)
        kls.append(kl(W, H))
    assert all(b < a for a, b in zip(kls, kls[1:])), kls
    Wo, Ho = orc.performKLNMF(Vh, K, 3, 0, W0=W0, H0=H0)
    assert relerr(W.cpu().numpy(), Wo) < 1e-5 and relerr(H.cpu().numpy(), Ho) < 1e-5
    np.testing.assert_allclose(torch.linalg.norm(W, dim=0).cpu().numpy(), 1.0, atol=1e-5)
    # all-TDOA argmax: exact agreement with the oracle on a slice of frames
    f = fn.getFrequenciesInHz(16000, 513)
    E = fn.getExpJOmegaTau(f, fn.getTDOAsInSeconds(0.1, D))
    coh, ang, mean = h.phat_angspec(X, h.to_device(np.ascontiguousarray(E)))
    Ed = h.to_device(np.ascontiguousarray(E))
    argmax, refined = h.tdoa_argmax(coh, Ed, W)              # tensor cores + float64 refinement of near-ties
    _, argmax64 = h.tdoa_gccnmf(coh, Ed, W)                  # float64 kernel: all 1.9 M decisions must agree
    assert torch.equal(argmax, argmax64)
    print('config 2: %d of %d argmax decisions refined in float64' % (int(refined.item()), argmax.numel()))
    cohh = coh.cpu().numpy()
    ref = orc.getGCCNMFAllTDOAs(cohh[:, 500:532], E, W.cpu().numpy())
    assert np.array_equal(argmax.cpu().numpy()[:, 500:532], np.argmax(ref, axis=1))
    assert fn.estimateTargetTDOAIndexesFromAngularSpectrum(mean.cpu().numpy(), 0.1, D, 1) == [12]


@pytest.mark.parametrize('S,F,T,K', [(1, 513, 1872, 1024), (3, 513, 311, 128), (2, 257, 100, 64)])
def test_masked_reconstruction_tensor_cores_vs_float32_kernel(h, S, F, T, K):
    """a8 on the plane GEMM (masked-H planes, 3 bf16 products per product) against the float32 SIMT kernel and float64 numpy:
    T = 1872 (whole vector segments), 311 (T % 4 != 0: scalar epilogue path, ragged last tile) and 100 (a single partial tile)."""
    import torch
    rng = np.random.default_rng(S * 1000 + T)
    W = (rng.random((F, K)) ** 2).astype(np.float32)
    H = (rng.random((K, 2 * T)) ** 3).astype(np.float32)
    masks = (rng.random((S, K, T)) < 0.3).astype(np.float32)
    X = (rng.standard_normal((2, F, T)) + 1j * rng.standard_normal((2, F, T))).astype(np.complex64)
    X[0, 3, 5] = 0                                                   # angle(0) = 0
    Wd, Hd, Md, Xd = h.to_device(W), h.to_device(H), h.to_device(masks), h.to_device(X)
    assert h.lib.gccnmf_masked_recon_workspace_bytes(S, F, T, K) > 256                  # the tensor-core path takes these shapes
    tc = h.masked_recon_phase(Md, Xd, Wd, Hd, tensor_cores=True).cpu().numpy()
    simt = h.masked_recon_phase(Md, Xd, Wd, Hd, tensor_cores=False).cpu().numpy()
    torch.cuda.synchronize()
    stereoH = np.array(np.hsplit(H, 2))
    ref = orc.getTargetSpectrogramEstimates(masks, X, W, stereoH)
    assert tc.shape == ref.shape == (S, 2, F, T)
    assert relerr(simt, ref) < 2e-6
    assert relerr(tc, ref) < 1e-5 and np.abs(tc - ref).max() < 3e-5 * np.abs(ref).max()


def test_fused_separate_call_equals_staged_pipeline(golden):
    """gccnmf_separate (one C-ABI call, target TDOAs picked by the device kernel with argrelmax semantics) against the staged
    pipeline (scipy peak picking on the host, like the reference): same targets, bit-identical W, H and signals -- both flows."""
    import torch
    from gcc_nmf_b200.pipeline import GCCNMFPipeline
    from gcc_nmf_b200.synth import synthetic_stereo
    g = golden('separation_mini')
    sr, N, hop, D, S, K, I = [int(v) for v in g['params']]
    pipe = GCCNMFPipeline(sr, N, hop, D, float(g['micSep']), K, I)
    x = pipe.h.to_device(g['samples'])
    staged = pipe.separate(x, S)
    fused = pipe.run_fused(x, S)
    torch.cuda.synchronize()
    pipe.raise_on_status(fused['status'])
    assert fused['targetTDOAIndexes'].cpu().tolist() == staged['targetTDOAIndexes'] == [int(i) for i in g['targetTDOAIndexes']]
    assert torch.equal(fused['W'], staged['W']) and torch.equal(fused['H'], staged['H'])
    assert torch.equal(fused['targetSignalEstimates'], staged['targetSignalEstimates'])
    # enhancement flow at a shape that takes every tensor-core path (KL-NMF planes, argmax GEMM, masked reconstruction)
    pipe2 = GCCNMFPipeline(16000, 1024, 256, 64, 0.1, 256, 10)
    x2 = pipe2.h.to_device(synthetic_stereo(6.0, seed=3))
    staged2 = pipe2.enhance(x2)
    y_staged = staged2['targetSignalEstimates'].clone()
    fused2 = pipe2.run_fused(x2, 0)
    torch.cuda.synchronize()
    pipe2.raise_on_status(fused2['status'])
    assert fused2['targetTDOAIndexes'].cpu().tolist() == staged2['targetTDOAIndexes']
    assert torch.equal(fused2['W'], staged2['W']) and torch.equal(fused2['targetSignalEstimates'], y_staged)


def test_device_peak_picking_matches_scipy(h, fn):
    """gccnmf_pick_targets against scipy.signal.argrelmax + top-S (gccNMFFunctions.py:94-116) on random spectra, plateaus and
    too-few-peaks cases."""
    import torch
    rng = np.random.default_rng(8)
    for trial in range(40):
        D = int(rng.choice([8, 16, 64, 128]))
        x = rng.standard_normal(D)
        if trial % 5 == 0:
            x[3:6] = x[3]                                  # a plateau is not a strict maximum
        S = int(rng.integers(1, 4))
        peaks = fn.argrelmax(x)[0]
        xd = h.to_device(x)
        targets = torch.zeros(S, dtype=torch.int32, device=h.device)
        status = torch.zeros(1, dtype=torch.int32, device=h.device)
        h.check(h.lib.gccnmf_pick_targets(h.h, xd.data_ptr(), D, S, targets.data_ptr(), status.data_ptr(), h.stream))
        if len(peaks) < S:
            assert int(status.item()) & 1
            with pytest.raises(ValueError):
                fn.estimateTargetTDOAIndexesFromAngularSpectrum(x, 0.1, D, S)
        else:
            assert int(status.item()) == 0
            assert targets.cpu().tolist() == [int(i) for i in fn.estimateTargetTDOAIndexesFromAngularSpectrum(x, 0.1, D, S)]
