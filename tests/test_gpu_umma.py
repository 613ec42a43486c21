"""GPU: the tcgen05 3xTF32 GEMM building block against a float64 product (and against what plain
TF32 would give, to show the error compensation is doing its job)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def h():
    from gcc_nmf_b200._lib import default_handle
    return default_handle()


def _check(h, M, N, Kc, tile_n, positive, bf16=False):
    h.set_option('nmf_split_bf16', 1 if bf16 else 0)
    import torch
    g = torch.Generator(device='cpu').manual_seed(M * 7 + N * 3 + Kc)
    ld = (Kc + 3) // 4 * 4
    A = torch.zeros(M, ld)
    B = torch.zeros(N, ld)
    A[:, :Kc] = torch.rand(M, Kc, generator=g) if positive else torch.randn(M, Kc, generator=g)
    B[:, :Kc] = torch.rand(N, Kc, generator=g) if positive else torch.randn(N, Kc, generator=g)
    Ad, Bd = A.to(h.device), B.to(h.device)
    D = h.gemm_tn_3xtf32(Ad, Bd, Kc=Kc, tile_n=tile_n)
    torch.cuda.synchronize()
    ref = (Ad[:, :Kc].double() @ Bd[:, :Kc].double().T)
    scale = (Ad[:, :Kc].abs().double() @ Bd[:, :Kc].abs().double().T)     # |a|.|b| bound for the error
    err = ((D.double() - ref).abs() / scale).max().item()
    h.set_option('nmf_split_bf16', 1)      # library default
    return err


@pytest.mark.parametrize('tile_n', [128, 256])
@pytest.mark.parametrize('shape', [(128, 128, 32), (128, 256, 64), (256, 384, 1024), (100, 70, 40), (513, 3744, 1024),
                                   (1024, 3744, 513), (513, 1024, 3744)])
@pytest.mark.parametrize('bf16', [False, True])
def test_gemm_3xtf32_matches_float64(h, shape, tile_n, bf16):
    M, N, Kc = shape
    for positive in (True, False):
        err = _check(h, M, N, Kc, tile_n, positive, bf16)
        # The operand split recovers float32 products (plain TF32 would give ~5e-4); what remains is the tensor
        # core's float32 accumulator, which truncates on every accumulation: a bias of up to ~2^-24 per MMA
        # into the same accumulator (3 Kc / 8 of them), fully coherent for all-positive data.
        # 3xBF16 (hi/lo bf16): operand split error 2^-17 per product (sign-symmetric) and half as many accumulations.
        bound = (8e-6 + 4e-8 * (3 * Kc / 16)) if bf16 else (2e-6 + 4e-8 * (3 * Kc / 8))
        assert err < bound, (shape, tile_n, positive, bf16, err, bound)


def test_klnmf_tensor_core_path_matches_oracle(h):
    """KL-NMF with the contractions on tcgen05 (F, T2 >= 128, K % 4 == 0) against the CPU oracle."""
    import torch
    from oracle import gccnmf_oracle as orc
    F, T2, K = 257, 512, 64
    assert h.lib.gccnmf_klnmf_uses_tensor_cores(h.h, F, T2, K) == 1
    rng = np.random.default_rng(5)
    V = (rng.random((F, T2)) ** 3 + 1e-3).astype(np.float32)
    W0, H0 = orc.initKLNMF(F, T2, K)

    def rel(a, b):
        return float(np.linalg.norm(a - b) / np.linalg.norm(b))
    for split, name in ((0, '3xTF32'), (1, '3xBF16')):
        h.set_option('nmf_split_bf16', split)
        for iters, tol in ((1, 2e-5), (30, 1e-4)):
            W, H = h.to_device(W0.copy()), h.to_device(H0.copy())
            h.klnmf(h.to_device(V), W, H, iters)
            Wo, Ho = orc.performKLNMF(V, K, iters, 0, W0=W0, H0=H0)
            eW, eH = rel(W.cpu().numpy(), Wo), rel(H.cpu().numpy(), Ho)
            print('tensor-core KL-NMF %s %d iterations: rel W %.2e  rel H %.2e' % (name, iters, eW, eH))
            assert eW < tol and eH < tol, (name, iters, eW, eH)
    # step protocol (multi-GPU) against the fused loop
    V_d = h.to_device(V)
    W1, H1 = h.to_device(W0.copy()), h.to_device(H0.copy())
    W2, H2 = W1.clone(), H1.clone()
    h.klnmf(V_d, W1, H1, 2)
    numer = torch.empty(F * K + K, dtype=torch.float32, device=V_d.device)
    h.klnmf_begin(V_d, W2, H2)
    for it in range(2):
        h.klnmf_step_numer(V_d, W2, H2, it, numer)
        h.klnmf_step_apply(W2, H2, numer)
    h.klnmf_end(W2, H2, 2)
    assert rel(W2.cpu().numpy(), W1.cpu().numpy()) < 1e-6 and rel(H2.cpu().numpy(), H1.cpu().numpy()) < 1e-6
    # fixed dictionary (H-only inference)
    Hi = h.to_device(H0.copy())
    h.klnmf(V_d, h.to_device(Wo), Hi, 5, update_W=False)
    Href = H0.copy()
    denom = np.sum(Wo, axis=0)[:, None] + np.float32(1e-16)
    for _ in range(5):
        Href *= np.dot(Wo.T, V / np.dot(Wo, Href)) / denom
    assert rel(Hi.cpu().numpy(), Href) < 2e-5


@pytest.mark.parametrize('D', [32, 64])
def test_tdoa_argmax_tensor_core_equals_float64_kernel(h, D):
    """Every argmax decision of the tensor-core + refinement path equals the float64 kernel's (and the oracle's
    on a slice), including planted exact ties and a NaN bin."""
    import torch
    from oracle import gccnmf_oracle as orc
    import gcc_nmf_b200.gccNMFFunctions as fn
    rng = np.random.default_rng(D)
    F, T, K = 257, 300, 128
    coh = np.exp(1j * rng.uniform(-np.pi, np.pi, (F, T))).astype(np.complex64)
    coh[:, 5] = coh[:, 4]                       # two identical frames
    coh[7, 9] = np.nan                          # one NaN bin (0/0 in the PHAT normalisation)
    W = (rng.random((F, K)) ** 4).astype(np.float32)
    W[:, 3] = W[:, 2]                           # two identical atoms
    E = fn.getExpJOmegaTau(fn.getFrequenciesInHz(16000, F), fn.getTDOAsInSeconds(0.1, D))
    cd, Ed, Wd = h.to_device(coh), h.to_device(np.ascontiguousarray(E)), h.to_device(W)
    fast, refined = h.tdoa_argmax(cd, Ed, Wd)
    _, exact = h.tdoa_gccnmf(cd, Ed, Wd, want_values=False, want_argmax=True)
    assert int(refined.item()) <= h.lib.gccnmf_tdoa_argmax_refine_capacity(K, T)
    assert torch.equal(fast, exact)
    with np.errstate(all='ignore'):
        ref = orc.getGCCNMFAllTDOAs(coh[:, :16], E, W)
    assert np.array_equal(fast.cpu().numpy()[:, :16], np.argmax(ref, axis=1))
    print('D=%d: %d of %d decisions refined in float64' % (D, int(refined.item()), K * T))
