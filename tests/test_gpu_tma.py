"""GPU: the TMA-fed plane GEMM (gcc-nmf_b200/csrc/tma_gemm.cuh) -- every operand layout (K-major / MN-major), tile width,
k-split and cluster shape against a float64 product -- and the KL-NMF path built on it (klnmf_tma.cu) against the
float32 SIMT path and the oracle, with the cluster shape forced both ways."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def h():
    from gcc_nmf_b200._lib import default_handle
    hd = default_handle()
    yield hd
    hd.set_option('gemm_cluster', -1)
    hd.set_option('gemm_pair', -1)
    hd.set_option('force_simt_nmf', 0)
    hd.set_option('nmf_pdl', 1)


def _gemm_error(h, M, N, Kc, a_mn, b_mn, tile_n, splits):
    import torch
    g = torch.Generator(device='cpu').manual_seed(M * 7 + N * 3 + Kc)
    A = torch.randn(M, Kc, generator=g).to(h.device)
    B = torch.rand(N, Kc, generator=g).to(h.device)
    Ain = A.T.contiguous() if a_mn else A
    Bin = B.T.contiguous() if b_mn else B
    DT = h.gemm_planes(Ain, Bin, a_mn, b_mn, tile_n=tile_n, splits=splits)
    torch.cuda.synchronize()
    D = DT.sum(0).T.double()
    ref = A.double() @ B.double().T
    scale = A.abs().double() @ B.abs().double().T      # |a|.|b| bound for the error
    return ((D - ref).abs() / scale).max().item()


# (M, N, Kc): one full tile; ragged everything (rows past the tiles, k tail, n tail); the three KL-NMF contraction shapes
# scaled down (M = 513 keeps the 513th-row SIMT tail, Kc = 513 the one-step k tail)
SHAPES = [(128, 128, 64), (200, 130, 70), (513, 640, 256), (256, 416, 513), (256, 513, 1024)]


@pytest.mark.parametrize('cluster', [11, 22, 12, 21])
@pytest.mark.parametrize('layout', [(False, False), (True, False), (True, True)])
def test_plane_gemm_matches_float64(h, layout, cluster):
    """3 x bf16 products per algorithmic product: error <= 2^-17 per product (sign-symmetric) plus the truncating
    float32 TMEM accumulator (~4e-8 per accumulation, coherent for all-positive data)."""
    a_mn, b_mn = layout
    h.set_option('gemm_cluster', cluster)      # 10 CN + CM; falls back to no cluster when it does not divide the tile grid
    try:
        for M, N, Kc in SHAPES:
            for tile_n in (128, 176, 208, 256):
                for splits in ((1, 3) if Kc >= 1024 else (1,)):
                    err = _gemm_error(h, M, N, Kc, a_mn, b_mn, tile_n, splits)
                    bound = 8e-6 + 4e-8 * (3 * Kc / 16)
                    assert err < bound, (layout, cluster, (M, N, Kc), tile_n, splits, err, bound)
    finally:
        h.set_option('gemm_cluster', -1)


def test_plane_gemm_rejects_bad_arguments(h):
    import torch
    from gcc_nmf_b200._lib import GCCNMFError, ParameterError
    A = torch.rand(128, 64, device=h.device)
    B = torch.rand(128, 64, device=h.device)
    with pytest.raises(GCCNMFError):
        h.gemm_planes(A, B, tile_n=100)                     # unsupported tile width
    with pytest.raises(ParameterError):
        h.gemm_planes(A, B, splits=99)                      # more k-splits than partial slabs
    with pytest.raises(GCCNMFError):
        h.gemm_planes(A, B.T.contiguous(), False, True)     # K-major A with MN-major B is not instantiated


def _rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.mark.parametrize('cluster', [-1, 11, 22])
def test_klnmf_tma_path_matches_simt_path_and_oracle(h, cluster):
    """F = 257 = 2 x 128 + 1 exercises the SIMT tail row; T2 = 512 and K = 256 give even tile grids, so the forced 2 x 2
    cluster shape really runs the multicast paths of all four contractions."""
    import torch
    from oracle import gccnmf_oracle as orc
    F, T2, K = 257, 512, 256
    assert h.lib.gccnmf_klnmf_uses_tensor_cores(h.h, F, T2, K) == 1
    rng = np.random.default_rng(9)
    V = (rng.random((F, T2)) ** 3 + 1e-3).astype(np.float32)
    W0, H0 = orc.initKLNMF(F, T2, K)
    Vd = h.to_device(V)
    h.set_option('gemm_cluster', cluster)
    try:
        out = {}
        for name, tma, pdl in (('simt', 0, 1), ('tma', 1, 1), ('tma_nopdl', 1, 0)):
            h.set_option('force_simt_nmf', 0 if tma else 1)
            h.set_option('nmf_pdl', pdl)
            for iters in (1, 25):
                W, H = h.to_device(W0.copy()), h.to_device(H0.copy())
                h.klnmf(Vd, W, H, iters)
                torch.cuda.synchronize()
                out[name, iters] = (W.cpu().numpy(), H.cpu().numpy())
        Wo, Ho = orc.performKLNMF(V, K, 25, 0, W0=W0, H0=H0)
        for name in ('simt', 'tma', 'tma_nopdl'):
            eW, eH = _rel(out[name, 25][0], Wo), _rel(out[name, 25][1], Ho)
            print('KL-NMF %s cluster %d, 25 iterations: rel W %.2e rel H %.2e' % (name, cluster, eW, eH))
            assert eW < 1e-4 and eH < 1e-4, (name, cluster, eW, eH)
        # one iteration: the hi/lo bf16 products against plain float32 FMAs -- only rounding-level differences remain
        assert _rel(out['tma', 1][0], out['simt', 1][0]) < 1e-5 and _rel(out['tma', 1][1], out['simt', 1][1]) < 1e-5
        # programmatic dependent launch changes the schedule, not the arithmetic
        assert np.array_equal(out['tma', 25][0], out['tma_nopdl', 25][0]) and np.array_equal(out['tma', 25][1], out['tma_nopdl', 25][1])
        # sparsity (alpha > 0) goes through the per-row reciprocal of the H update
        h.set_option('force_simt_nmf', 0)
        h.set_option('nmf_pdl', 1)
        W, H = h.to_device(W0.copy()), h.to_device(H0.copy())
        h.klnmf(Vd, W, H, 5, sparsity_alpha=0.3)
        Wa, Ha = orc.performKLNMF(V, K, 5, 0.3, W0=W0, H0=H0)
        assert _rel(W.cpu().numpy(), Wa) < 2e-5 and _rel(H.cpu().numpy(), Ha) < 2e-5
    finally:
        h.set_option('gemm_cluster', -1)
        h.set_option('force_simt_nmf', 0)
        h.set_option('nmf_pdl', 1)


def test_debug_timing_records_every_cta(h):
    """gccnmf_debug_timing: 8 stamps per CTA of every plane GEMM launched while it is armed, plus one record per W update."""
    import torch
    from oracle import gccnmf_oracle as orc
    F, T2, K = 257, 640, 128
    rng = np.random.default_rng(2)
    V = h.to_device((rng.random((F, T2)) + 1e-3).astype(np.float32))
    W0, H0 = orc.initKLNMF(F, T2, K)
    W, H = h.to_device(W0), h.to_device(H0)
    buf = torch.zeros(1 << 18, dtype=torch.int64, device=h.device)
    h.lib.gccnmf_debug_timing(h.h, buf.data_ptr(), 1)
    try:
        h.klnmf(V, W, H, 1)
        torch.cuda.synchronize()
    finally:
        used = h.lib.gccnmf_debug_timing(h.h, None, 1)
    assert used > 0 and used % 8 == 0
    s = buf.cpu().numpy()[:used].reshape(-1, 8)
    assert (s[:, 0] > 0).all() and (s[:, 7] >= s[:, 0]).all()          # %globaltimer at CTA start / end
    gemm = s[:, 6] > 0
    assert gemm.sum() == len(s) - 1                                     # the W update of the one iteration leaves one record
    assert (s[gemm, 6] > s[gemm, 1]).all()                              # clock64: epilogue end after kernel entry
    assert h.lib.gccnmf_debug_timing(h.h, None, 1) == 0                 # disarmed and rewound


@pytest.mark.parametrize('layout', [(False, False), (True, False), (True, True)])
def test_plane_gemm_cta_pairs(h, layout):
    """set_option('gemm_pair', 1): two m tiles issue one 256-row tcgen05.mma.cta_group::2.  With K-major B tiles of <= 128 columns the
    pair also runs the dual-N loop (N = 2 BN = [leader: B_hi ; peer: B_lo], two MMAs per k-step); wider tiles split both planes of B
    in halves (three MMAs per k-step)."""
    a_mn, b_mn = layout
    h.set_option('gemm_pair', 1)
    try:
        for M, N, Kc in [(256, 256, 64), (512, 640, 256), (256, 416, 513), (1024, 513, 1024)]:
            for tile_n in ((128, 256) if b_mn else (128, 176, 208, 256)):
                err = _gemm_error(h, M, N, Kc, a_mn, b_mn, tile_n, 1)
                assert err < 8e-6 + 4e-8 * (3 * Kc / 16), (layout, (M, N, Kc), tile_n, err)
    finally:
        h.set_option('gemm_pair', -1)


def test_klnmf_pairs_on_and_off_agree(h):
    """The W.H contractions run on CTA pairs + dual-N by default (gemm_pair = -1); gemm_pair = 0 runs the same four products per
    k-step on single CTAs: identical arithmetic, so identical results."""
    import torch
    from oracle import gccnmf_oracle as orc
    F, T2, K = 513, 640, 128
    rng = np.random.default_rng(12)
    V = h.to_device((rng.random((F, T2)) ** 3 + 1e-3).astype(np.float32))
    W0, H0 = orc.initKLNMF(F, T2, K)
    out = {}
    try:
        for pair in (-1, 0):
            h.set_option('gemm_pair', pair)
            W, H = h.to_device(W0.copy()), h.to_device(H0.copy())
            h.klnmf(V, W, H, 10)
            torch.cuda.synchronize()
            out[pair] = (W.cpu().numpy(), H.cpu().numpy())
    finally:
        h.set_option('gemm_pair', -1)
    assert np.array_equal(out[-1][0], out[0][0]) and np.array_equal(out[-1][1], out[0][1])


def test_plane_gemm_dual_n_narrow_tiles(h):
    """104- and 112-column tiles of the dual-N loop (2 x 104 = 208 / 224 columns per MMA): K-major B only; the 40-column
    half of a 104-column tile exercises the 8-column tcgen05.ld remainder of the epilogue."""
    for tile_n in (104, 112):
        for (M, N, Kc) in [(128, 104, 64), (513, 640, 256), (256, 416, 513)]:
            for a_mn in (False, True):
                err = _gemm_error(h, M, N, Kc, a_mn, False, tile_n, 1)
                assert err < 8e-6 + 4e-8 * (3 * Kc / 16), (tile_n, (M, N, Kc), a_mn, err)
