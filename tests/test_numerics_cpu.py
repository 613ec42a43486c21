"""CPU: the arithmetic claims behind the tensor-core KL-NMF path, checked in numpy (no GPU).

The plane GEMM (gcc-nmf_b200/csrc/tma_gemm.cuh) represents each float32 operand as hi = bf16(x), lo = bf16(x - hi) and
issues hi.hi + hi.lo + lo.hi per product.  DESIGN.md section 4.1 states: representation error <= 2^-17 |x|, product error
<= ~2^-16 |a b| (the dropped lo.lo term plus the two representation errors), sign-symmetric -- so a K-term contraction is
accurate to ~2^-16 / sqrt(K) relative to sum |a||b| for random data, well inside float32's own 2^-24 sqrt(K) accumulation."""
import numpy as np


def bf16_round(x):
    """Round-to-nearest-even float32 -> bfloat16, returned as float32 (what cvt.rn.bf16.f32 does)."""
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def split(x):
    hi = bf16_round(x)
    lo = bf16_round(np.asarray(x, np.float32) - hi)
    return hi, lo


def test_hi_lo_split_represents_16_mantissa_bits():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(200000) * np.exp(rng.uniform(-20, 20, 200000))).astype(np.float32)
    hi, lo = split(x)
    err = np.abs(x.astype(np.float64) - (hi.astype(np.float64) + lo.astype(np.float64)))
    assert (err <= 2.0 ** -17 * np.abs(x)).all()
    assert np.abs(lo).max() <= np.abs(x).max() * 2.0 ** -8          # lo is the rounding error of an 8-bit mantissa
    # exact for values that already fit 16 significant bits
    y = (rng.integers(-2 ** 15, 2 ** 15, 1000) * 2.0 ** rng.integers(-10, 10, 1000)).astype(np.float32)
    h2, l2 = split(y)
    assert np.array_equal(h2.astype(np.float64) + l2.astype(np.float64), y.astype(np.float64))


def test_three_product_contraction_error():
    rng = np.random.default_rng(1)
    for positive in (True, False):
        for K in (64, 513, 1024, 3744):
            A = (rng.random((48, K)) if positive else rng.standard_normal((48, K))).astype(np.float32)
            B = (rng.random((40, K)) if positive else rng.standard_normal((40, K))).astype(np.float32)
            ah, al = (t.astype(np.float64) for t in split(A))
            bh, bl = (t.astype(np.float64) for t in split(B))
            three = al @ bh.T + ah @ bl.T + ah @ bh.T           # exact accumulation: isolates the operand-split error
            exact = A.astype(np.float64) @ B.astype(np.float64).T
            scale = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T
            err = np.abs(three - exact) / scale
            assert err.max() < 2.0 ** -16, (positive, K, err.max())
            # random rounding errors average out over the contraction: far below the per-product bound
            assert err.max() < 4.0 * 2.0 ** -16 / np.sqrt(K) + 2.0 ** -24, (positive, K, err.max())
            # one bf16 product (no compensation) is three orders of magnitude worse -- why the split is needed at all
            single = np.abs(ah @ bh.T - exact) / scale
            assert single.max() > 30 * err.max()


def test_workspace_sizes_cover_the_buffers():
    import __graft_entry__ as entry
    entry.build()
    from gcc_nmf_b200 import _lib
    lib = _lib.load_library()
    F, T2, K = 513, 3744, 1024
    ws = lib.gccnmf_klnmf_workspace_bytes(F, T2, K)
    # H^T float32 + planes, V^T, R^T planes, W planes, 8 k-split slabs: DESIGN.md section 3
    need = T2 * K * 4 + 2 * T2 * K * 2 + T2 * 520 * 4 + 2 * T2 * 520 * 2 + 2 * F * K * 2 + 8 * F * K * 4
    assert need <= ws <= need * 1.05
    assert lib.gccnmf_klnmf_workspace_bytes(F, 2 * T2, K) > ws
    assert lib.gccnmf_gemm_planes_workspace_bytes(513, 3744, 1024) >= 2 * 2 * (513 + 3744) * 1024
    assert lib.gccnmf_klnmf_workspace_bytes(0, 10, 10) == 0
