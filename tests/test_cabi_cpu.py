"""CPU: the C-ABI library loads without a GPU, exports every symbol include/gccnmf_b200.h declares,
host-only helpers behave, and the product path refuses to run without a device (no CPU fallback)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as entry
    entry.build()
    from gcc_nmf_b200 import _lib
    return _lib.load_library()


def header_symbols():
    text = open(os.path.join(ROOT, 'include', 'gccnmf_b200.h')).read()
    return sorted(set(re.findall(r'GCCNMF_API [\w\s\*]+?\b(gccnmf_\w+)\(', text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from gcc_nmf_b200 import _lib
    names = header_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), 'missing export: ' + name
    assert sorted(_lib.SIGNATURES) == names


def test_host_helpers(lib):
    assert lib.gccnmf_abi_version() == 2
    assert lib.gccnmf_stft_num_frames(160000, 1024, 512) == 311          # config 1
    assert lib.gccnmf_stft_num_frames(480000, 1024, 256) == 1872         # config 2
    assert lib.gccnmf_stft_num_frames(100, 1024, 256) < 0                # buffer too short
    assert lib.gccnmf_stft_num_frames(4096, 1024, 0) < 0                 # invalid hop
    assert lib.gccnmf_istft_length(1024, 256, 1872, 1) == 478976
    assert lib.gccnmf_istft_length(1024, 512, 311, 1) == 158720
    assert lib.gccnmf_klnmf_workspace_bytes(513, 3744, 1024) > 513 * 3744 * 4
    assert lib.gccnmf_status_string(-5).decode().startswith('no CUDA device')


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from gcc_nmf_b200 import _lib
    with pytest.raises(_lib.GCCNMFError):
        _lib.Handle(0)
    import gcc_nmf_b200.gccNMFFunctions as fn
    import numpy as np
    with pytest.raises(_lib.GCCNMFError):
        fn.performKLNMF(np.ones((8, 8), np.float32), 2, 1, 0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'gcc-nmf_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, re.M), f


def test_klnmf_tile_plan_fills_the_chip(lib):
    """Host-side tile planner of the TMA KL-NMF path (klnmf_tma.cu make_plan): one wave on 148 SMs at the headline shape,
    and sane choices at the other BASELINE.json shapes."""
    import ctypes
    out = (ctypes.c_int * 8)()
    assert lib.gccnmf_klnmf_tile_plan(148, 513, 3744, 1024, out) == 0            # config 2
    bn_wh, bn_h, bn_w, splits_w, slots, c_wh, c_h, c_w = list(out)
    assert (bn_h, bn_w, splits_w) == (208, 176, 6)
    assert bn_wh in (104, 112, 128) and c_wh == 4 * ((3744 + bn_wh - 1) // bn_wh)  # W.H: dual-N loop, 2 MMAs of N = 2 bn per k-step
    assert slots == 18 and (c_h, c_w) == (144, 144) and c_wh <= 148                # every launch is one wave of <= 148 CTAs
    assert lib.gccnmf_klnmf_tile_plan(148, 513, 622, 128, out) == 0               # config 1: few tiles -> k-splits for the numerator
    assert out[3] >= 2 and max(out[5], out[6], out[7]) <= 148
    assert lib.gccnmf_klnmf_tile_plan(148, 1025, 37494, 4096, out) == 0           # config 4: many waves, widest tiles
    assert out[0] == 256 and out[1] in (208, 256) and out[3] == 1
    for bn in (out[0], out[1], out[2]):
        assert bn in (112, 128, 176, 208, 256)
    assert lib.gccnmf_klnmf_tile_plan(148, 513, 3744, 1020, out) < 0              # K % 8 != 0: float32 SIMT path
    assert lib.gccnmf_klnmf_tile_plan(0, 513, 3744, 1024, out) < 0
    # fewer SMs -> the planner may not use more CTAs than a wave when a single-wave choice exists
    assert lib.gccnmf_klnmf_tile_plan(132, 513, 3744, 1024, out) == 0
    assert out[7] <= 132


def test_pull_exchange_buffer_layout(lib):
    """Size of the symmetric buffer of the pull exchange (gccnmf_klnmf_step_pull): 2 x (numerator + packed row sums), 2 x row-sum slots,
    the slice-owner buffer, the arrival counters and one flag per 32 x 128 tile of U; identical on every rank when built from the largest
    shard (host logic only)."""
    F, T2, K = 513, 3744, 1024
    n = lib.gccnmf_klnmf_pull_buffer_floats(F, T2, K)
    slots = (T2 + 127) // 128
    tiles = ((F + 31) // 32) * ((K + 127) // 128)
    assert n == 2 * (F * K + K) + 2 * slots * K + F * K + 64 + (tiles + 63) // 64 * 64
    assert lib.gccnmf_klnmf_pull_buffer_floats(F, T2 + 2, K) >= n           # uneven shards: every rank uses the largest 2T
    assert lib.gccnmf_klnmf_pull_buffer_floats(1025, 4688, 4096) > 3 * 1025 * 4096
    assert lib.gccnmf_klnmf_pull_buffer_floats(0, T2, K) == 0


def test_documented_options_are_the_accepted_ones():
    """Every option gccnmf_set_option accepts (csrc/api.cu) is listed in the header's option block, and vice versa."""
    api = open(os.path.join(ROOT, 'gcc-nmf_b200', 'csrc', 'api.cu')).read()
    accepted = set(re.findall(r'strcmp\(name, "([a-z_0-9]+)"\)', api))
    hdr = open(os.path.join(ROOT, 'include', 'gccnmf_b200.h')).read()
    block = hdr[hdr.index('Options (A/B switches'):hdr.index('GCCNMF_API int gccnmf_set_option')]
    documented = set(re.findall(r'^ \*   "([a-z_0-9]+)"', block, re.M))
    assert accepted == documented, (sorted(accepted - documented), sorted(documented - accepted))
