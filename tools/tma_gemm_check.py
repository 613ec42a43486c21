"""GPU diagnostics for the TMA-fed plane GEMM (gcc-nmf_b200/csrc/tma_gemm.cuh) and the KL-NMF path built on it.

  python tools/tma_gemm_check.py gemm      every operand-layout combination against a float64 product
  python tools/tma_gemm_check.py nmf       KL-NMF, TMA path vs oracle and vs the float32 SIMT path (small + config 2)
  python tools/tma_gemm_check.py time      KL-NMF stage time at config 2: float32 SIMT path, TMA path, TMA + PDL
  python tools/tma_gemm_check.py all       each of the above in its own process (a trap in one does not hide the others)
"""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def handle():
    from gcc_nmf_b200._lib import default_handle
    h = default_handle()
    if os.environ.get('GEMM_PAIR'):          # cta_group::2 CTA pairs for every contraction whose tile shape allows them
        h.set_option('gemm_pair', int(os.environ['GEMM_PAIR']))
    return h


def gemm():
    import torch
    h = handle()
    shapes = [(128, 128, 64), (128, 256, 96), (256, 300, 1024), (513, 3744, 1024), (1024, 3744, 513), (1024, 513, 3744), (200, 130, 70)]
    bad = 0
    for cluster in (11, 22, 12, 21):
      h.set_option('gemm_cluster', cluster)
      for a_mn, b_mn in ((False, False), (True, False), (True, True)):
        for M, N, Kc in (shapes if cluster in (11, 22) else shapes[2:6]):
            for tile_n in (128, 176, 208, 256):
                for splits in ((1, 3) if Kc >= 1024 else (1,)):
                    g = torch.Generator(device='cpu').manual_seed(M * 7 + N * 3 + Kc)
                    A = torch.randn(M, Kc, generator=g)
                    B = torch.rand(N, Kc, generator=g)
                    Ad, Bd = A.to(h.device), B.to(h.device)
                    Ain = Ad.T.contiguous() if a_mn else Ad
                    Bin = Bd.T.contiguous() if b_mn else Bd
                    try:
                        DT = h.gemm_planes(Ain, Bin, a_mn, b_mn, tile_n=tile_n, splits=splits)
                        torch.cuda.synchronize()
                    except Exception as e:   # noqa: BLE001
                        print('FAIL a_mn=%d b_mn=%d %s tile %d splits %d: %s' % (a_mn, b_mn, (M, N, Kc), tile_n, splits, e))
                        return 1
                    D = DT.sum(0).T.double()
                    ref = Ad.double() @ Bd.double().T
                    scale = Ad.abs().double() @ Bd.abs().double().T
                    err = ((D - ref).abs() / scale)
                    e = err.max().item()
                    ok = e < 2e-5
                    bad += 0 if ok else 1
                    msg = 'cluster %d a_mn=%d b_mn=%d M=%d N=%d K=%d tile %d splits %d: max err/|a||b| %.2e %s' % (cluster, a_mn, b_mn, M, N, Kc, tile_n, splits, e, 'ok' if ok else 'BAD')
                    if not ok:
                        # structure of the error: which rows / columns are wrong
                        wrong = err > 2e-5
                        rows = wrong.any(1).nonzero().flatten().tolist()
                        cols = wrong.any(0).nonzero().flatten().tolist()
                        msg += ' | wrong %.1f%% rows %s.. cols %s.. | D[0,:4]=%s ref=%s' % (
                            100 * wrong.float().mean().item(), rows[:12], cols[:12], D[0, :4].tolist(), ref[0, :4].tolist())
                    print(msg, flush=True)
    h.set_option('gemm_cluster', -1)
    print('gemm: %d bad' % bad)
    return 1 if bad else 0


def _rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def nmf():
    import torch
    from oracle import gccnmf_oracle as orc
    h = handle()
    rc = 0
    for (F, T2, K, iters) in ((257, 512, 64, 30), (513, 3744, 1024, 100)):
        rng = np.random.default_rng(5)
        V = (rng.random((F, T2)) ** 3 + 1e-3).astype(np.float32)
        W0, H0 = orc.initKLNMF(F, T2, K)
        t0 = time.time()
        Wo, Ho = orc.performKLNMF(V, K, iters, 0, W0=W0, H0=H0)
        res = {}
        for name, tma in (('simt', 0), ('tma', 1)):
            h.set_option('force_simt_nmf', 1 - tma)
            W, H = h.to_device(W0.copy()), h.to_device(H0.copy())
            h.klnmf(h.to_device(V), W, H, iters)
            torch.cuda.synchronize()
            res[name] = (W.cpu().numpy(), H.cpu().numpy())
            eW, eH = _rel(res[name][0], Wo), _rel(res[name][1], Ho)
            ok = eW < 1e-4 and eH < 1e-4
            rc |= 0 if ok else 1
            print('KL-NMF %s F=%d T2=%d K=%d %d it: rel W %.2e rel H %.2e %s (oracle %.1fs)' % (name, F, T2, K, iters, eW, eH, 'ok' if ok else 'BAD', time.time() - t0), flush=True)
        # 1 iteration: tight comparison of the two paths
        out = {}
        for name, tma in (('simt', 0), ('tma', 1)):
            h.set_option('force_simt_nmf', 1 - tma)
            W, H = h.to_device(W0.copy()), h.to_device(H0.copy())
            h.klnmf(h.to_device(V), W, H, 1)
            out[name] = (W.cpu().numpy(), H.cpu().numpy())
        print('  1 iteration tma vs simt: rel W %.2e rel H %.2e' % (_rel(out['tma'][0], out['simt'][0]), _rel(out['tma'][1], out['simt'][1])))
        # fixed dictionary
        h.set_option('force_simt_nmf', 0)
        Hi = h.to_device(H0.copy())
        h.klnmf(h.to_device(V), h.to_device(Wo), Hi, 5, update_W=False)
        Href = H0.copy()
        denom = np.sum(Wo, axis=0)[:, None] + np.float32(1e-16)
        for _ in range(5):
            Href *= np.dot(Wo.T, V / np.dot(Wo, Href)) / denom
        e = _rel(Hi.cpu().numpy(), Href)
        rc |= 0 if e < 5e-5 else 1
        print('  fixed dictionary, 5 H updates: rel H %.2e' % e, flush=True)
    return rc


def timing():
    import torch
    from oracle import gccnmf_oracle as orc
    h = handle()
    F, T2, K, iters = 513, 3744, 1024, 100
    rng = np.random.default_rng(5)
    V = h.to_device((rng.random((F, T2)) ** 3 + 1e-3).astype(np.float32))
    W0, H0 = orc.initKLNMF(F, T2, K)
    W0d, H0d = h.to_device(W0), h.to_device(H0)
    variants = (('float32 simt', 0, 0, -1), ('tma no pdl, auto clusters', 1, 0, -1), ('tma+pdl, no clusters', 1, 1, 11), ('tma+pdl, clusters 1x2', 1, 1, 12),
                ('tma+pdl, clusters 2x1', 1, 1, 21), ('tma+pdl, clusters 2x2', 1, 1, 22), ('tma+pdl, auto clusters', 1, 1, -1))
    if os.environ.get('TIME_VARIANTS') == 'short':
        variants = (('tma+pdl, auto clusters', 1, 1, -1),)
    for name, tma, pdl, cl in variants:
        h.set_option('force_simt_nmf', 1 - tma)
        h.set_option('nmf_pdl', pdl)
        h.set_option('gemm_cluster', cl)
        t_cpu = 0.0
        ms = []
        for rep in range(4):
            W, H = W0d.clone(), H0d.clone()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            tc = time.perf_counter()
            h.klnmf(V, W, H, iters)
            t_cpu = time.perf_counter() - tc
            e1.record()
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
        print('KL-NMF config 2 (%s): %s ms per 100 iterations (host time of the call %.2f ms); W finite %s' % (
            name, ['%.2f' % m for m in ms], t_cpu * 1e3, bool(torch.isfinite(W).all())), flush=True)
    h.set_option('nmf_pdl', 1)
    h.set_option('gemm_cluster', -1)
    # per-GEMM CTA phase stamps
    g = torch.Generator(device='cpu').manual_seed(1)
    for (a_mn, b_mn, M, N, Kc, tile, splits, label) in ((0, 0, 513, 3744, 1024, 128, 1, 'G1/G3'), (1, 0, 1024, 3744, 513, 208, 1, 'G2'),
                                                       (1, 1, 1024, 513, 3744, 176, 6, 'G4')):
        A = torch.rand(M, Kc, generator=g).to(h.device)
        B = torch.rand(N, Kc, generator=g).to(h.device)
        Ain = A.T.contiguous() if a_mn else A
        Bin = B.T.contiguous() if b_mn else B
        ctas = ((N + tile - 1) // tile) * ((M + 127) // 128 + 1) * splits
        stamps = torch.zeros(ctas * 8, dtype=torch.int64, device=h.device)
        for _ in range(3):
            h.gemm_planes(Ain, Bin, bool(a_mn), bool(b_mn), tile_n=tile, splits=splits, timing=stamps)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            h.gemm_planes(Ain, Bin, bool(a_mn), bool(b_mn), tile_n=tile, splits=splits)
        e1.record()
        e1.synchronize()
        s = stamps.cpu().numpy().reshape(-1, 8)[:, 1:7]
        s = s[s[:, 0] > 0]
        s = s[s[:, 5] > 0]
        s = s[s[:, 2] > 0]            # (a cta_group::2 pair's non-leader issues no MMA: leaders only)
        d = lambda a, b: float(np.median(s[:, b] - s[:, a]))   # noqa: E731
        print('%s M=%d N=%d K=%d tile %d splits %d: %.1f us per call incl. operand split; CTA medians: start->first full %.0f, '
              'mma issue loop %.0f, last issue->accum done %.0f, epilogue %.0f, total %.0f cycles (%d CTAs)' % (
                  label, M, N, Kc, tile, splits, e0.elapsed_time(e1) / 20 * 1e3, d(0, 1), d(1, 2), d(2, 4), d(4, 5), d(0, 5), len(s)), flush=True)
    return 0


def variants():
    """KL-NMF stage time at config 2 over runtime switches: k-splits of the W-update numerator summed inside clusters (vs slabs),
    persisting L2 window over G^T, streaming hints for the k-split slabs; plus the result's distance from the default build.  (Measured earlier in the
    round: operand preload and CTA pairs change nothing; 104-column dual-N tiles beat 112 and 128.)"""
    import torch
    from oracle import gccnmf_oracle as orc
    h = handle()
    F, T2, K, iters = 513, 3744, 1024, 100
    rng = np.random.default_rng(5)
    V = h.to_device((rng.random((F, T2)) ** 3 + 1e-3).astype(np.float32))
    W0, H0 = orc.initKLNMF(F, T2, K)
    W0d, H0d = h.to_device(W0), h.to_device(H0)
    ref = None
    for split2 in (0, 1, 0, 1):
        h.set_option('wh_split2', split2)
        ms = []
        for rep in range(3):
            W, H = W0d.clone(), H0d.clone()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            h.klnmf(V, W, H, iters)
            e1.record()
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
        Wn = W.cpu().numpy()
        if ref is None:
            ref = Wn
            Wo, Ho = orc.performKLNMF(V.cpu().numpy(), K, iters, 0, W0=W0, H0=H0)
        print('wh_split2 %d: %s ms per 100 iterations | rel W vs first variant %.2e, vs oracle %.2e finite %s' % (
            split2, ['%.2f' % m for m in ms], _rel(Wn, ref), _rel(Wn, Wo), bool(np.isfinite(Wn).all())), flush=True)
    h.set_option('wh_split2', 0)
    h.set_option('w_cluster_reduce', 1)
    h.set_option('wh_tile', 0)
    h.set_option('gemm_preload', 1)
    h.set_option('gemm_streaming', 0)
    h.set_option('gemm_pair', -1)
    return 0


def stamps():
    """%globaltimer / clock64 stamps of every CTA of the plane GEMMs inside the live KL-NMF loop (config 2)."""
    import ctypes
    import torch
    from oracle import gccnmf_oracle as orc
    h = handle()
    F, T2, K = 513, 3744, 1024
    rng = np.random.default_rng(5)
    V = h.to_device((rng.random((F, T2)) ** 3 + 1e-3).astype(np.float32))
    W0, H0 = orc.initKLNMF(F, T2, K)
    h.set_option('force_simt_nmf', 0)
    WH = int(os.environ.get('WH_TILE', '128'))
    h.set_option('wh_tile', WH)
    split2 = int(os.environ.get('WH_SPLIT2', '0'))
    h.set_option('wh_split2', split2)
    h.set_option('gemm_cluster', int(os.environ.get('GEMM_CLUSTER', '-1')))
    for pdl in (0, 1):
        h.set_option('nmf_pdl', pdl)
        W, H = h.to_device(W0), h.to_device(H0)
        h.klnmf(V, W, H, 5)
        buf = torch.zeros(4 << 20, dtype=torch.int64, device=h.device)
        torch.cuda.synchronize()
        h.lib.gccnmf_debug_timing(h.h, buf.data_ptr(), 1)
        h.klnmf(V, W, H, 3)
        torch.cuda.synchronize()
        used = h.lib.gccnmf_debug_timing(h.h, None, 1)
        s = buf.cpu().numpy()[:used].reshape(-1, 8)
        nt = 2 * ((T2 + 207) // 208) if split2 else (T2 + WH - 1) // WH
        grids = [('G1', nt * 4, 120), ('G2', 18 * 8, 144), ('G3', nt * 4, 120), ('G4', 3 * 8 * 6, 144), ('W update', 1, 0)] * 3
        off, prev_end = 0, None
        print('pdl=%d: %d CTA records' % (pdl, len(s)))
        for name, ctas, tc in grids:
            k = s[off:off + ctas]
            off += ctas
            if name == 'W update':          # one record: [0] start (after the dependency wait), [2] operands loaded, [7] end of CTA 0
                print('W update: CTA 0 %.1f us (loads %.1f us) | gap after previous GEMM %.1f us' % (
                    (k[0, 7] - k[0, 0]) / 1e3, (k[0, 2] - k[0, 0]) / 1e3, (k[0, 0] - prev_end) / 1e3), flush=True)
                prev_end = k[0, 7]
                continue
            k = k[k[:, 0] > 0]
            t0 = k[:, 0].min()
            st, en = k[:, 0] - t0, k[:, 7] - t0
            is_tc = (k[:, 1] > 0) & (k[:, 2] > 0) & (k[:, 3] > 0)      # the CTAs that issue MMAs (a pair's leader; every CTA otherwise)
            follower = (k[:, 1] > 0) & (k[:, 2] > 0) & (k[:, 3] == 0)  # the non-leader CTA of a cta_group::2 pair
            dur = (k[:, 7] - k[:, 0])
            cyc = lambda a, b: np.median((k[is_tc, b] - k[is_tc, a]))   # noqa: E731
            msg = '%s: span %.1f us | CTA start offset p50 %.1f max %.1f us | tc CTA dur p50 %.1f max %.1f us' % (
                name, en.max() / 1e3, np.median(st) / 1e3, st.max() / 1e3, np.median(dur[is_tc]) / 1e3, dur[is_tc].max() / 1e3)
            msg += ' | cycles: prologue %.0f, main %.0f, drain %.0f, epilogue %.0f, producer done at %.0f, total %.0f' % (
                cyc(1, 2), cyc(2, 3), cyc(3, 5), cyc(5, 6), cyc(1, 4), cyc(1, 6))
            if follower.any():
                fc = lambda a, b: np.median((k[follower, b] - k[follower, a]))   # noqa: E731
                msg += ' | pair followers (%d): prologue %.0f, producer done at %.0f, accumulator at %.0f, epilogue %.0f' % (
                    int(follower.sum()), fc(1, 2), fc(1, 4), fc(1, 5), fc(5, 6))
            if prev_end is not None:
                msg += ' | gap after previous GEMM %.1f us' % ((t0 - prev_end) / 1e3)
            prev_end = k[:, 7].max()
            print(msg, flush=True)
    h.set_option('nmf_pdl', 0)
    return 0


def prof():
    """A short KL-NMF run on the TMA path for an ncu launch list (--cache-control none: warm L2 like the live loop)."""
    import torch
    from oracle import gccnmf_oracle as orc
    h = handle()
    F, T2, K = 513, 3744, 1024
    rng = np.random.default_rng(5)
    V = h.to_device((rng.random((F, T2)) ** 3 + 1e-3).astype(np.float32))
    W0, H0 = orc.initKLNMF(F, T2, K)
    h.set_option('force_simt_nmf', 0)
    h.set_option('nmf_pdl', int(os.environ.get('PROF_PDL', '0')))
    W, H = h.to_device(W0), h.to_device(H0)
    h.klnmf(V, W, H, int(os.environ.get('PROF_ITERS', '6')))
    torch.cuda.synchronize()
    return 0


if __name__ == '__main__':
    what = sys.argv[1] if len(sys.argv) > 1 else 'all'
    if what == 'all':
        rc = 0
        for part in ('gemm', 'nmf', 'time'):
            print('==== %s' % part, flush=True)
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), part], timeout=600)
                rc |= r.returncode
            except subprocess.TimeoutExpired:
                print('TIMEOUT in %s' % part)
                rc |= 1
        sys.exit(rc)
    sys.exit({'gemm': gemm, 'nmf': nmf, 'time': timing, 'prof': prof, 'stamps': stamps, 'variants': variants}[what]())
