"""Per-chunk latency of the real-time path (BASELINE.json configs[2]: online RT-GCC-NMF, pretrained W K=1024, 512-FFT,
hop 128, 64 TDOAs, one frame per chunk): p50 / p99 of `GCCNMFProcessor.processFrames` over N chunks after warm-up.

    python tools/rt_latency.py [--chunks 2000] [--K 1024] [--N 512] [--D 64] [--frames-per-chunk 1]

Two clocks per chunk: the wall clock around the call (host -> device copy of the windowed frames, the kernels, device ->
host copy of the result: what `GCCNMFProcess.run` would see, gccNMF/realtime/gccNMFProcessor.py:94-101) and CUDA events
around the same region (device time only).  The dictionary is random (spectral shape does not change the work); the input is
the synthetic two-source mixture of `gcc_nmf_b200.synth`.  Written at the end of round 1 with no GPU time left: it has
not produced a committed number yet.
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--chunks', type=int, default=2000)
    ap.add_argument('--warmup', type=int, default=50)
    ap.add_argument('--K', type=int, default=1024)
    ap.add_argument('--N', type=int, default=512)
    ap.add_argument('--D', type=int, default=64)
    ap.add_argument('--frames-per-chunk', type=int, default=1)
    args = ap.parse_args()
    import torch
    from gcc_nmf_b200.realtime.gccNMFProcessor import GCCNMFProcessor
    from gcc_nmf_b200.realtime.utils import CircularBuffer
    from gcc_nmf_b200.synth import synthetic_stereo

    sr, N, nT, K, D = 16000, args.N, args.frames_per_chunk, args.K, args.D
    hop = N // 4
    rng = np.random.default_rng(0)
    W = (rng.random((N // 2 + 1, K)) ** 3).astype(np.float32)
    proc = GCCNMFProcessor(sr, N, nT, {'Pretrained': {K: W}}, 'Pretrained', K, 0, 0.1, True, 6,
                           gccPHATHistory=CircularBuffer((D, 128)), tdoaHistory=CircularBuffer((1, 128)))
    proc.numTDOAs = D
    proc.reset()
    proc.setTargetTDOARange(10.0, 5.0, 2.0, 0.0)
    total = args.warmup + args.chunks
    x = synthetic_stereo(max(2.0, (total * nT * hop + N) / sr + 0.1))
    window = np.sqrt(np.hamming(N).astype(np.float32))[None, :, None]
    wall, dev = [], []
    for c in range(total):
        s = c * nT * hop
        frames = np.stack([x[:, s + i * hop:s + i * hop + N] for i in range(nT)], axis=-1) * window
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        y = proc.processFrames(frames.astype(np.float32))
        e1.record()
        e1.synchronize()
        t1 = time.perf_counter()
        if c >= args.warmup:
            wall.append((t1 - t0) * 1e3)
            dev.append(e0.elapsed_time(e1))
    assert y.shape == (2, N, nT) and np.isfinite(y).all()
    pct = lambda a, q: float(np.percentile(a, q))   # noqa: E731
    print('config: %d-FFT hop %d, K=%d, %d TDOAs, %d frame(s) per chunk, %d chunks' % (N, hop, K, D, nT, args.chunks))
    print('wall clock per chunk (host copies + kernels): p50 %.3f ms  p99 %.3f ms  max %.3f ms' % (pct(wall, 50), pct(wall, 99), max(wall)))
    print('CUDA events per chunk:                        p50 %.3f ms  p99 %.3f ms' % (pct(dev, 50), pct(dev, 99)))
    print('real-time budget per chunk (hop / sample rate): %.3f ms' % (nT * hop / sr * 1e3))


if __name__ == '__main__':
    main()
