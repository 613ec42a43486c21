"""Per-chunk latency of the real-time path -- BASELINE.json configs[2]: online RT-GCC-NMF with per-frame coefficient inference,
pretrained-size dictionary K = 1024, 512-FFT, hop 128, 64 TDOAs, one frame (= one hop of new audio) per chunk.

    python tools/rt_latency.py [--chunks 2000] [--K 1024] [--N 512] [--hop 128] [--D 64] [--inference 10] [--json out.json]

Every chunk is one audio block of `hop` new samples per channel handed to the fused block path (csrc/rt.cu): pinned host block
-> H2D -> rings + FFT + PHAT + GCC-NMF argmax mask [+ n H-only KL updates] + Wiener filter + inverse FFT + overlap-add + block
emit -> D2H, as ONE CUDA-graph launch and one stream synchronisation -- what `GCCNMFProcess.run` sees per block
(gccNMF/realtime/gccNMFProcessor.py:94-101).  Three clocks per chunk:
  wall     time.perf_counter around `processBlock` (host copy into the pinned buffer, graph launch, synchronisation)
  device   CUDA events recorded on the engine's stream around the graph launch (H2D + kernels + D2H)
  frames   the same for `GCCNMFProcessor.processFrames` (no rings: windowed frames in, frames out), kernel-by-kernel launches
The dictionary is random (its spectral shape does not change the work); the input is the synthetic two-source mixture.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def percentiles(a):
    a = np.asarray(a)
    return {'p50_ms': float(np.percentile(a, 50)), 'p90_ms': float(np.percentile(a, 90)), 'p99_ms': float(np.percentile(a, 99)),
            'max_ms': float(a.max()), 'mean_ms': float(a.mean())}


def run(K, N, hop, D, inference, chunks, warmup, use_graph=True, frames_per_block=1):
    import torch
    from gcc_nmf_b200.realtime.gccNMFProcessor import GCCNMFProcessor
    from gcc_nmf_b200.realtime.utils import CircularBuffer
    from gcc_nmf_b200.synth import synthetic_stereo
    sr, nT = 16000, frames_per_block
    B = nT * hop
    rng = np.random.default_rng(0)
    W = (rng.random((N // 2 + 1, K)) ** 3).astype(np.float32)
    proc = GCCNMFProcessor(sr, N, nT, {'Pretrained': {K: W}}, 'Pretrained', K, 0, 0.1, True, 6,
                           gccPHATHistory=None, tdoaHistory=None, coefficientInferenceIterations=inference)
    proc.numTDOAs = D
    proc.reset()
    proc.setTargetTDOARange(10.0, 5.0, 2.0, 0.0)
    total = warmup + chunks
    x = synthetic_stereo(max(2.0, (total * B + N) / sr + 0.1))
    wall, dev = [], []
    y = None
    for c in range(total):
        block = x[:, c * B:(c + 1) * B]
        if c == 0:
            proc.processBlock(block, hop, B, useGraph=use_graph)       # builds the engine + graph
            continue
        stream = proc.engine.stream
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        y = proc.processBlock(block, hop, B, useGraph=use_graph)
        e1.record(stream)
        t1 = time.perf_counter()
        e1.synchronize()
        if c >= warmup:
            wall.append((t1 - t0) * 1e3)
            dev.append(e0.elapsed_time(e1))
    assert y.shape == (2, B) and np.isfinite(y).all()
    # processFrames (drop-in entry, no rings)
    window = np.sqrt(np.hamming(N).astype(np.float32))[None, :, None]
    nT = 1
    proc2 = GCCNMFProcessor(sr, N, nT, {'Pretrained': {K: W}}, 'Pretrained', K, 0, 0.1, True, 6, coefficientInferenceIterations=inference)
    proc2.numTDOAs = D
    proc2.reset()
    wall_f = []
    for c in range(min(total, 600)):
        frames = (x[:, c * hop:c * hop + N][:, :, None] * window).astype(np.float32)
        t0 = time.perf_counter()
        proc2.processFrames(frames)
        if c >= 100:
            wall_f.append((time.perf_counter() - t0) * 1e3)
    launches = 5 + 2 * inference
    return {'config': {'fft': N, 'hop': hop, 'K': K, 'D': D, 'frames_per_block': frames_per_block, 'block_samples': B, 'inference_iterations': inference, 'chunks': chunks,
                       'cuda_graph': bool(use_graph), 'kernels_per_block': launches},
            'block_wall': percentiles(wall), 'block_device': percentiles(dev), 'processFrames_wall': percentiles(wall_f),
            'budget_ms': B / sr * 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--chunks', type=int, default=2000)
    ap.add_argument('--warmup', type=int, default=100)
    ap.add_argument('--K', type=int, default=1024)
    ap.add_argument('--N', type=int, default=512)
    ap.add_argument('--hop', type=int, default=128)
    ap.add_argument('--D', type=int, default=64)
    ap.add_argument('--inference', type=int, nargs='*', default=[0, 10])
    ap.add_argument('--frames-per-block', type=int, nargs='*', default=[1, 2],
                    help='1: one hop per block (configs[2] per-frame latency); 2: block = 2 hops, the smallest block whose ring output is a complete '
                         'overlap-add at N = 4 hop (the ring emits block [-3B, -2B), utils.py:115)')
    ap.add_argument('--json', default=None)
    args = ap.parse_args()
    results = []
    for n in args.inference:
      for fpb in args.frames_per_block:
        for graph in (True, False):
            r = run(args.K, args.N, args.hop, args.D, n, args.chunks, args.warmup, use_graph=graph, frames_per_block=fpb)
            results.append(r)
            print('%d-FFT hop %d K=%d D=%d, %d frame(s) per block, %d inference iterations, graph=%s: block wall p50 %.3f p99 %.3f max %.3f ms | device p50 %.3f p99 %.3f ms | '
                  'processFrames wall p50 %.3f ms | budget %.1f ms' % (
                      args.N, args.hop, args.K, args.D, fpb, n, graph, r['block_wall']['p50_ms'], r['block_wall']['p99_ms'], r['block_wall']['max_ms'],
                      r['block_device']['p50_ms'], r['block_device']['p99_ms'], r['processFrames_wall']['p50_ms'], r['budget_ms']), flush=True)
    if args.json:
        json.dump(results, open(args.json, 'w'), indent=1)


if __name__ == '__main__':
    main()
