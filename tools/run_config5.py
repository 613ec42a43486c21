"""BASELINE.json configs[4]: low-latency asymmetric-window path -- 1024-sample analysis window, 64-sample hop, K = 256, 128 TDOAs,
one independent stream per GPU (no exchange of any kind: "replicas only"):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 tools/run_config5.py [--json out.json]

Every rank runs notebooks/lowLatencySpeechEnhancement.ipynb's frame loop (:511-584, asymmetric analysis window of :371-380 with
m = 64) over its own 30 s stream through `online.performOnlineSpeechEnhancement` (all frames in one batch: the loop's only carried
state is a prefix maximum) with a dictionary pre-learnt on the stream's own spectrogram (K = 256, 100 iterations; the CHiME matrix
of the reference is 513-bin / K <= 1024 and absent on the GPU box).  Reported: frames/s per stream (host call, wall clock, including
the host <-> device copies of the drop-in API) and the aggregate over the ranks (sum: the streams are independent)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--seconds', type=float, default=30.0)
    ap.add_argument('--json', default=None)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    import gcc_nmf_b200.gccNMFFunctions as fn
    from gcc_nmf_b200._lib import default_handle
    from gcc_nmf_b200.online import getAsymmetricAnalysisWindow, getAsymmetricSynthesisWindow, performOnlineSpeechEnhancement
    from gcc_nmf_b200.synth import synthetic_stereo
    default_handle(local)
    sr, N, hop, K, D, m = 16000, 1024, 64, 256, 128, 64
    x = synthetic_stereo(args.seconds, seed=1234 + rank)
    X = fn.computeComplexMixtureSpectrogram(x, N, 256, np.hanning)
    W, _ = fn.performKLNMF(np.concatenate(np.abs(X), axis=-1), K, 100, 0)
    analysis, synthesis = getAsymmetricAnalysisWindow(N, m, 0), getAsymmetricSynthesisWindow(N, m, 0)
    run = lambda: performOnlineSpeechEnhancement(x, sr, W, analysis, synthesis, hop, D, 0.1, 0.05 * D, gainPerFrame=False, device=local)   # noqa: E731
    res = run()
    frames = res[4].shape[0]
    times = []
    for _ in range(args.steps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        res = run()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    best = min(times)
    t = torch.tensor([frames / best], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if rank == 0:
        line = {'workload': 'BASELINE.json configs[4]: low-latency asymmetric window, N=1024 hop=64 K=256 D=128, %.0f s per stream, one stream per GPU' % args.seconds,
                'n_gpus': world, 'frames_per_stream': int(frames), 'seconds_per_pass_rank0': best, 'frames_per_s_rank0': frames / best,
                'frames_per_s_all_streams': float(t.item()), 'realtime_factor_rank0': args.seconds / best, 'finite': bool(np.isfinite(res[2]).all()),
                'all_seconds_rank0': times}
        print(json.dumps(line))
        if args.json:
            json.dump(line, open(args.json, 'w'), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
