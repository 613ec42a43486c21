"""BASELINE.json configs[3]: batch dictionary learning -- 10 min synthetic stereo @16 kHz, 2048-FFT (hop 512: librosa's default
N / 4, the config does not name one), K = 4096, 200 KL-NMF iterations, frame-sharded over the GPUs of one box:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/run_config4.py [--json out.json]

Each rank holds 600 s / world of the recording (the synthetic generator is seeded per 75 s clip, so the recording is the same for
every world size that divides 8); ONE dictionary is learnt jointly: per iteration every rank contributes its (F K + K)-float
numerator (16.8 MB) and the sum is formed inside the NVSwitch (multimem.ld_reduce in the W-update kernel).  The whole enhancement
flow runs (STFT ... iSTFT); the KL-NMF stage is what configs[3] is about and is reported separately.  Times are CUDA events, max
over ranks; one warm-up pass, `--steps` timed passes."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--steps', type=int, default=2)
    ap.add_argument('--seconds', type=float, default=600.0)
    ap.add_argument('--K', type=int, default=4096)
    ap.add_argument('--iterations', type=int, default=200)
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
    cfg = dict(sampleRate=16000, windowSize=2048, hopSize=512, numTDOAs=64, microphoneSeparationInMetres=0.1,
               dictionarySize=args.K, numIterations=args.iterations)
    if world > 1:
        from gcc_nmf_b200.distributed import ShardedGCCNMFPipeline
        pipe = ShardedGCCNMFPipeline(device=local, clip_seconds=args.seconds / world, **cfg)
        x = torch.from_numpy(pipe.local_samples()).to(pipe.h.device)
        total_frames = pipe.total_frames
    else:
        from gcc_nmf_b200.pipeline import GCCNMFPipeline
        from gcc_nmf_b200.synth import synthetic_stereo
        pipe = GCCNMFPipeline(device=local, **cfg)
        x = pipe.h.to_device(synthetic_stereo(args.seconds))
        total_frames = pipe.num_frames(x.shape[1])
    torch.cuda.synchronize()
    r = pipe.enhance(x)
    torch.cuda.synchronize()
    step_ms, nmf_ms = [], []
    for _ in range(args.steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = pipe.enhance(x, collect_stage_times=True)
        e1.record()
        e1.synchronize()
        t = torch.tensor([e0.elapsed_time(e1), pipe.stage_times_ms()['nmf']], dtype=torch.float64, device=pipe.h.device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_ms.append(float(t[0])); nmf_ms.append(float(t[1]))
    W = r['W']
    finite = bool(torch.isfinite(W).all() and torch.isfinite(r['targetSignalEstimates']).all())
    if rank == 0:
        F, K, I = cfg['windowSize'] // 2 + 1, args.K, args.iterations
        ms, nm = float(np.min(step_ms)), float(np.min(nmf_ms))
        flops = 16.0 * F * K * total_frames * I
        line = {'workload': 'BASELINE.json configs[3]: %.0f s synthetic stereo, 2048-FFT hop 512, K=%d, %d KL-NMF iterations, frame-sharded' % (args.seconds, K, I),
                'n_gpus': world, 'total_frames': int(total_frames), 'frames_per_rank': int(total_frames // world), 'steps': args.steps,
                'ms_per_pass': ms, 'frames_per_s': total_frames / (ms * 1e-3), 'nmf_ms': nm, 'nmf_ms_per_iteration': nm / I,
                'nmf_algorithmic_tflops_all_gpus': flops / (nm * 1e-3) / 1e12, 'exchange_bytes_per_iteration': (F * K + K) * 4,
                'collective': getattr(pipe, 'collective', None), 'finite': finite, 'all_step_ms': step_ms, 'all_nmf_ms': nmf_ms}
        print(json.dumps(line))
        if args.json:
            json.dump(line, open(args.json, 'w'), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
