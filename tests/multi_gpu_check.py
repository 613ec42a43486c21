"""Run under torchrun on >= 2 GPUs (not collected by pytest: needs NCCL + several devices):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py

Checks that the frame-sharded enhancement pipeline (joint dictionary, one all-reduce per KL-NMF
iteration, iSTFT seam exchange) reproduces the single-GPU pipeline run on the whole recording."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank, world = dist.get_rank(), dist.get_world_size()
    from gcc_nmf_b200.distributed import ShardedGCCNMFPipeline
    from gcc_nmf_b200.pipeline import GCCNMFPipeline
    from gcc_nmf_b200.synth import synthetic_stereo
    cfg = dict(sampleRate=16000, windowSize=1024, hopSize=256, numTDOAs=64, microphoneSeparationInMetres=0.1,
               dictionarySize=256, numIterations=20)
    clip = 4.0
    sp = ShardedGCCNMFPipeline(device=local, clip_seconds=clip, **cfg)
    x_local = torch.from_numpy(sp.local_samples()).to(sp.h.device)
    r = sp.enhance(x_local)
    y_local = r['targetSignalEstimates'][0].cpu().numpy()
    sizes = [None] * world
    dist.all_gather_object(sizes, y_local.shape[1])
    ys = [None] * world
    dist.all_gather_object(ys, y_local)
    Ws = [None] * world
    dist.all_gather_object(Ws, r['W'].cpu().numpy())
    if rank == 0:
        full = np.concatenate([synthetic_stereo(clip, seed=1234 + c) for c in range(world)], axis=1)
        single = GCCNMFPipeline(device=local, handle=sp.h, **cfg)
        r1 = single.enhance(single.h.to_device(full))
        y1 = r1['targetSignalEstimates'][0].cpu().numpy()
        y = np.concatenate(ys, axis=1)
        W1 = r1['W'].cpu().numpy()

        def rel(a, b):
            return float(np.linalg.norm(a - b) / np.linalg.norm(b))
        print('collective:', sp.collective)
        print('world', world, 'frames', sp.total_frames, 'target', r['targetTDOAIndexes'], r1['targetTDOAIndexes'])
        print('W identical across ranks:', all(np.array_equal(Ws[0], w) for w in Ws))
        print('rel W sharded vs single: %.3e' % rel(Ws[0], W1))
        print('signal shape', y.shape, y1.shape, 'rel signal: %.3e' % rel(y, y1))
        ok = (y.shape == y1.shape and r['targetTDOAIndexes'] == r1['targetTDOAIndexes'] and rel(Ws[0], W1) < 1e-4)
        print('MULTI_GPU_CHECK', 'PASS' if ok else 'FAIL')
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
