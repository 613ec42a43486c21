"""Where a sharded KL-NMF iteration spends its time (run under torchrun, 1 rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tools/multi_stamps.py

For each form of the numerator exchange (two-shot / one-shot inside the switch, NCCL all-reduce) and, for reference, the same loop with
no exchange at all (one C call per iteration, and the fused single-call loop): CUDA-event time per iteration and, from the
%globaltimer stamps every plane-GEMM CTA records (gccnmf_debug_timing), the iteration period, the span of each contraction and the
gap between the end of the numerator contraction and the start of the next iteration's first contraction (exchange + W update)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from gcc_nmf_b200 import distributed as gd
    from gcc_nmf_b200._lib import Handle
    from gcc_nmf_b200 import gccNMFFunctions as fn
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    h = Handle(local)
    comm = gd.ShardComm()
    F, T2, K = 513, 3744, 1024
    rng = np.random.default_rng(5 + rank)
    V = h.to_device((rng.random((F, T2)) ** 3 + 1e-3).astype(np.float32))
    W0, H0 = fn._seededInit(F, T2, K, 1e-16, 0)
    W0d, H0d = h.to_device(W0), h.to_device(H0)
    numer = h.empty((F * K + K,), torch.float32)
    nt = (T2 + 103) // 104
    gemms = [('G1', nt * 4), ('G2', 18 * 8), ('G3', nt * 4), ('G4', 3 * 8 * 6)]
    # one 8-slot record per launch of the exchange kernels: PACK (row sums [+ numerator] -> symmetric buffer, arrival signal),
    # RB (two-shot: slice reduction + multicast), AP (W update)
    extra = {'fused': ['AP'], 'stepwise-no-exchange': ['PACK', 'AP'], 'two-shot': ['PACK', 'RB', 'AP'], 'one-shot': ['PACK', 'AP'],
             'nccl': ['PACK', 'AP'], 'pull-one-shot': ['AP'], 'pull-two-shot': ['RB', 'AP'], 'pull-in-w-update': ['XW']}
    modes = ['fused', 'stepwise-no-exchange']
    if world > 1:
        modes += ['pull-in-w-update', 'pull-one-shot', 'pull-two-shot', 'two-shot', 'one-shot', 'nccl']
    if os.environ.get('STAMP_MODES'):
        modes = [m for m in modes if m in os.environ['STAMP_MODES'].split(',')]
    iters = 24
    ghz = 1.75        # SM clock under this load (cycles per ns), from clock64 / globaltimer of whole CTAs
    for mode in modes:
        mm = px = None
        if mode.startswith('pull'):
            px = gd.PullExchange.create(h.lib, F, T2, K, h.device, None, two_shot={'pull-one-shot': 0, 'pull-two-shot': 1, 'pull-in-w-update': 2}[mode])
            sup = h.lib.gccnmf_klnmf_pull_supported(h.h, F, T2, K)
            if px is not None:
                px.direct = bool(sup & 2)
                extra[mode] = ['XW'] if px.two_shot == 2 else ((['RB', 'AP'] if px.two_shot else ['AP']) if px.direct else (['PACK', 'RB', 'AP'] if px.two_shot else ['PACK', 'AP']))
            if px is None or sup < 1:
                if rank == 0:
                    print(mode, ': not available on this box (buffer %s, supported %s)' % (px is not None, sup))
                continue
        if mode == 'two-shot':
            mm = gd.MultimemTwoShot.create(F * K + K, h.device, None)
        elif mode == 'one-shot':
            mm = gd.MultimemNumerator.create(F * K + K, h.device, None)
        if mode in ('two-shot', 'one-shot') and mm is None:
            if rank == 0:
                print(mode, ': no multicast support on this box')
            continue

        def run(n):
            W, H = W0d.clone(), H0d.clone()
            if mode == 'fused':
                h.klnmf(V, W, H, n)
            elif mode == 'stepwise-no-exchange':
                h.klnmf_begin(V, W, H)
                for it in range(n):
                    h.klnmf_step_numer(V, W, H, it, numer, 0.0, 1e-16)
                    h.klnmf_step_apply(W, H, numer)
                h.klnmf_end(W, H, n)
            elif mode == 'nccl':
                gd.klnmf_sharded(h, comm, V, W, H, n, 0.0, 1e-16, numer)
            elif px is not None:
                gd.klnmf_sharded_pull(h, px, V, W, H, n, 0.0, 1e-16)
            else:
                gd.klnmf_sharded_multimem(h, mm, V, W, H, n, 0.0, 1e-16)
            return W

        run(8)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(100)
        e1.record()
        e1.synchronize()
        ms100 = e0.elapsed_time(e1)
        buf = torch.zeros(8 << 20, dtype=torch.int64, device=h.device)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        h.lib.gccnmf_debug_timing(h.h, buf.data_ptr(), 1)
        run(iters)
        torch.cuda.synchronize()
        used = h.lib.gccnmf_debug_timing(h.h, None, 1)
        s = buf.cpu().numpy()[:used].reshape(-1, 8)
        starts, ends, recs = {}, {}, {}
        grids = gemms + [(n, 1) for n in extra[mode]]
        per_iter = sum(c for _, c in grids)
        ok = len(s) >= per_iter * iters
        if ok:
            off = 0
            for it in range(iters):
                for name, ctas in grids:
                    k = s[off:off + ctas]
                    off += ctas
                    if ctas == 1 and name in ('PACK', 'RB', 'AP', 'XW'):
                        recs[it, name] = k[0].astype(np.int64)
                        continue
                    k = k[(k[:, 0] > 0) & (k[:, 2] > 0)]
                    # under programmatic dependent launch a CTA is resident long before it may touch memory: its work starts when
                    # its first pipeline stage is full = CTA start + (clock64 at first full stage - clock64 at start) / SM clock
                    work = k[:, 0] + (k[:, 2] - k[:, 1]) / ghz
                    starts[it, name], ends[it, name] = work.min(), k[:, 7].max()
            its = range(4, iters - 1)
            period = np.median([starts[i + 1, 'G1'] - starts[i, 'G1'] for i in its]) / 1e3
            gap = np.median([starts[i + 1, 'G1'] - ends[i, 'G4'] for i in its]) / 1e3
            spans = {n: np.median([ends[i, n] - starts[i, n] for i in its]) / 1e3 for n, _ in gemms}
            inner = {a + '>' + b: np.median([starts[i, b] - ends[i, a] for i in its]) / 1e3 for a, b in (('G1', 'G2'), ('G2', 'G3'), ('G3', 'G4'))}
            msg = 'period %.1f us | G4 end -> next G1 first loads landed %.1f us | spans %s | gaps %s' % (
                period, gap, {k: round(float(v), 1) for k, v in spans.items()}, {k: round(float(v), 1) for k, v in inner.items()})
            # the exchange, relative to the end of the numerator contraction (medians, us)
            med = lambda f: round(float(np.median([f(i) for i in its])) / 1e3, 1)   # noqa: E731
            tl = {}
            if 'PACK' in extra[mode]:
                tl['pack start'] = med(lambda i: recs[i, 'PACK'][0] - ends[i, 'G4'])
                tl['pack signalled'] = med(lambda i: recs[i, 'PACK'][7] - ends[i, 'G4'])
            if 'RB' in extra[mode] and mode == 'pull-two-shot':
                tl['rb start'] = med(lambda i: recs[i, 'RB'][0] - ends[i, 'G4'])
                tl['rb arrivals seen'] = med(lambda i: recs[i, 'RB'][1] - ends[i, 'G4'])
                tl['rb cta0 summed'] = med(lambda i: recs[i, 'RB'][2] - ends[i, 'G4'])
                tl['rb signalled'] = med(lambda i: recs[i, 'RB'][7] - ends[i, 'G4'])
            elif 'RB' in extra[mode]:
                tl['rb start'] = med(lambda i: recs[i, 'RB'][0] - ends[i, 'G4'])
                tl['rb arrivals seen'] = med(lambda i: recs[i, 'RB'][1] - ends[i, 'G4'])
                tl['rb cta0 stored'] = med(lambda i: recs[i, 'RB'][2] - ends[i, 'G4'])
                tl['rb cta0 fenced'] = med(lambda i: recs[i, 'RB'][3] - ends[i, 'G4'])
                tl['rb signalled'] = med(lambda i: recs[i, 'RB'][7] - ends[i, 'G4'])
            if 'XW' in extra[mode]:
                tl['w update start'] = med(lambda i: recs[i, 'XW'][0] - ends[i, 'G4'])
                tl['tile published + flagged'] = med(lambda i: recs[i, 'XW'][1] - ends[i, 'G4'])
                tl['peers tile flagged'] = med(lambda i: recs[i, 'XW'][2] - ends[i, 'G4'])
                tl['operands in'] = med(lambda i: recs[i, 'XW'][3] - ends[i, 'G4'])
                tl['cta0 end'] = med(lambda i: recs[i, 'XW'][7] - ends[i, 'G4'])
                msg += ' | after G4 end: %s' % tl
                print('rank %d %-22s: %.2f ms per 100 iterations | %s' % (rank, mode, ms100, msg), flush=True)
                continue
            tl['apply start'] = med(lambda i: recs[i, 'AP'][0] - ends[i, 'G4'])
            tl['apply arrivals seen'] = med(lambda i: recs[i, 'AP'][1] - ends[i, 'G4'])
            tl['apply operands in'] = med(lambda i: recs[i, 'AP'][2] - ends[i, 'G4'])
            tl['apply cta0 end'] = med(lambda i: recs[i, 'AP'][7] - ends[i, 'G4'])
            msg += ' | after G4 end: %s' % tl
        else:
            msg = 'stamps: %d records, expected %d' % (len(s), per_iter * iters)
        print('rank %d %-22s: %.2f ms per 100 iterations | %s' % (rank, mode, ms100, msg), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
