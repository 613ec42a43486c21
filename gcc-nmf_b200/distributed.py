"""Frame-sharded GCC-NMF across the GPUs of one box.

One long recording is cut into contiguous frame ranges, one per rank (one process per GPU,
torch.distributed / NCCL over NVLink).  Every stage of the path is independent per frame except:

  1. the KL-NMF W update sums over ALL frames (gccNMFFunctions.py:77): each rank computes its partial
     numerator (V_s/(W H_s)).H_s^T and rowsum(H_s); ONE all-reduce of F*K + K floats per iteration
     makes W identical everywhere; normalisation (:79-81) is then computed redundantly per rank;
  2. the mean angular spectrum used for peak picking (runGCCNMF.py:46): one D-float all-reduce;
  3. the iSTFT overlap-add needs N - hop samples from the neighbouring rank at each seam.

Column order matches the single-GPU run on the whole recording: the reference's V is
[left frames | right frames] (runGCCNMF.py:40); rank r holds both channels of its own frames, and
takes the matching column slices of the globally seeded H0, so 1-GPU and n-GPU runs agree up to
float32 reduction order.

`ShardComm` abstracts the three exchanges so the host logic is testable on CPU with gloo.
"""
import os

import numpy as np

from . import gccNMFFunctions as fn


def shard_frames(total_frames, world, rank):
    """Contiguous, balanced frame range [t0, t1) of `rank` (first `total % world` ranks get one more)."""
    base, extra = divmod(total_frames, world)
    t0 = rank * base + min(rank, extra)
    return t0, t0 + base + (1 if rank < extra else 0)


def shard_sample_range(t0, t1, windowSize, hopSize):
    """Samples a rank needs to form frames [t0, t1): [t0 hop, (t1 - 1) hop + N)."""
    return t0 * hopSize, (t1 - 1) * hopSize + windowSize


def sharded_nmf_init(numFrequencies, totalFrames, dictionarySize, epsilon, seedValue, t0, t1):
    """Global seeded draw (gccNMFFunctions.py:70-73 for V of shape (F, 2 T_total)), then this rank's
    column slices [L frames t0:t1 | R frames t0:t1]."""
    W0, H0 = fn._seededInit(numFrequencies, 2 * totalFrames, dictionarySize, epsilon, seedValue)
    H0s = np.ascontiguousarray(np.concatenate([H0[:, t0:t1], H0[:, totalFrames + t0:totalFrames + t1]], axis=1))
    return W0, H0s


class ShardComm(object):
    """The exchanges of the sharded path on torch.distributed (NCCL on GPUs, gloo in CPU tests)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def all_reduce_sum(self, tensor):
        if self.world > 1:
            self.dist.all_reduce(tensor, op=self.dist.ReduceOp.SUM, group=self.group)
        return tensor

    def seam_exchange(self, tail, head_like):
        """Send `tail` (my samples past my last owned sample) to rank+1, receive rank-1's tail
        (zeros for rank 0).  Returns the received halo, shaped like `head_like`."""
        import torch
        recv = torch.zeros_like(head_like)
        if self.world == 1:
            return recv
        ops = []
        if self.rank + 1 < self.world:
            ops.append(self.dist.P2POp(self.dist.isend, tail.contiguous(), self.rank + 1, self.group))
        if self.rank > 0:
            ops.append(self.dist.P2POp(self.dist.irecv, recv, self.rank - 1, self.group))
        for req in self.dist.batch_isend_irecv(ops):
            req.wait()
        return recv


def klnmf_sharded(ops, comm, V_s, W, H_s, numIterations, sparsityAlpha, epsilon, numer):
    """gccNMFFunctions.py:75-81 on frame shards.  `ops` provides the C-ABI step protocol
    (klnmf_begin / klnmf_step_numer / klnmf_step_apply / klnmf_end, include/gccnmf_b200.h);
    `numer` is an (F*K + K) float32 buffer: the single all-reduce per iteration."""
    ops.klnmf_begin(V_s, W, H_s)
    for it in range(numIterations):
        ops.klnmf_step_numer(V_s, W, H_s, it, numer, sparsityAlpha, epsilon)
        comm.all_reduce_sum(numer)
        ops.klnmf_step_apply(W, H_s, numer)
    ops.klnmf_end(W, H_s, numIterations)
    return W, H_s


class MultimemNumerator(object):
    """Two symmetric buffers of (F*K + K) floats + one uint32 arrival counter each, bound to an NVLink multicast object (torch
    symmetric memory), for the W update of a frame-sharded run with the exchange fused into the kernels
    (gccnmf_klnmf_step_multimem): the numerator pack writes the local partial and adds 1 to the counter of EVERY rank through
    the multicast address, the W-update kernel waits until its own copy of the counter shows one arrival per rank and reads the
    cross-rank SUM with multimem.ld_reduce (reduction inside the NVSwitch).  No host-launched barrier, no all-reduce.
    Double-buffered by iteration parity.  `create` returns None when the platform has no multicast support (the caller then
    uses the NCCL all-reduce)."""
    COUNTER_PAD = 64        # floats after the numerator: the uint32 arrival counter lives in the first of them

    def __init__(self, buffers, handles, numel, world):
        self.buffers, self.handles, self.numel, self.world = buffers, handles, numel, world
        self.uses = [0, 0]              # how often each buffer has been used (identical on every rank)

    @classmethod
    def create(cls, numel, device, group):
        try:
            import torch
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm_mem
            group = group if group is not None else dist.group.WORLD
            try:
                symm_mem.enable_symm_mem_for_group(group.group_name)
            except Exception:
                pass
            buffers, handles = [], []
            for _ in range(2):
                t = symm_mem.empty(numel + cls.COUNTER_PAD, dtype=torch.float32, device=device)
                hdl = symm_mem.rendezvous(t, group)
                if not getattr(hdl, 'has_multicast_support', False) or not int(hdl.multicast_ptr):
                    return None
                t.zero_()
                buffers.append(t)
                handles.append(hdl)
            torch.cuda.synchronize(device)
            for hdl in handles:
                hdl.barrier(channel=0, timeout_ms=20000)       # every rank's counters are zero before anyone signals
            return cls(buffers, handles, numel, dist.get_world_size(group))
        except Exception:
            return None

    def step_args(self, it):
        """(numer_local, numer_multicast, counter_local, counter_multicast, arrivals_expected) for iteration `it`."""
        p = it & 1
        self.uses[p] += 1
        local, mc = int(self.buffers[p].data_ptr()), int(self.handles[p].multicast_ptr)
        off = 4 * self.numel
        return local, mc, local + off, mc + off, self.world * self.uses[p]


class MultimemTwoShot(object):
    """One symmetric allocation [numerator (F*K + K) | reduced (F*K + K) | 2 uint32 arrival counters], bound to an NVLink multicast
    object, for gccnmf_klnmf_step_multimem2: pack -> every rank sums its 1 / world slice inside the NVSwitch (multimem.ld_reduce) and
    multicasts it to all ranks (multimem.st) -> W update from the local `reduced` copy.  Per GPU and iteration the links carry one
    numerator each way whatever the world size.  `create` returns None without multicast support."""
    COUNTER_PAD = 64

    def __init__(self, buffer, handle, numel, rank, world):
        self.buffer, self.handle, self.numel, self.rank, self.world = buffer, handle, numel, rank, world
        self.iterations = 0             # arrivals so far / world (identical on every rank)

    @classmethod
    def create(cls, numel, device, group):
        try:
            import torch
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm_mem
            group = group if group is not None else dist.group.WORLD
            try:
                symm_mem.enable_symm_mem_for_group(group.group_name)
            except Exception:
                pass
            t = symm_mem.empty(2 * numel + cls.COUNTER_PAD, dtype=torch.float32, device=device)
            hdl = symm_mem.rendezvous(t, group)
            if not getattr(hdl, 'has_multicast_support', False) or not int(hdl.multicast_ptr):
                return None
            t.zero_()
            torch.cuda.synchronize(device)
            hdl.barrier(channel=0, timeout_ms=20000)           # every rank's counters are zero before anyone signals
            return cls(t, hdl, numel, dist.get_rank(group), dist.get_world_size(group))
        except Exception:
            return None

    def step_args(self, it):
        """(rank, world, numer_local, numer_mc, reduced_local, reduced_mc, counters_local, counters_mc, arrivals_expected)."""
        self.iterations += 1
        local, mc = int(self.buffer.data_ptr()), int(self.handle.multicast_ptr)
        red, cnt = 4 * self.numel, 8 * self.numel
        return (self.rank, self.world, local, mc, local + red, mc + red, local + cnt, mc + cnt, self.world * self.iterations)


class PullExchange(object):
    """Symmetric buffer (torch symmetric memory: every rank's allocation mapped into every process) for gccnmf_klnmf_step_pull: the
    numerator contraction writes this rank's partial into it and signals every rank; the W updates read the partials (or, two-shot,
    the owners' slice sums) with plain peer loads.  No multicast object, no NCCL call, nothing on the host inside the loop.
    form 0 one-shot (world - 1 numerators inbound per GPU), 1 two-shot (one numerator in each direction for any world size), 2 the
    exchange inside the W update, tile by tile (no pack kernel, no kernel boundary inside the exchange; world - 1 inbound)."""

    def __init__(self, buffer, handle, bases, rank, world, layout_T2, two_shot):
        self.buffer, self.handle, self.bases, self.rank, self.world = buffer, handle, bases, rank, world
        self.layout_T2, self.two_shot = layout_T2, two_shot
        self.epoch = 0                  # iterations executed on this buffer so far (identical on every rank)
        self.direct = False             # agreed by the caller: every rank supports the direct form (gccnmf_klnmf_pull_supported == 2)

    @classmethod
    def create(cls, lib, F, layout_T2, K, device, group, two_shot=None):
        try:
            import ctypes
            import torch
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm_mem
            group = group if group is not None else dist.group.WORLD
            try:
                symm_mem.enable_symm_mem_for_group(group.group_name)
            except Exception:
                pass
            world, rank = dist.get_world_size(group), dist.get_rank(group)
            if world > 8:
                return None
            numel = int(lib.gccnmf_klnmf_pull_buffer_floats(F, layout_T2, K))
            t = symm_mem.empty(numel, dtype=torch.float32, device=device)
            hdl = symm_mem.rendezvous(t, group)
            ptrs = [int(p) for p in hdl.buffer_ptrs]
            if len(ptrs) != world or not all(ptrs):
                return None
            t.zero_()
            torch.cuda.synchronize(device)
            hdl.barrier(channel=0, timeout_ms=20000)           # every rank's buffer is zero before anyone signals
            bases = (ctypes.c_void_p * world)(*ptrs)
            if two_shot is None:
                two_shot = 1 if world >= 4 else 0
            return cls(t, hdl, bases, rank, world, layout_T2, int(two_shot))
        except Exception as e:                                     # noqa: BLE001  (no symmetric memory / no peer mapping: the caller falls back)
            if os.environ.get('GCCNMF_DEBUG_EXCHANGE'):
                import traceback
                traceback.print_exc()
                print('PullExchange.create failed:', repr(e), flush=True)
            return None


def klnmf_sharded_pull(ops, px, V_s, W, H_s, numIterations, sparsityAlpha, epsilon):
    """klnmf_sharded with the pull exchange (gccnmf_klnmf_step_pull): one C call per iteration, five or six kernels, no host step."""
    ops.klnmf_begin(V_s, W, H_s)
    for it in range(numIterations):
        ops.klnmf_step_pull(V_s, W, H_s, it, px.epoch, px.rank, px.world, px.bases, px.layout_T2, px.two_shot, px.direct, sparsity_alpha=sparsityAlpha,
                            epsilon=epsilon)
    px.epoch += numIterations
    ops.klnmf_end(W, H_s, numIterations)
    return W, H_s


def klnmf_sharded_multimem(ops, mm, V_s, W, H_s, numIterations, sparsityAlpha, epsilon):
    """klnmf_sharded with the exchange fused into the kernels: one C call per iteration, nothing on the host between the numerator
    and the W update (gccnmf_klnmf_step_multimem: one-shot; gccnmf_klnmf_step_multimem2: two-shot)."""
    step = ops.klnmf_step_multimem2 if isinstance(mm, MultimemTwoShot) else ops.klnmf_step_multimem
    ops.klnmf_begin(V_s, W, H_s)
    for it in range(numIterations):
        step(V_s, W, H_s, it, *mm.step_args(it), sparsity_alpha=sparsityAlpha, epsilon=epsilon)
    ops.klnmf_end(W, H_s, numIterations)
    return W, H_s


def overlap_add_seams(comm, y_local, owned, halo):
    """y_local: (B, owned + halo) un-trimmed local overlap-add.  Adds the left neighbour's tail onto
    this rank's head and returns the owned part (the last rank keeps its tail)."""
    import torch
    tail = y_local[:, owned:owned + halo]
    recv = comm.seam_exchange(tail, y_local[:, :halo])
    y_local[:, :halo] += recv
    if comm.rank + 1 < comm.world:
        return y_local[:, :owned]
    return y_local


class ShardedGCCNMFPipeline(object):
    """Enhancement flow (offlineSpeechEnhancement.ipynb cells 12-41) on one recording of
    world x clip_seconds, frame-sharded; the dictionary is learnt jointly."""

    def __init__(self, sampleRate, windowSize, hopSize, numTDOAs, microphoneSeparationInMetres, dictionarySize,
                 numIterations, sparsityAlpha=0.0, epsilon=1e-16, seedValue=0, targetTDOAWindowSizePercent=0.05,
                 device=0, clip_seconds=30.0, comm=None, handle=None):
        import torch
        from ._lib import Handle
        self.torch = torch
        self.comm = comm if comm is not None else ShardComm()
        self.h = handle if handle is not None else Handle(device)
        from .pipeline import _BufferOwner
        import weakref
        self._token = _BufferOwner()
        weakref.finalize(self, self.h.release, self._token)
        self.sr, self.N, self.hop, self.D = sampleRate, int(windowSize), int(hopSize), int(numTDOAs)
        self.micSep, self.K, self.I = microphoneSeparationInMetres, int(dictionarySize), int(numIterations)
        self.alpha, self.eps, self.seed = float(sparsityAlpha), float(epsilon), seedValue
        self.windowPercent = targetTDOAWindowSizePercent
        self.F = self.N // 2 + 1
        world, rank = self.comm.world, self.comm.rank
        self.total_samples = int(round(clip_seconds * sampleRate)) * world
        self.total_frames = 1 + (self.total_samples - self.N) // self.hop
        self.t0, self.t1 = shard_frames(self.total_frames, world, rank)
        self.s0, self.s1 = shard_sample_range(self.t0, self.t1, self.N, self.hop)
        self.clip_seconds = clip_seconds
        self.frequenciesInHz = fn.getFrequenciesInHz(sampleRate, self.F)
        self.hypothesisTDOAs = fn.getTDOAsInSeconds(microphoneSeparationInMetres, self.D)
        self.E_host = np.ascontiguousarray(fn.getExpJOmegaTau(self.frequenciesInHz, self.hypothesisTDOAs))
        self.window = self.h.to_device(np.hanning(self.N))
        self.E = self.h.to_device(self.E_host)
        W0, H0s = sharded_nmf_init(self.F, self.total_frames, self.K, self.eps, self.seed, self.t0, self.t1)
        self.W0, self.H0s = self.h.to_device(W0), self.h.to_device(H0s)
        self.numer = self.h.empty((self.F * self.K + self.K,), torch.float32)
        self.multimem = None            # decided on the first call (needs the agreed NMF path)
        self.pull = None
        self.collective = 'nccl-all-reduce'
        self.stage_events = None

    def local_samples(self):
        """This rank's slice of the synthetic recording (each 30 s clip seeded by its index)."""
        from .synth import synthetic_stereo
        n_clip = int(round(self.clip_seconds * self.sr))
        first, last = self.s0 // n_clip, (self.s1 - 1) // n_clip
        x = np.concatenate([synthetic_stereo(self.clip_seconds, self.sr, seed=1234 + c) for c in range(first, last + 1)], axis=1)
        return np.ascontiguousarray(x[:, self.s0 - first * n_clip:self.s1 - first * n_clip])

    def _mark(self, name):
        if self.stage_events is not None:
            ev = self.torch.cuda.Event(enable_timing=True)
            ev.record()
            self.stage_events.append((name, ev))

    def stage_times_ms(self):
        ev = self.stage_events
        return {ev[i + 1][0]: ev[i][1].elapsed_time(ev[i + 1][1]) for i in range(len(ev) - 1)}

    def _all_min(self, value):
        t = self.torch.tensor([int(value)], dtype=self.torch.int32, device=self.h.device)
        if self.comm.world > 1:
            self.comm.dist.all_reduce(t, op=self.comm.dist.ReduceOp.MIN, group=self.comm.group)
        return int(t.item())

    def _all_and(self, value, bits=8):
        """Bitwise AND over the ranks (NCCL has no BAND: MIN over the separated bits)."""
        t = self.torch.tensor([(int(value) >> b) & 1 for b in range(bits)], dtype=self.torch.int32, device=self.h.device)
        if self.comm.world > 1:
            self.comm.dist.all_reduce(t, op=self.comm.dist.ReduceOp.MIN, group=self.comm.group)
        return sum(int(v) << b for b, v in enumerate(t.tolist()))

    def _all_max(self, value):
        t = self.torch.tensor([int(value)], dtype=self.torch.int32, device=self.h.device)
        if self.comm.world > 1:
            self.comm.dist.all_reduce(t, op=self.comm.dist.ReduceOp.MAX, group=self.comm.group)
        return int(t.item())

    def _agree_on_nmf_path(self, T2):
        """All ranks must run the W update with the same kernels (bit-identical W): if any rank's shard
        shape falls back to the SIMT contractions, every rank does.  Decided once."""
        if getattr(self, '_path_agreed', False):
            return
        flag = self.torch.tensor([1 if self.h.klnmf_uses_tensor_cores(self.F, T2, self.K) else 0],
                                 dtype=self.torch.int32, device=self.h.device)
        if self.comm.world > 1:
            self.comm.dist.all_reduce(flag, op=self.comm.dist.ReduceOp.MIN, group=self.comm.group)
        if int(flag.item()) == 0:
            self.h.set_option('force_simt_nmf', 1)
        elif self.comm.world > 1 and os.environ.get('GCCNMF_COLLECTIVE', 'auto') != 'nccl':
            # exchange fused into the kernels (all ranks must agree that it is available).  GCCNMF_COLLECTIVE: auto (default: measured at
            # 2 ranks, DESIGN.md section 5 -- one-shot pull for 2 ranks, whose traffic grows with the world size, two-shot inside the
            # switch from 3 ranks up), pull1 / pull2 (one- / two-shot pull), multimem / multimem1 (two- / one-shot inside the switch), nccl
            mode = os.environ.get('GCCNMF_COLLECTIVE', 'auto')
            if mode == 'auto':
                mode = 'pull1' if self.comm.world == 2 else 'multimem' 
            agree = lambda ok: int(self._all_min(1 if ok else 0)) == 1       # noqa: E731
            if mode.startswith('pull'):
                layout_T2 = int(self._all_max(T2))
                px = PullExchange.create(self.h.lib, self.F, layout_T2, self.K, self.h.device, self.comm.group,
                                         two_shot={'pull1': 0, 'pull2': 1, 'pullw': 2}.get(mode))     # ('pull': by world size)
                level = self.h.lib.gccnmf_klnmf_pull_supported(self.h.h, self.F, T2, self.K) if px is not None else 0
                level = self._all_and(level)                                  # bits: 1 pull available, 2 direct contraction, 4 form 2
                if px is not None and px.two_shot == 2 and not level & 4:
                    px.two_shot = 0
                if level & 1:
                    px.direct = bool(level & 2)
                    self.pull = px
                    self.collective = ('pull exchange, %s: %s; %s; no system-scope fence, no multimem, no NCCL call in the loop' % (
                                           {0: 'one-shot', 1: 'two-shot', 2: 'inside the W update'}[px.two_shot],
                                           'each W-update CTA sums this rank\'s k-split slabs for its tile, publishes it, flags every rank and reads '
                                           'the same tile of the other ranks' if px.two_shot == 2 else
                                           'the numerator contraction writes its partial into the symmetric buffer and signals every rank from its '
                                           'last CTA' if px.direct else 'the pack kernel sums the k-split slabs into the symmetric buffer and signals '
                                           'every rank',
                                           'five launches per iteration' if px.two_shot == 2 else
                                           'each rank sums its 1/world slice with plain peer loads, the W updates fetch every word from its owner'
                                           if px.two_shot == 1 else 'the W update reads every rank\'s partial with plain peer loads, added in rank order'))
                else:
                    mode = 'multimem'
            if self.pull is None and mode.startswith('multimem'):
                one_shot = mode == 'multimem1'
                mm = (MultimemNumerator if one_shot else MultimemTwoShot).create(self.F * self.K + self.K, self.h.device, self.comm.group)
                if agree(mm is not None):
                    self.multimem = mm
                    self.collective = ('one-shot in the switch: multimem.red arrival signal + multimem.ld_reduce of the whole sum in the W-update '
                                       'kernel') if one_shot else (
                                       'two-shot in the switch: each rank multimem.ld_reduce-s its 1/world slice and multimem.st-s it to every '
                                       'rank, W update from local memory')
        self._path_agreed = True

    def enhance(self, samples, collect_stage_times=False):
        h, torch, comm = self.h, self.torch, self.comm
        self.stage_events = [] if collect_stage_times else None
        self._mark('start')
        key = self._token
        X, V = h.stft(samples, self.window, self.N, self.hop, conjugate=True, want_V=True, out_key=key)
        Ts = X.shape[2]
        assert Ts == self.t1 - self.t0
        self._mark('stft')
        coh, ang, mean = h.phat_angspec(X, self.E, out_key=key)
        total = mean * float(Ts)                      # back to the local sum over frames
        comm.all_reduce_sum(total)
        mean_host = (total / float(self.total_frames)).cpu().numpy()
        self._mark('angular')
        W, H = h.buffer((key, 'W'), self.W0.shape, self.W0.dtype), h.buffer((key, 'H'), self.H0s.shape, self.H0s.dtype)
        W.copy_(self.W0)
        H.copy_(self.H0s)
        self._agree_on_nmf_path(V.shape[1])
        if self.pull is not None:
            klnmf_sharded_pull(h, self.pull, V, W, H, self.I, self.alpha, self.eps)
        elif self.multimem is not None:
            klnmf_sharded_multimem(h, self.multimem, V, W, H, self.I, self.alpha, self.eps)
        else:
            klnmf_sharded(h, comm, V, W, H, self.I, self.alpha, self.eps, self.numer)
        self._mark('nmf')
        argmax, refined = h.tdoa_argmax(coh, self.E, W, out_key=key)
        self._mark('gccnmf')
        if int(refined.item()) > h.lib.gccnmf_tdoa_argmax_refine_capacity(self.K, Ts):
            _, argmax = h.tdoa_gccnmf(coh, self.E, W, want_values=False, want_argmax=True)   # exact float64 kernel
        target = int(fn.estimateTargetTDOAIndexesFromAngularSpectrum(mean_host, self.micSep, self.D, 1)[0])
        window = (self.hypothesisTDOAs[-1] - self.hypothesisTDOAs[0]) * self.windowPercent
        lut = fn.getTargetTDOALookup(self.hypothesisTDOAs, target, window)
        mask = h.argmax_mask(argmax, h.to_device(lut.astype(np.uint8)), out_key=key)
        self._mark('mask')
        est = h.masked_recon_phase(mask[None], X, W, H, out_key=key)
        self._mark('recon')
        y = h.istft_ola(est.reshape(2, self.F, Ts), self.window, self.N, self.hop,
                        gain=np.float32(self.hop / float(self.N) * 2), center=False, conjugate=True, out_key=key)
        y = overlap_add_seams(comm, y, self.hop * Ts, self.N - self.hop)
        # global centre trim (librosaSTFT.py:283-284): N/2 samples off each end of the whole recording
        if comm.rank == 0:
            y = y[:, self.N // 2:]
        if comm.rank == comm.world - 1:
            y = y[:, :y.shape[1] - self.N // 2]
        self._mark('istft')
        return dict(X=X, V=V, W=W, H=H, coherence=coh, angularSpectrogram=ang, targetTDOAIndexes=[target],
                    argMaxGCCNMF=argmax, targetCoefficientMasks=mask[None], targetSpectrogramEstimates=est,
                    targetSignalEstimates=y.contiguous()[None])

    def enhance_host(self, samples_host, out_host=None):
        torch = self.torch
        r = self.enhance(samples_host.to(self.h.device, non_blocking=True))
        y = r['targetSignalEstimates']
        if out_host is None:
            out_host = torch.empty(y.shape, dtype=torch.float32, pin_memory=True)
        out_host.copy_(y, non_blocking=True)
        torch.cuda.current_stream(self.h.device).synchronize()
        return out_host
