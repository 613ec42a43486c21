"""Diagnostics: KL-NMF accuracy and speed of the two operand-split modes against the CPU oracle at the headline shape."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, '.')
from gcc_nmf_b200._lib import default_handle
from gcc_nmf_b200.synth import synthetic_stereo
import gcc_nmf_b200.gccNMFFunctions as fn
from oracle import gccnmf_oracle as orc
h = default_handle()
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
x = h.to_device(synthetic_stereo(30.0))
X, V = h.stft(x, h.to_device(np.hanning(1024)), 1024, 256, conjugate=True, want_V=True)
Vh = V.cpu().numpy()
W0, H0 = fn._seededInit(513, 3744, 1024, 1e-16, 0)
t0 = time.time(); Wo, Ho = orc.performKLNMF(Vh, 1024, iters, 0, W0=W0, H0=H0); print('oracle %d iterations: %.1f s' % (iters, time.time() - t0))
rel = lambda a, b: float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b.astype(np.float64)))
for name, opt in (('3xTF32', 0), ('3xBF16', 1)):
    h.set_option('nmf_split_bf16', opt)
    for rep in range(2):
        W, H = h.to_device(W0.copy()), h.to_device(H0.copy())
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); h.klnmf(V, W, H, iters); e1.record(); torch.cuda.synchronize()
    Wg, Hg = W.cpu().numpy(), H.cpu().numpy()
    print('%s: %.2f ms for %d iterations | rel W %.3e  rel H %.3e | max-rel W %.3e' % (
        name, e0.elapsed_time(e1), iters, rel(Wg, Wo), rel(Hg, Ho), float(np.max(np.abs(Wg - Wo)) / np.max(Wo))))
