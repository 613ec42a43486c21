"""Diagnostics: per-CTA phase breakdown of the tcgen05 GEMM (clock64 stamps) at the KL-NMF shapes."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from gcc_nmf_b200._lib import default_handle, _ptr
h = default_handle()
import os
h.set_option("nmf_split_bf16", 1 if os.environ.get("SPLIT") == "bf16" else 0)
shapes = {'G1/G3 (F x T2 over K)': (512, 3744, 1024, 128), 'G2 (K x T2 over F)': (1024, 3744, 513, 256), 'G4-like (F x K over T2/4)': (512, 1024, 936, 128)}
for name, (M, N, Kc, bn) in shapes.items():
    ld = (Kc + 3) // 4 * 4
    A = torch.rand(M, ld, device=h.device); B = torch.rand(N, ld, device=h.device)
    if ld != Kc: A[:, Kc:] = 0; B[:, Kc:] = 0
    D = torch.empty(M, N, device=h.device)
    ctas = ((N + bn - 1) // bn) * ((M + 127) // 128)
    T = torch.zeros(ctas * 6, dtype=torch.int64, device=h.device)
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h.check(h.lib.gccnmf_gemm_tn_3xtf32_timed(h.h, _ptr(A), ld, _ptr(B), ld, _ptr(D), N, M, N, Kc, bn, _ptr(T), h.stream))
        e1.record(); torch.cuda.synchronize()
    t = T.cpu().numpy().reshape(ctas, 6).astype(np.float64)
    kb = (Kc + 31) // 32
    d = lambda a, b: np.median(t[:, b] - t[:, a])
    print('%-28s %6.1f us | cycles: start->first full %6.0f | first full->last MMA issued %7.0f (%.0f/k-block, floor %d) | loaders done->acc complete %6.0f | epilogue %6.0f | total %7.0f'
          % (name, e0.elapsed_time(e1) * 1e3, d(0, 1), d(1, 2), d(1, 2) / kb, 768 * bn // 128, d(3, 4), d(4, 5), d(0, 5)))
