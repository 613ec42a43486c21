"""Seeded synthetic stereo generator for the benchmark configurations.

Pure numpy/scipy host code: two amplitude-modulated low-passed noise sources with integer
inter-microphone delays plus sensor noise.  Used by bench.py, the tests and the golden-vector
script; nothing here touches the GPU.
"""
import numpy as np
from scipy.signal import lfilter

SAMPLE_RATE = 16000


def synthetic_stereo(duration_s=30.0, sample_rate=SAMPLE_RATE, seed=1234, num_sources=2):
    """Return (2, n) float32 stereo samples in [-1, 1 - 2**-15].

    source s: N(0,1) noise -> one-pole low-pass y[n] = x[n] + 0.9 y[n-1]
              -> amplitude modulation 0.5 (1 + sin(2 pi (3 + s) t)).
    mic delays (left, right) in samples: source 0 -> (0, 3), source 1 -> (2, 0);
    source gains 1.0 / 0.7; + 0.01 N(0,1) sensor noise per channel; scaled to std 0.1.
    """
    n = int(round(duration_s * sample_rate))
    rng = np.random.default_rng(seed)
    t = np.arange(n) / float(sample_rate)
    delays = [(0, 3), (2, 0), (1, 1), (3, 0)]
    gains = [1.0, 0.7, 0.5, 0.4]
    stereo = np.zeros((2, n))
    for s in range(num_sources):
        x = lfilter([1.0], [1.0, -0.9], rng.standard_normal(n))
        x *= 0.5 * (1.0 + np.sin(2 * np.pi * (3 + s) * t))
        for ch in range(2):
            d = delays[s][ch]
            stereo[ch, d:] += gains[s] * x[:n - d]
    stereo += 0.01 * rng.standard_normal((2, n))
    stereo *= 0.1 / stereo.std()
    np.clip(stereo, -1.0, 1.0 - 2.0 ** -15, out=stereo)
    return stereo.astype(np.float32)
