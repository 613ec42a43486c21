"""wav in / out with the reference's sample conventions (gccNMF/wavfile.py:34-48, 57-125): integer PCM <-> float in
[-1, 1) through the dtype's half range (unsigned PCM is re-centred), clip protection to 0.99 when a signal reaches 1."""
import logging

import numpy as np

CLIP_PROTECTION_MAX_SAMPLE_VALUE = 0.99      # wavfile.py:32


def _half_range_and_offset(int_dtype):
    info = np.iinfo(int_dtype)
    half = 2 ** (info.bits - 1)
    return info, half, info.min + half       # offset: 0 for signed PCM, 128 for uint8


def pcm2float(sig, dtype='float32'):
    """wavfile.py:57-90."""
    sig = np.asarray(sig)
    if sig.dtype.kind not in 'iu':
        raise TypeError("'sig' must be an array of integers")
    if np.dtype(dtype).kind != 'f':
        raise TypeError("'dtype' must be a floating point type")
    _, half, offset = _half_range_and_offset(sig.dtype)
    return (sig.astype(dtype) - offset) / half


def float2pcm(sig, dtype='int16'):
    """wavfile.py:92-125: scale by the half range (2^15 for int16), clip to the integer range, truncate."""
    sig = np.asarray(sig)
    if sig.dtype.kind != 'f':
        raise TypeError("'sig' must be a float array")
    if np.dtype(dtype).kind not in 'iu':
        raise TypeError("'dtype' must be an integer type")
    info, half, offset = _half_range_and_offset(dtype)
    return (sig * half + offset).clip(info.min, info.max).astype(dtype)


def wavread(filePath):
    """wavfile.py:34-37 -> (channels, n) float32, sample rate."""
    from scipy.io import wavfile
    sampleRate, pcm = wavfile.read(filePath)
    return pcm2float(pcm).T, sampleRate


def wavwrite(samples_float32, filePath, sampleRate, clipProtection=True):
    """wavfile.py:39-48: samples (channels, n); a peak >= 1 is rescaled to 0.99 (or raises without clip protection)."""
    from scipy.io import wavfile
    samples_float32 = np.asarray(samples_float32)
    peak = np.max(np.abs(samples_float32))
    if peak >= 1:
        if not clipProtection:
            raise ValueError('wavwrite: max abs signal value exceeds 1')
        logging.warning('wavwrite: max abs signal value exceeds 1, rescaling to %2f' % CLIP_PROTECTION_MAX_SAMPLE_VALUE)
        samples_float32 = samples_float32 / peak * CLIP_PROTECTION_MAX_SAMPLE_VALUE
    wavfile.write(filePath, sampleRate, float2pcm(samples_float32.astype(np.float32)).T)
