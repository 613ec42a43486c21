"""Drop-in for gccNMF/librosaSTFT.py (`stft` :20-181, `istft` :183-286) running on the B200.

Same names, argument meaning and error behaviour (ParameterError) as the reference's vendored
librosa functions, restricted to what the GCC-NMF path uses: a callable or vector window of
length n_fft == win_length.  `center=True` in stft (reflect padding) is host-side padding followed
by the same kernel.  numpy in, numpy out; the work is done by gccnmf_stft / gccnmf_istft_ola.
"""
import numpy as np

from ._lib import ParameterError, default_handle

__all__ = ['stft', 'istft', 'ParameterError']


def _window_vector(window, n_fft, scale=1.0):
    if window is None:
        import scipy.signal
        w = scipy.signal.get_window('hann', n_fft, fftbins=True) * scale   # librosaSTFT.py:133-135 / :252-255
    elif callable(window):
        w = window(n_fft)
    else:
        w = np.asarray(window)
        if w.size != n_fft:
            raise ParameterError('Size mismatch between n_fft and len(window)')
    return np.ascontiguousarray(w, dtype=np.float64)


def stft(y, n_fft=2048, hop_length=None, win_length=None, window=None, center=True, dtype=np.complex64, device=None):
    """gccNMF/librosaSTFT.py:20-181 -> (1 + n_fft/2, T) complex64, conjugated like :179."""
    if win_length is None:
        win_length = n_fft
    if win_length != n_fft:
        raise ParameterError('win_length != n_fft is not used on the GCC-NMF path and not supported')
    if hop_length is None:
        hop_length = int(win_length / 4)
    if hop_length < 1:
        raise ParameterError('Invalid hop_length: {:d}'.format(hop_length))
    y = np.asarray(y)
    if y.ndim != 1:
        raise ParameterError('Invalid shape for monophonic audio: ndim={:d}, shape={}'.format(y.ndim, y.shape))
    if not np.isfinite(y).all():
        raise ParameterError('Audio buffer is not finite everywhere')          # librosaSTFT.py:486-487
    w = _window_vector(window, n_fft)
    if center:
        y = np.pad(y, int(n_fft // 2), mode='reflect')
    if len(y) < n_fft:
        raise ParameterError('Buffer is too short (n={:d}) for frame_length={:d}'.format(len(y), n_fft))
    h = default_handle(device)
    X = h.stft(h.to_device(y.astype(np.float32)[None, :]), h.to_device(w), n_fft, hop_length, conjugate=True)
    return X[0].cpu().numpy().astype(dtype, copy=False)


def istft(stft_matrix, hop_length=None, win_length=None, window=None, center=True, dtype=np.float32, device=None):
    """gccNMF/librosaSTFT.py:183-286 -> float32 signal (centre-trimmed by default like :283-284)."""
    stft_matrix = np.asarray(stft_matrix)
    n_fft = 2 * (stft_matrix.shape[0] - 1)
    if win_length is None:
        win_length = n_fft
    if win_length != n_fft:
        raise ParameterError('Size mismatch between n_fft and window size')
    if hop_length is None:
        hop_length = int(win_length / 4)
    w = _window_vector(window, n_fft, scale=2.0 / 3)
    h = default_handle(device)
    spec = h.to_device(np.ascontiguousarray(stft_matrix, dtype=np.complex64)[None])
    y = h.istft_ola(spec, h.to_device(w), n_fft, hop_length, gain=1.0, center=center, conjugate=True)
    return y[0].cpu().numpy().astype(dtype, copy=False)
