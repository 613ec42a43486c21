"""Online / low-latency speech enhancement: the frame loops of
notebooks/onlineSpeechEnhancement.ipynb cell 23 (:406-447) and lowLatencySpeechEnhancement.ipynb cell 30
(:511-584), plus the asymmetric windows of cell 21 (:371-392).

The notebooks' loop is sequential only through one piece of state -- the accumulated maximum of the GCC-PHAT
angular spectrum (:416), a prefix maximum over time -- so here ALL frames are processed in one batch:
STFT (numpy.fft.rfft convention: not conjugated) -> PHAT coherence + angular spectrogram -> prefix max +
argmax (target TDOA per frame) -> all-TDOA GCC-NMF argmax per atom -> 0/1 atom mask -> Wiener-like filter ->
inverse FFT -> overlap-add of the raw frames (the notebooks never apply the synthesis window, :575-578).

Precision: the notebook keeps complex128 spectra (numpy.fft.rfft of float64); this path stores complex64 like
the offline path and the reference's real-time class (gccNMFProcessor.py:202), so float outputs agree to
float32 rounding and integer decisions (targets, atom masks) can differ only on float32-level near-ties.
"""
import numpy as np

from . import gccNMFFunctions as fn
from ._lib import default_handle


def getAsymmetricAnalysisWindow(k, m, d):
    """lowLatencySpeechEnhancement.ipynb cell 21 (:371-380)."""
    rising = np.sqrt(np.hanning(2 * (k - m - d) + 1)[:2 * (k - m - d)])
    falling = np.sqrt(np.hanning(2 * m + 1)[:2 * m])
    window = np.zeros(k)
    window[d:k - m] = rising[:k - m - d]
    window[k - m:] = falling[-m:]
    return window


def getAsymmetricSynthesisWindow(k, m, d):
    """lowLatencySpeechEnhancement.ipynb cell 21 (:382-392) (constructed by the notebook, never applied)."""
    risingAnalysis = np.sqrt(np.hanning(2 * (k - m - d) + 1)[:2 * (k - m - d)])
    risingNormalized = np.hanning(2 * m + 1)[:m] / risingAnalysis[k - 2 * m - d:k - m - d]
    falling = np.sqrt(np.hanning(2 * m + 1)[:2 * m])
    window = np.zeros(k)
    window[-2 * m:-m] = risingNormalized
    window[-m:] = falling[-m:]
    return window


def performOnlineSpeechEnhancement(stereoSamples, sampleRate, W, analysisWindow, synthesisWindow, hopSize, numTDOAs,
                                   microphoneSeparationInMetres, targetTDOAEpsilon, numInferenceIterations=0,
                                   gainPerFrame=False, device=0):
    """Returns the notebook's tuple (lowLatencySpeechEnhancement.ipynb:583-584):
    inputSpectrogram, outputSpectrogram, targetEstimateSamplesOLA, gccPHATAccumulatedMax, targetTDOAs,
    angularSpectrogram, atomMasks, wienerFilters.

    gainPerFrame=True is the online notebook (:376, :447: frames counted with the analysis window, gain applied per
    frame); False is the low-latency notebook (:513, :580: frames counted with len(synthesisWindow), gain at the end).
    """
    if numInferenceIterations != 0:
        raise NotImplementedError('numInferenceIterations > 0 calls inferCoefficientsKLNMF, which the reference never defines '
                                  '(onlineSpeechEnhancement.ipynb:433); run gccNMFFunctions.inferCoefficientsKLNMF separately')
    stereoSamples = np.ascontiguousarray(stereoSamples, dtype=np.float32)
    analysisWindow = np.ascontiguousarray(analysisWindow, dtype=np.float64)
    N = len(analysisWindow)
    numSamples = stereoSamples.shape[1]
    W = np.ascontiguousarray(W, dtype=np.float32)
    F, K = W.shape
    synthLen = len(synthesisWindow)
    numFrames = (numSamples - (N if gainPerFrame else synthLen)) // hopSize
    numFrames = min(numFrames, 1 + (numSamples - N) // hopSize)
    gainFactor = hopSize / float(synthLen) * 2
    h = default_handle(device)
    torch = h.torch
    used = (numFrames - 1) * hopSize + N
    x = h.to_device(stereoSamples[:, :used])
    X = h.stft(x, h.to_device(analysisWindow), N, hopSize, conjugate=False)                      # :410 rfft(frame * window)
    frequenciesInHz = fn.getFrequenciesInHz(sampleRate, F)
    E = h.to_device(np.ascontiguousarray(fn.getExpJOmegaTau(frequenciesInHz, fn.getTDOAsInSeconds(microphoneSeparationInMetres, numTDOAs))))
    coh, ang, _ = h.phat_angspec(X, E, want_mean=False)                                           # :414-415
    accMax, targets = h.online_targets(ang)                                                       # :416-417
    Wd = h.to_device(W)
    argmax, refined = h.tdoa_argmax(coh, E, Wd)                                                   # :422-423
    if int(refined.item()) > h.lib.gccnmf_tdoa_argmax_refine_capacity(K, numFrames):
        _, argmax = h.tdoa_gccnmf(coh, E, Wd, want_values=False, want_argmax=True)
    atomMasks = h.atom_mask(argmax, targets=targets, epsilon=targetTDOAEpsilon, mode=0)          # :424-425
    Y, wiener = h.wiener_apply(atomMasks, Wd, X, want_filter=True)                                # :429-431, :440
    ones = np.full(N, gainFactor if gainPerFrame else 1.0)
    y = h.istft_ola(Y, h.to_device(ones), N, hopSize, gain=np.float32(1.0 if gainPerFrame else gainFactor),
                    center=False, conjugate=False)                                                # :443-447 / :575-580
    out = np.zeros_like(stereoSamples)
    out[:, :y.shape[1]] = y.cpu().numpy()
    wf = wiener.cpu().numpy().astype(np.float64)
    return (X.cpu().numpy(), Y.cpu().numpy(), out, accMax[:, -1].cpu().numpy(), targets.cpu().numpy().astype(np.float64),
            ang.cpu().numpy(), atomMasks.cpu().numpy().astype(np.float64), np.stack([wf, wf]))
