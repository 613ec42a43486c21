"""CPU oracle for the GCC-NMF separation hot path.

TEST INFRASTRUCTURE ONLY.  This module is a numpy restatement of the reference algorithm
(seanwood/gcc-nmf @ 0b13e9e).  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
CPU-baseline / `--impl reference` legs may import it, and only as the checker or the timed
CPU baseline -- never on the product path (the product fails loudly without its CUDA library).

Every function cites the reference lines it restates (paths relative to the reference root).
The restatement keeps the reference's dtypes and operation order (float64 STFT rounded to
complex64, float32 KL-NMF, complex128 angular spectrogram, float32 inverse FFT ...) so that it
is usable both as the parity oracle and as an honest CPU timing baseline.

Parity status: PINNED against outputs of the unmodified reference executed in the build
container (`oracle/make_golden.py` -> `tests/golden/*.npz`, checked by
`tests/test_oracle_golden.py`) for every stage that exists as reference library code or as a
runnable notebook cell.  Two pieces are UNPINNED because the reference has nothing to run:
`inferCoefficientsKLNMF` (called at notebooks/onlineSpeechEnhancement.ipynb:433 but defined
nowhere) and `GCCNMFProcessor.processFrames` (needs Theano, not installed, no network).
Third-party arithmetic (numpy 2.3.5, scipy 1.18.1 pocketfft) is not vendored by the reference
and no versions are pinned there; the versions above are the ones the fixtures were made with.
"""
import numpy as np
import scipy.fftpack as fftpack
from scipy.signal import argrelmax

SPEED_OF_SOUND_IN_METRES_PER_SECOND = 340.29  # gccNMF/gccNMFFunctions.py:38, gccNMF/defs.py


# ----------------------------------------------------------------------------- helpers
def getTDOAsInSeconds(microphoneSeparationInMetres, numTDOAs):
    """gccNMF/gccNMFFunctions.py:50-56."""
    maxTDOA = microphoneSeparationInMetres / SPEED_OF_SOUND_IN_METRES_PER_SECOND
    return np.linspace(-maxTDOA, maxTDOA, numTDOAs)


def getFrequenciesInHz(sampleRate, numFrequencies):
    """gccNMF/gccNMFFunctions.py:58-59."""
    return np.linspace(0, sampleRate / 2, numFrequencies)


def getExpJOmegaTau(frequenciesInHz, tdoasInSeconds):
    """exp(-2 pi i f tau) table, (F, D) complex128; gccNMF/gccNMFFunctions.py:89,128."""
    return np.exp(np.outer(frequenciesInHz, -(2j * np.pi) * tdoasInSeconds))


# ----------------------------------------------------------------------------- a1: STFT
def stft(y, n_fft, hop_length, window=np.hanning):
    """gccNMF/librosaSTFT.py:20-181 as called with center=False, win_length=n_fft.

    window(n_fft) is float64, so window*frame is float64 and scipy.fftpack.fft runs in double
    (:177); the first 1+n_fft/2 bins are conjugated (:179) and stored as complex64 (:21,:165).
    Frame count: 1 + int((len(y) - n_fft) / hop) (:425).
    """
    fft_window = np.asarray(window(n_fft) if callable(window) else window, dtype=np.float64).reshape(-1, 1)
    n_frames = 1 + int((len(y) - n_fft) / hop_length)
    if n_frames < 1:
        raise ValueError('Buffer is too short (n=%d) for frame_length=%d' % (len(y), n_fft))
    y = np.ascontiguousarray(y)
    frames = np.lib.stride_tricks.as_strided(y, shape=(n_fft, n_frames),
                                             strides=(y.itemsize, hop_length * y.itemsize))
    F = 1 + n_fft // 2
    out = np.empty((F, n_frames), dtype=np.complex64, order='F')
    n_columns = int(2 ** 8 * 2 ** 10 / (F * out.itemsize))  # MAX_MEM_BLOCK, librosaSTFT.py:18,170
    for s in range(0, n_frames, n_columns):
        e = min(s + n_columns, n_frames)
        out[:, s:e] = fftpack.fft(fft_window * frames[:, s:e], axis=0)[:F].conj()
    return out


def computeComplexMixtureSpectrogram(stereoSamples, windowSize, hopSize, windowFunction=np.hanning):
    """gccNMF/gccNMFFunctions.py:61-67 -> (2, F, T) complex64 (window argument ignored there: hanning)."""
    return np.array([stft(np.squeeze(stereoSamples[c]).copy(), windowSize, hopSize, np.hanning)
                     for c in range(2)])


# ----------------------------------------------------------------------------- a9: iSTFT
def istft(stft_matrix, hop_length, win_length=None, window=np.hanning, center=True):
    """gccNMF/librosaSTFT.py:183-286.

    Hermitian rebuild concat(conj(col), col[-2:0:-1]) (:278); complex64 input keeps
    scipy.fftpack.ifft in single precision (:279); `.real * window(float64)`; accumulation into a
    float32 buffer (:273,:281); `center=True` default trims n_fft/2 samples from each end (:283-284).
    """
    n_fft = 2 * (stft_matrix.shape[0] - 1)
    ifft_window = np.asarray(window(n_fft) if callable(window) else window, dtype=np.float64)
    n_frames = stft_matrix.shape[1]
    y = np.zeros(n_fft + hop_length * (n_frames - 1), dtype=np.float32)
    for i in range(n_frames):
        sample = i * hop_length
        spec = stft_matrix[:, i].flatten()
        spec = np.concatenate((spec.conj(), spec[-2:0:-1]), 0)
        ytmp = ifft_window * fftpack.ifft(spec).real
        y[sample:sample + n_fft] = y[sample:sample + n_fft] + ytmp
    if center:
        y = y[n_fft // 2:-(n_fft // 2)]
    return y


def getTargetSignalEstimates(targetSpectrogramEstimates, windowSize, hopSize, windowFunction=np.hanning):
    """gccNMF/gccNMFFunctions.py:153-163 -> (S, 2, hop*(T-1)) float32 * (2 hop / N)."""
    S, C = targetSpectrogramEstimates.shape[:2]
    gain = hopSize / float(windowSize) * 2
    return np.array([[istft(targetSpectrogramEstimates[s, c], hopSize, windowSize, windowFunction)
                      for c in range(C)] for s in range(S)]) * gain


# ----------------------------------------------------------------------------- a2: KL-NMF
def initKLNMF(numFrequencies, numColumns, dictionarySize, epsilon=1e-16, seedValue=0):
    """gccNMF/gccNMFFunctions.py:70-73: legacy global MT19937, W drawn first, then H."""
    np.random.seed(seedValue)
    W = np.random.random((numFrequencies, dictionarySize)).astype(np.float32) + epsilon
    H = np.random.random((dictionarySize, numColumns)).astype(np.float32) + epsilon
    return W, H


def klnmfIteration(V, W, H, sparsityAlpha=0, epsilon=1e-16):
    """One pass of gccNMF/gccNMFFunctions.py:76-81, in place on W, H."""
    H *= np.dot(W.T, V / np.dot(W, H)) / (np.sum(W, axis=0)[:, np.newaxis] + sparsityAlpha + epsilon)
    W *= np.dot(V / np.dot(W, H), H.T) / np.sum(H, axis=1)
    norms = np.sqrt(np.sum(W ** 2, 0))
    W /= norms
    H *= norms[:, np.newaxis]
    return W, H


def performKLNMF(V, dictionarySize, numIterations, sparsityAlpha, epsilon=1e-16, seedValue=0, W0=None, H0=None):
    """gccNMF/gccNMFFunctions.py:69-83 (optionally from caller-supplied W0, H0 for teacher forcing)."""
    if W0 is None:
        W, H = initKLNMF(V.shape[0], V.shape[1], dictionarySize, epsilon, seedValue)
    else:
        W, H = W0.copy(), H0.copy()
    for _ in range(numIterations):
        klnmfIteration(V, W, H, sparsityAlpha, epsilon)
    return W, H


def inferCoefficientsKLNMF(V, W, numIterations, sparsityAlpha, epsilon=1e-16, seedValue=0):
    """H-only KL-NMF with a fixed dictionary.  PARITY UNPINNED.

    Called at notebooks/onlineSpeechEnhancement.ipynb:433 and lowLatencySpeechEnhancement.ipynb:564
    but defined nowhere in the reference; restated from gccNMF/gccNMFFunctions.py:70,73 (seeded H
    init; W is not drawn because it is given) and :76 (H update) repeated numIterations times.
    """
    np.random.seed(seedValue)
    H = np.random.random((W.shape[1], V.shape[1])).astype(np.float32) + epsilon
    denom = np.sum(W, axis=0)[:, np.newaxis] + sparsityAlpha + epsilon
    for _ in range(numIterations):
        H *= np.dot(W.T, V / np.dot(W, H)) / denom
    return H


# ----------------------------------------------------------------------------- a3, a4, a5
def getSpectralCoherence(complexMixtureSpectrogram):
    """gccNMF/runGCCNMF.py:44 (unguarded; 0/0 -> NaN), complex64 in -> complex64 out."""
    X = complexMixtureSpectrogram
    return X[0] * X[1].conj() / np.abs(X[0]) / np.abs(X[1])


def getAngularSpectrogram(spectralCoherenceV, frequenciesInHz, microphoneSeparationInMetres, numTDOAs):
    """gccNMF/gccNMFFunctions.py:85-92, same einsum (materialises (D,F,T) complex128) -> (D,T) float64."""
    tdoasInSeconds = getTDOAsInSeconds(microphoneSeparationInMetres, numTDOAs)
    expJOmega = getExpJOmegaTau(frequenciesInHz, tdoasInSeconds)
    FREQ, TIME, TDOA = range(3)
    return np.sum(np.einsum(spectralCoherenceV, [FREQ, TIME], expJOmega, [FREQ, TDOA], [TDOA, FREQ, TIME]).real, axis=1)


def getAngularSpectrogramFast(spectralCoherenceV, frequenciesInHz, microphoneSeparationInMetres, numTDOAs):
    """Same quantity as getAngularSpectrogram via two real float64 GEMMs (test helper for big shapes):
    A[tau,t] = sum_f Re(C) Re(E) - Im(C) Im(E)."""
    E = getExpJOmegaTau(frequenciesInHz, getTDOAsInSeconds(microphoneSeparationInMetres, numTDOAs))
    C = spectralCoherenceV.astype(np.complex128)
    return E.real.T @ C.real - E.imag.T @ C.imag


def estimateTargetTDOAIndexesFromAngularSpectrum(angularSpectrum, microphoneSeparationInMetres, numTDOAs, numSources):
    """gccNMF/gccNMFFunctions.py:94-116, numSources branch (the KMeans branch references
    un-imported names in the reference and cannot run).  Too few peaks raises ValueError here
    (the reference logs and calls os._exit through an un-imported `os`)."""
    peakIndexes = argrelmax(angularSpectrum)[0]
    if not numSources:
        raise NotImplementedError('numSources=None branch is broken in the reference (:105-110)')
    sourcePeakIndexes = peakIndexes[np.argsort(angularSpectrum[peakIndexes])[-numSources:]]
    if len(sourcePeakIndexes) != numSources:
        raise ValueError('not enough peaks in the angular spectrum')
    return sorted(sourcePeakIndexes)


# ----------------------------------------------------------------------------- a6, a7, a8
def getTargetTDOAGCCNMFs(coherenceV, microphoneSeparationInMetres, numTDOAs, frequenciesInHz, targetTDOAIndexes, W, stereoH):
    """gccNMF/gccNMFFunctions.py:118-135 -> (S, K, T) float32."""
    hypothesisTDOAs = getTDOAsInSeconds(microphoneSeparationInMetres, numTDOAs)
    numChannels, numAtom, numTime = stereoH.shape
    expJOmegaTau = getExpJOmegaTau(frequenciesInHz, hypothesisTDOAs)
    TIME, FREQ, TDOA, ATOM = range(4)
    out = np.empty((len(targetTDOAIndexes), numAtom, numTime), np.float32)
    for i, tdoaIndex in enumerate(targetTDOAIndexes):
        gccChunk = np.einsum(coherenceV, [FREQ, TIME], expJOmegaTau[:, tdoaIndex], [FREQ], [FREQ, TIME])
        out[i] = np.einsum(W, [FREQ, ATOM], gccChunk, [FREQ, TIME], [ATOM, TIME]).real
    return out


def getTargetCoefficientMasks(targetTDOAGCCNMFs, numTargets):
    """gccNMF/gccNMFFunctions.py:137-143: nanargmax over targets -> one-hot float32 (S, K, T)."""
    nanArgMax = np.nanargmax(targetTDOAGCCNMFs, axis=0)
    masks = np.zeros_like(targetTDOAGCCNMFs)
    for s in range(numTargets):
        masks[s][np.where(nanArgMax == s)] = 1
    return masks


def getTargetSpectrogramEstimates(targetCoefficientMasks, complexMixtureSpectrogram, W, stereoH):
    """gccNMF/gccNMFFunctions.py:145-151 -> (S, 2, F, T) complex64."""
    S = targetCoefficientMasks.shape[0]
    est = np.zeros((S,) + complexMixtureSpectrogram.shape, np.complex64)
    for s, mask in enumerate(targetCoefficientMasks):
        for c, coefficients in enumerate(stereoH):
            est[s, c] = np.dot(W, coefficients * mask)
    return est * np.exp(1j * np.angle(complexMixtureSpectrogram))


# ----------------------------------------------------------------------------- a10: enhancement mask
def getGCCNMFAllTDOAs(spectralCoherenceV, expJOmegaTau, W):
    """notebooks/offlineSpeechEnhancement.ipynb cell 27 (:444-450): per-frame
    dot((C[:,t,None] * E).real.T, W) stacked and transposed -> (K, D, T) float64."""
    numTime = spectralCoherenceV.shape[1]
    gccNMFs = []
    for t in range(numTime):
        gccNMFs.append(np.dot((spectralCoherenceV[:, t, np.newaxis] * expJOmegaTau).real.T, W))
    return np.array(gccNMFs).T


def getEnhancementCoefficientMask(gccNMF, hypothesisTDOAs, targetTDOAIndex, targetTDOAWindowSizePercent=0.05):
    """notebooks/offlineSpeechEnhancement.ipynb cell 29 (:466-472) and cell 4 (:117-118).
    Returns (argmax (K,T) int, mask (1,K,T) bool)."""
    targetTDOAWindowSize = (hypothesisTDOAs[-1] - hypothesisTDOAs[0]) * targetTDOAWindowSizePercent
    argMaxGCCNMF = np.argmax(gccNMF, axis=1)
    gccNMFMaxTDOA = np.take(hypothesisTDOAs, argMaxGCCNMF)
    distanceToTargetTDOA = np.abs(gccNMFMaxTDOA - hypothesisTDOAs[targetTDOAIndex])
    return argMaxGCCNMF, np.array([distanceToTargetTDOA < targetTDOAWindowSize])


# ----------------------------------------------------------------------------- full offline flows
def runSeparation(stereoSamples, sampleRate, windowSize, hopSize, numTDOAs, microphoneSeparationInMetres,
                  numTargets, dictionarySize=128, numIterations=100, sparsityAlpha=0, stages=None):
    """gccNMF/runGCCNMF.py:36-52 in the same order (NMF parameters exposed instead of the
    hard-coded 128/100 of :41).  Returns a dict of every intermediate."""
    r = {}
    X = r['X'] = computeComplexMixtureSpectrogram(stereoSamples, windowSize, hopSize)
    numChannels, numFrequencies, numTime = X.shape
    f = r['frequenciesInHz'] = np.linspace(0, sampleRate / 2.0, numFrequencies)
    V = r['V'] = np.concatenate(np.abs(X), axis=-1)
    W, H = performKLNMF(V, dictionarySize, numIterations, sparsityAlpha)
    r['W'], r['H'] = W, H
    stereoH = r['stereoH'] = np.array(np.hsplit(H, numChannels))
    coh = r['coherence'] = getSpectralCoherence(X)
    A = r['angularSpectrogram'] = getAngularSpectrogram(coh, f, microphoneSeparationInMetres, numTDOAs)
    m = r['meanAngularSpectrum'] = np.mean(A, axis=-1)
    idx = r['targetTDOAIndexes'] = estimateTargetTDOAIndexesFromAngularSpectrum(m, microphoneSeparationInMetres, numTDOAs, numTargets)
    G = r['targetTDOAGCCNMFs'] = getTargetTDOAGCCNMFs(coh, microphoneSeparationInMetres, numTDOAs, f, idx, W, stereoH)
    M = r['targetCoefficientMasks'] = getTargetCoefficientMasks(G, numTargets)
    S = r['targetSpectrogramEstimates'] = getTargetSpectrogramEstimates(M, X, W, stereoH)
    r['targetSignalEstimates'] = getTargetSignalEstimates(S, windowSize, hopSize)
    return r


def runEnhancement(stereoSamples, sampleRate, windowSize, hopSize, numTDOAs, microphoneSeparationInMetres,
                   dictionarySize, numIterations, sparsityAlpha=0, targetTDOAWindowSizePercent=0.05,
                   fastAngularSpectrogram=False, timings=None):
    """notebooks/offlineSpeechEnhancement.ipynb cells 12-41 in the notebook's order
    (STFT, coherence + angular spectrogram + peak, KL-NMF, all-TDOA GCC-NMF, argmax mask,
    reconstruction, iSTFT) with numSources = 1.  Returns a dict of every intermediate."""
    import time
    r = {}
    tm = timings if timings is not None else {}
    t0 = time.perf_counter()
    X = r['X'] = computeComplexMixtureSpectrogram(stereoSamples, windowSize, hopSize)
    numChannels, numFrequencies, numTime = X.shape
    f = r['frequenciesInHz'] = getFrequenciesInHz(sampleRate, numFrequencies)
    tm['stft'] = time.perf_counter() - t0; t0 = time.perf_counter()
    hypothesisTDOAs = getTDOAsInSeconds(microphoneSeparationInMetres, numTDOAs)
    coh = r['coherence'] = getSpectralCoherence(X)
    angfn = getAngularSpectrogramFast if fastAngularSpectrogram else getAngularSpectrogram
    A = r['angularSpectrogram'] = angfn(coh, f, microphoneSeparationInMetres, numTDOAs)
    m = r['meanAngularSpectrum'] = np.mean(A, axis=-1)
    idx = r['targetTDOAIndexes'] = estimateTargetTDOAIndexesFromAngularSpectrum(m, microphoneSeparationInMetres, numTDOAs, 1)
    tm['angular'] = time.perf_counter() - t0; t0 = time.perf_counter()
    V = r['V'] = np.concatenate(np.abs(X), axis=-1)
    W, H = performKLNMF(V, dictionarySize, numIterations, sparsityAlpha)
    r['W'], r['H'] = W, H
    stereoH = r['stereoH'] = np.array(np.hsplit(H, numChannels))
    tm['nmf'] = time.perf_counter() - t0; t0 = time.perf_counter()
    E = getExpJOmegaTau(f, hypothesisTDOAs)
    gccNMF = getGCCNMFAllTDOAs(coh, E, W)
    tm['gccnmf'] = time.perf_counter() - t0; t0 = time.perf_counter()
    r['argMaxGCCNMF'], M = getEnhancementCoefficientMask(gccNMF, hypothesisTDOAs, idx[0], targetTDOAWindowSizePercent)
    r['targetCoefficientMasks'] = M
    tm['mask'] = time.perf_counter() - t0; t0 = time.perf_counter()
    S = r['targetSpectrogramEstimates'] = getTargetSpectrogramEstimates(M, X, W, stereoH)
    tm['recon'] = time.perf_counter() - t0; t0 = time.perf_counter()
    r['targetSignalEstimates'] = getTargetSignalEstimates(S, windowSize, hopSize)
    tm['istft'] = time.perf_counter() - t0
    return r


# ----------------------------------------------------------------------------- a11: online / low-latency
def getAsymmetricAnalysisWindow(k, m, d):
    """notebooks/lowLatencySpeechEnhancement.ipynb cell 21 (:371-380)."""
    rising = np.sqrt(np.hanning(2 * (k - m - d) + 1)[:2 * (k - m - d)])
    falling = np.sqrt(np.hanning(2 * m + 1)[:2 * m])
    window = np.zeros(k)
    window[d:k - m] = rising[:k - m - d]
    window[k - m:] = falling[-m:]
    return window


def getAsymmetricSynthesisWindow(k, m, d):
    """notebooks/lowLatencySpeechEnhancement.ipynb cell 21 (:382-392).  Built by the notebook but
    never multiplied into the output frames (:575-578); kept for completeness."""
    risingAnalysis = np.sqrt(np.hanning(2 * (k - m - d) + 1)[:2 * (k - m - d)])
    risingNormalized = np.hanning(2 * m + 1)[:m] / risingAnalysis[k - 2 * m - d:k - m - d]
    falling = np.sqrt(np.hanning(2 * m + 1)[:2 * m])
    window = np.zeros(k)
    window[-2 * m:-m] = risingNormalized
    window[-m:] = falling[-m:]
    return window


def performOnlineSpeechEnhancement(stereoSamples, sampleRate, W, analysisWindow, synthesisWindowLength, hopSize,
                                   numTDOAs, microphoneSeparationInMetres, targetTDOAEpsilon,
                                   numInferenceIterations=0, sparsityAlpha=0, epsilon=1e-16, seedValue=0,
                                   gainPerFrame=False):
    """Frame loop of notebooks/onlineSpeechEnhancement.ipynb cell 23 (:406-447) and
    lowLatencySpeechEnhancement.ipynb cell 30 (:511-584).

    numpy.fft.rfft (NOT conjugated, float64); accumulated-max GCC-PHAT localisation; per-frame
    (D x F).(F x K) GCC-NMF; argmax over TDOA per atom; 0/1 atom mask; Wiener-like filter;
    irfft; overlap-add of the raw irfft frame (no synthesis window), gain 2 hop / len(synthesis).
    gainPerFrame=True applies the gain per frame as the online notebook does (:447), False once
    at the end as the low-latency notebook does (:580).
    """
    numSamples = stereoSamples.shape[1]
    analysisWindowSize = len(analysisWindow)
    numFrequencies, dictionarySize = W.shape
    if gainPerFrame:
        numFrames = (numSamples - analysisWindowSize) // hopSize          # online :376
    else:
        numFrames = (numSamples - synthesisWindowLength) // hopSize       # low-latency :513
    gainFactor = hopSize / float(synthesisWindowLength) * 2
    frequenciesInHz = getFrequenciesInHz(sampleRate, numFrequencies)
    expJOmegaTau = getExpJOmegaTau(frequenciesInHz, getTDOAsInSeconds(microphoneSeparationInMetres, numTDOAs))
    out = np.zeros_like(stereoSamples)
    accMax = np.full(numTDOAs, -np.inf)
    targetTDOAs = np.full(numFrames, np.nan)
    angularSpectrogram = np.zeros((numTDOAs, numFrames))
    atomMasks = np.zeros((dictionarySize, numFrames))
    wienerFilters = np.zeros((2, numFrequencies, numFrames))
    for frameIndex in range(numFrames):
        s = frameIndex * hopSize
        e = s + analysisWindowSize
        X = np.fft.rfft(stereoSamples[:, s:e] * analysisWindow)
        coh = X[0] * X[1].conj() / np.abs(X[0]) / np.abs(X[1])
        gccPHAT = np.dot(coh, expJOmegaTau).real
        accMax = np.max(np.array([gccPHAT, accMax]), axis=0)
        target = np.argmax(accMax)
        targetTDOAs[frameIndex] = target
        angularSpectrogram[:, frameIndex] = gccPHAT
        gccNMF = np.dot((coh[:, np.newaxis] * expJOmegaTau).real.T, W)
        est = np.argmax(gccNMF, axis=0)
        atomMask = np.zeros(dictionarySize)
        atomMask[np.abs(est - target) < targetTDOAEpsilon] = 1
        atomMasks[:, frameIndex] = atomMask
        if numInferenceIterations == 0:
            wiener = np.sum(atomMask * W, axis=1) / np.sum(W, axis=1)
        else:
            stereoH = inferCoefficientsKLNMF(np.abs(X).T, W, numInferenceIterations, sparsityAlpha, epsilon, seedValue)
            wiener = (np.dot(W, stereoH * atomMask[:, np.newaxis]) / np.dot(W, stereoH)).T
        wienerFilters[:, :, frameIndex] = wiener
        rec = np.fft.irfft(wiener * X)
        if gainPerFrame:
            out[:, s:e] += rec * gainFactor
        else:
            out[:, s:e] += rec
    if not gainPerFrame:
        out *= gainFactor
    return dict(output=out, targetTDOAs=targetTDOAs, angularSpectrogram=angularSpectrogram,
                atomMasks=atomMasks, wienerFilters=wienerFilters, gccPHATAccumulatedMax=accMax)


# ----------------------------------------------------------------------------- a13: RT chunk processor
TARGET_MODE_BOXCAR = 0               # gccNMFProcessor.py:35-37
TARGET_MODE_MULTIPLE = 1
TARGET_MODE_WINDOW_FUNCTION = 2


class GCCNMFProcessorOracle(object):
    """numpy restatement of gccNMF/realtime/gccNMFProcessor.py:167-276.  The reference class needs Theano, which is absent
    here; it is PINNED to the output of the unmodified reference class executed over oracle/theano_numpy_shim.py (a numpy
    evaluation of the reference's own graph): tests/golden/realtime_mini.npz, tests/test_oracle_golden.py.  What that cannot
    pin is Theano's own arithmetic (BLAS summation order, elementwise kernels) -- say "pinned to the reference's code on numpy
    arithmetic" when quoting it.  dtypes follow numpy's promotion, which for this graph is also Theano's: the float32 dot,
    then float64 from `argmax (int64) - targetTDOAIndex (float32)` onwards.

    Localisation history: the (numTDOAs, numTDOAHistory) float64 ring of realtime/utils.py:34-65
    (zero-initialised) holding nanmean-over-frequency GCC-PHAT columns (:214-215), kept here in
    time order; target index = argmax of nanmean over the last localizationWindowSize columns (:221-226).
    """

    def __init__(self, sampleRate, windowSize, numTimePerChunk, W, numTDOAs, microphoneSeparationInMetres,
                 localizationEnabled=False, localizationWindowSize=6, historyLength=128):
        self.W = np.asarray(W, np.float32)
        self.numFrequencies, self.numAtom = self.W.shape
        self.numTDOAs = numTDOAs
        self.windowFunction = np.sqrt(np.hamming(windowSize).astype(np.float32))[:, np.newaxis]   # :186
        f = np.linspace(0, sampleRate / 2, self.numFrequencies).astype(np.float32)                # :245
        maxTDOA = microphoneSeparationInMetres / SPEED_OF_SOUND_IN_METRES_PER_SECOND
        tdoas = np.linspace(-maxTDOA, maxTDOA, numTDOAs).astype(np.float32)                       # :247
        self.expJOmegaTau = np.exp(np.outer(f, -(2j * np.pi) * tdoas)).astype(np.complex64)       # :248
        self.separationEnabled = True
        self.localizationEnabled = localizationEnabled
        self.localizationWindowSize = localizationWindowSize
        self.targetMode = TARGET_MODE_WINDOW_FUNCTION
        self.targetTDOAIndex = np.float32(10.0)      # :196-199
        self.targetTDOAEpsilon = np.float32(2.0)
        self.targetTDOABeta = np.float32(1.0)
        self.targetTDOANoiseFloor = np.float32(0.0)
        self.gccPHATHistory = np.zeros((numTDOAs, historyLength), np.float64)   # utils.py:35-38 (initValue=0, c_double)

    def setTargetTDOARange(self, idx, eps, beta, floor):
        """:272-276."""
        self.targetTDOAIndex, self.targetTDOAEpsilon = np.float32(idx), np.float32(eps)
        self.targetTDOABeta, self.targetTDOANoiseFloor = np.float32(beta), np.float32(floor)

    def processFrames(self, windowedSamples):
        """:201-231.  windowedSamples (2, N, nT) float32 -> (2, N, nT)."""
        X = np.fft.rfft(windowedSamples * self.windowFunction, axis=1).astype(np.complex64)        # :202
        coh = X[0] * X[1].conj() / np.abs(X[0]) / np.abs(X[1])                                     # :253
        realGCC = (coh[:, :, np.newaxis] * self.expJOmegaTau[:, np.newaxis]).real                  # :254, (F,nT,D)
        self.lastRealGCC = realGCC
        if self.separationEnabled:
            gccNMF = np.dot(realGCC.T, self.W)                                                     # :259 (D,nT,K) float32
            dist = np.abs(np.argmax(gccNMF, axis=0).T - self.targetTDOAIndex)                      # (K,nT): int64 - float32 -> float64
            self.lastArgmax = np.argmax(gccNMF, axis=0).T
            if self.targetMode == TARGET_MODE_BOXCAR:
                HMask = np.where(dist < self.targetTDOAEpsilon, 1.0, 0.0)                          # :263
            else:
                HMask = (np.exp(-(dist / self.targetTDOAEpsilon) ** self.targetTDOABeta)
                         / (1 + self.targetTDOANoiseFloor) + self.targetTDOANoiseFloor)            # :265
            tfMask = (np.dot(self.W, HMask).T / np.sum(self.W, axis=-1, keepdims=False)).T         # :267-269
            self.lastHMask, self.lastTFMask = HMask, tfMask
            out = tfMask * X                                                                       # :209
        else:
            out = X.copy()
        with np.errstate(all='ignore'):
            newColumns = np.nanmean(realGCC, axis=0).T                                              # :214-215 (D,nT)
        nT = newColumns.shape[-1]
        self.gccPHATHistory = np.roll(self.gccPHATHistory, -nT, axis=1)   # time-ordered view of utils.py:45-65
        self.gccPHATHistory[:, -nT:] = newColumns
        if self.localizationEnabled:
            with np.errstate(all='ignore'):
                self.targetTDOAIndex = np.float32(np.argmax(
                    np.nanmean(self.gccPHATHistory[:, -self.localizationWindowSize:], axis=-1)))   # :221-226
        return np.fft.irfft(out, axis=1) * self.windowFunction                                     # :231


class OverlapAddProcessorOracle(object):
    """gccNMF/realtime/utils.py:72-116 without the shared-memory plumbing."""

    def __init__(self, numChannels, windowSize, hopSize, blockSize, windowsPerBlock):
        self.windowSize, self.hopSize, self.blockSize, self.windowsPerBlock = windowSize, hopSize, blockSize, windowsPerBlock
        self.bufferSize = blockSize * 8
        self.inputBuffer = np.zeros((numChannels, self.bufferSize), np.float32)
        self.outputBuffer = np.zeros((numChannels, self.bufferSize), np.float32)
        self.windowedSamples = np.zeros((numChannels, windowSize, windowsPerBlock), np.float32)

    def processFrames(self, inputFrames, processFramesFunction):
        B = self.blockSize
        self.inputBuffer[:, :-B] = self.inputBuffer[:, B:]
        self.inputBuffer[:, -B:] = inputFrames
        self.outputBuffer[:, :-B] = self.outputBuffer[:, B:]
        self.outputBuffer[:, -B:] = 0
        idx = np.arange(self.bufferSize - self.windowSize - (self.windowsPerBlock - 1) * self.hopSize,
                        self.bufferSize - self.windowSize + 1, self.hopSize)
        for i, w in enumerate(idx):
            self.windowedSamples[..., i] = self.inputBuffer[:, w:w + self.windowSize]
        processed = processFramesFunction(self.windowedSamples)
        for i, w in enumerate(idx):
            self.outputBuffer[:, w:w + self.windowSize] += processed[..., i]
        return self.outputBuffer[:, -3 * B:-2 * B].copy()


# ----------------------------------------------------------------------------- a14: pre-training helpers
def getOrderedDictionary(W):
    """gccNMF/realtime/gccNMFPretraining.py:60-66: atoms sorted by spectral centroid."""
    numFreq = W.shape[0]
    centroids = np.squeeze(np.sum(np.arange(numFreq)[:, np.newaxis] * W, axis=0, keepdims=True) / np.sum(W, axis=0, keepdims=True))
    return np.squeeze(W[:, np.argsort(centroids)])
