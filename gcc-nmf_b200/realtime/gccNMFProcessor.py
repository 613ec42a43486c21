"""Drop-in for the per-chunk processor of gccNMF/realtime/gccNMFProcessor.py:167-276.

Same constructor arguments, attributes and methods (`processFrames`, `reset`, `setTargetTDOARange`,
settable `numTDOAs`, `separationEnabled`, `localizationEnabled`, `localizationWindowSize`, `targetMode`); the
Theano graph (:238-270) is replaced by the sm_100a kernels: windowed FFT -> PHAT coherence + real GCC ->
GCC-NMF argmax over TDOA per atom -> boxcar / window atom mask -> Wiener-like TF mask (W.mask)/rowsum(W) ->
inverse FFT x synthesis window.  `numHUpdates` is accepted and ignored exactly as in the reference (it is
plumbed everywhere there but never used).  History buffers (the reference's SharedMemoryCircularBuffer
objects) are optional duck-typed objects with `.set(values)` / `.getUnraveledArray()`.

The argmax over TDOA is the exact float64 one (tensor cores + float64 refinement); the reference evaluates it in
Theano float32 (and cannot be run here: Theano is not installed), so near-ties may resolve differently.
"""
import logging

import numpy as np

from .. import gccNMFFunctions as fn
from .._lib import default_handle

TARGET_MODE_BOXCAR = 0
TARGET_MODE_WINDOW_FUNCTION = 1


class GCCNMFProcessor(object):
    def __init__(self, sampleRate, windowSize, numTimePerChunk, dictionariesW, dictionaryType, dictionarySize, numHUpdates,
                 microphoneSeparationInMetres, localizationEnabled, localizationWindowSize, gccPHATHistory=None, tdoaHistory=None,
                 inputSpectrogramHistory=None, outputSpectrogramHistory=None, coefficientMaskHistories=None, device=0):
        self.sampleRate = sampleRate
        self.windowSize = windowSize
        self.numTimePerChunk = numTimePerChunk
        self.dictionariesW = dictionariesW
        self.dictionaryType = dictionaryType
        self.dictionarySize = dictionarySize
        self.numHUpdates = numHUpdates
        self.microphoneSeparationInMetres = microphoneSeparationInMetres
        self.gccPHATHistory = gccPHATHistory
        self.tdoaHistory = tdoaHistory
        self.inputSpectrogramHistory = inputSpectrogramHistory
        self.outputSpectrogramHistory = outputSpectrogramHistory
        self.coefficientMaskHistories = coefficientMaskHistories
        self.windowFunction = np.sqrt(np.hamming(self.windowSize).astype(np.float32))[:, np.newaxis]    # :186
        self.synthesisWindowFunction = self.windowFunction
        self.numTDOAs = None
        self.separationEnabled = True
        self.localizationEnabled = localizationEnabled
        self.localizationWindowSize = localizationWindowSize
        self.targetMode = TARGET_MODE_WINDOW_FUNCTION
        self.targetTDOAIndex = np.float32(10.0)      # :196-199 (Theano shared scalars in the reference)
        self.targetTDOAEpsilon = np.float32(2.0)
        self.targetTDOABeta = np.float32(1.0)
        self.targetTDOANoiseFloor = np.float32(0.0)
        self.h = default_handle(device)
        self._built = False

    # ------------------------------------------------------------------ :233-270
    def reset(self):
        logging.info('GCCNMFProcessor: resetting...')
        self.buildFunctions()
        logging.info('GCCNMFProcessor: done reset.')

    def buildFunctions(self):
        """Device constants that the reference bakes into its Theano functions (:241-248)."""
        h = self.h
        self.W = np.ascontiguousarray(self.dictionariesW[self.dictionaryType][self.dictionarySize], dtype=np.float32)
        self.numFrequencies, self.numAtom = self.W.shape
        self.frequenciesInHz = np.linspace(0, self.sampleRate / 2, self.numFrequencies).astype(np.float32)
        self.maxTDOA = self.microphoneSeparationInMetres / fn.SPEED_OF_SOUND_IN_METRES_PER_SECOND
        self.hypothesisTDOAs = np.linspace(-self.maxTDOA, self.maxTDOA, self.numTDOAs).astype(np.float32)
        self.expJOmegaTau = np.exp(np.outer(self.frequenciesInHz, -(2j * np.pi) * self.hypothesisTDOAs)).astype(np.complex64)
        self._W = h.to_device(self.W)
        self._E = h.to_device(np.ascontiguousarray(self.expJOmegaTau.astype(np.complex128)))
        self._analysis = h.to_device(self.windowFunction[:, 0].astype(np.float64))
        self._synthesis = h.to_device(self.synthesisWindowFunction[:, 0].astype(np.float64))
        self._built = True

    buildTheanoFunctions = buildFunctions      # the reference's name (:238)

    def setTargetTDOARange(self, targetTDOAIndex, targetTDOAEpsilon, targetTDOABeta, targetTDOANoiseFloor):
        """:272-276."""
        self.targetTDOAIndex = np.float32(targetTDOAIndex)
        self.targetTDOAEpsilon = np.float32(targetTDOAEpsilon)
        self.targetTDOABeta = np.float32(targetTDOABeta)
        self.targetTDOANoiseFloor = np.float32(targetTDOANoiseFloor)

    # ------------------------------------------------------------------ :201-231
    def processFrames(self, windowedSamples):
        """windowedSamples (2, N, nT) float32 -> (2, N, nT) float32."""
        if not self._built:
            self.buildFunctions()
        h = self.h
        windowedSamples = np.asarray(windowedSamples, dtype=np.float32)
        _, N, nT = windowedSamples.shape
        # frames as a non-overlapping signal: (2, nT * N), hop = N
        x = h.to_device(np.ascontiguousarray(windowedSamples.transpose(0, 2, 1)).reshape(2, nT * N))
        X = h.stft(x, self._analysis, N, N, conjugate=False)                                   # :202 rfft(frames * window)
        coh, ang, _ = h.phat_angspec(X, self._E, want_mean=False)                              # :253-255 (sum over f of realGCC = F * nanmean)
        if self.separationEnabled:
            argmax, refined = h.tdoa_argmax(coh, self._E, self._W)                             # :259 + argmax of :263/:265
            if int(refined.item()) > h.lib.gccnmf_tdoa_argmax_refine_capacity(self.numAtom, nT):
                _, argmax = h.tdoa_gccnmf(coh, self._E, self._W, want_values=False, want_argmax=True)
            mode = 0 if self.targetMode == TARGET_MODE_BOXCAR else 1
            mask = h.atom_mask(argmax, targets=None, target_scalar=float(self.targetTDOAIndex), epsilon=float(self.targetTDOAEpsilon),
                               mode=mode, beta=float(self.targetTDOABeta), noise_floor=float(self.targetTDOANoiseFloor))
            out = h.wiener_apply(mask, self._W, X)                                             # :267-269, :209
            if self.coefficientMaskHistories:
                self.coefficientMaskHistories[self.dictionarySize].set(1 - mask.cpu().numpy())
        else:
            out = X
        if self.inputSpectrogramHistory:
            self.inputSpectrogramHistory.set(-np.mean(np.abs(X.cpu().numpy()), axis=0) ** (1 / 3.0))
        gccPHAT = ang.cpu().numpy() / float(self.numFrequencies)                               # :214 nanmean over frequency, (D, nT)
        if self.gccPHATHistory:
            self.gccPHATHistory.set(gccPHAT)
        if self.tdoaHistory:
            if self.localizationEnabled:
                history = self.gccPHATHistory.getUnraveledArray()
                with np.errstate(all='ignore'):
                    tdoaIndex = np.argmax(np.nanmean(history[:, -self.localizationWindowSize:], axis=-1))   # :221-222
                self.targetTDOAIndex = np.float32(tdoaIndex)
            self.tdoaHistory.set(np.array([[self.targetTDOAIndex]]))
        if self.outputSpectrogramHistory:
            self.outputSpectrogramHistory.set(-np.nanmean(np.abs(out.cpu().numpy()), axis=0) ** (1 / 3.0))
        y = h.istft_ola(out, self._synthesis, N, N, gain=1.0, center=False, conjugate=False)   # :231 irfft * synthesis window
        return np.ascontiguousarray(y.cpu().numpy().reshape(2, nT, N).transpose(0, 2, 1))
