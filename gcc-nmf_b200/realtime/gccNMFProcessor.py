"""Drop-in for the per-chunk processor of gccNMF/realtime/gccNMFProcessor.py:167-276.

Same constructor arguments, attributes and methods (`processFrames`, `reset`, `setTargetTDOARange`, settable `numTDOAs`,
`separationEnabled`, `localizationEnabled`, `localizationWindowSize`, `targetMode`).  The Theano graph (:238-270) and the
numpy code around it (:201-231) run as ONE stream-ordered sequence of sm_100a kernels without a host synchronisation inside
(csrc/rt.cu, `RealtimeEngine`): windowed FFT -> PHAT coherence -> real GCC -> float32 GCC-NMF `dot(realGCC.T, W)` -> argmax
over TDOA per atom -> boxcar / window atom mask (float64, like the int64 - float32 promotion of the reference's graph) ->
(W . mask) / rowsum(W) -> inverse FFT x synthesis window, plus the GCC-PHAT history ring and the sliding-window
localisation, whose target TDOA index stays on the device from one call to the next.  Per call the host does one H2D of
the windowed frames, one D2H of the result and one synchronisation.

`numHUpdates` is accepted and, exactly as in the reference, unused (it is plumbed everywhere there and never read);
`coefficientInferenceIterations` (an extension, default 0) switches on the per-frame H-only KL updates of
notebooks/onlineSpeechEnhancement.ipynb:433-438 in the same kernel sequence.  History buffers (the reference's
SharedMemoryCircularBuffer objects) are optional duck-typed objects with `.set(values)` / `.getUnraveledArray()`.
"""
import logging

import numpy as np

from .. import gccNMFFunctions as fn
from . import engine as rt

TARGET_MODE_BOXCAR = 0               # gccNMFProcessor.py:35-37
TARGET_MODE_MULTIPLE = 1             # declared by the reference, not implemented there (its graph has no such branch, :262-265)
TARGET_MODE_WINDOW_FUNCTION = 2


class GCCNMFProcessor(object):
    def __init__(self, sampleRate, windowSize, numTimePerChunk, dictionariesW, dictionaryType, dictionarySize, numHUpdates,
                 microphoneSeparationInMetres, localizationEnabled, localizationWindowSize, gccPHATHistory=None, tdoaHistory=None,
                 inputSpectrogramHistory=None, outputSpectrogramHistory=None, coefficientMaskHistories=None, device=0,
                 coefficientInferenceIterations=0):
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
        self.coefficientInferenceIterations = coefficientInferenceIterations
        self.device = device
        self.engine = None
        self._sent = None
        self._target_dirty = True

    # ------------------------------------------------------------------ :233-270
    def reset(self):
        logging.info('GCCNMFProcessor: resetting...')
        self.buildFunctions()
        logging.info('GCCNMFProcessor: done reset.')

    def buildFunctions(self, hopSize=None, blockSize=None):
        """Device constants that the reference bakes into its Theano functions (:241-248) and the device-resident state.
        hopSize / blockSize only matter for the ring entry (`processBlock`); `processFrames` gets its frames cut by the caller."""
        self.W = np.ascontiguousarray(self.dictionariesW[self.dictionaryType][self.dictionarySize], dtype=np.float32)
        self.numFrequencies, self.numAtom = self.W.shape
        self.frequenciesInHz = np.linspace(0, self.sampleRate / 2, self.numFrequencies).astype(np.float32)
        self.maxTDOA = self.microphoneSeparationInMetres / fn.SPEED_OF_SOUND_IN_METRES_PER_SECOND
        self.hypothesisTDOAs = np.linspace(-self.maxTDOA, self.maxTDOA, self.numTDOAs).astype(np.float32)
        self.expJOmegaTau = np.exp(np.outer(self.frequenciesInHz, -(2j * np.pi) * self.hypothesisTDOAs)).astype(np.complex64)
        historyLength = self.gccPHATHistory.size() if self.gccPHATHistory else 128
        if self.engine is not None:
            self.engine.close()
        self._geometry = (int(hopSize or self.windowSize), int(blockSize or self.windowSize * self.numTimePerChunk))
        self.engine = rt.RealtimeEngine(self.W, self.expJOmegaTau, self.windowFunction[:, 0], self.synthesisWindowFunction[:, 0],
                                        hopSize=self._geometry[0], blockSize=self._geometry[1], windowsPerBlock=self.numTimePerChunk,
                                        historyLength=historyLength, numInferenceIterations=self.coefficientInferenceIterations,
                                        device=self.device)
        if self.targetMode not in (TARGET_MODE_BOXCAR, TARGET_MODE_WINDOW_FUNCTION):
            raise ValueError('targetMode %r: the reference builds a mask for TARGET_MODE_BOXCAR and TARGET_MODE_WINDOW_FUNCTION only' % (self.targetMode,))
        self._builtTargetMode = self.targetMode        # the reference bakes the mode into the graph at build time (:262-265)
        self._sent = None
        self._target_dirty = True

    buildTheanoFunctions = buildFunctions      # the reference's name (:238)

    def setTargetTDOARange(self, targetTDOAIndex, targetTDOAEpsilon, targetTDOABeta, targetTDOANoiseFloor):
        """:272-276."""
        self.targetTDOAIndex = np.float32(targetTDOAIndex)
        self.targetTDOAEpsilon = np.float32(targetTDOAEpsilon)
        self.targetTDOABeta = np.float32(targetTDOABeta)
        self.targetTDOANoiseFloor = np.float32(targetTDOANoiseFloor)
        self._target_dirty = True

    def _sync_params(self):
        """Pushes the Python-side attributes to the device when they changed (stream-ordered, no synchronisation)."""
        localize = bool(self.tdoaHistory) and bool(self.gccPHATHistory) and bool(self.localizationEnabled)     # :216-222
        now = (float(self.targetTDOAEpsilon), float(self.targetTDOABeta), float(self.targetTDOANoiseFloor), int(self._builtTargetMode),
               bool(self.separationEnabled), localize, int(self.localizationWindowSize))
        if self._target_dirty or now != self._sent:
            self.engine.set_params(float(self.targetTDOAIndex) if self._target_dirty else None, now[0], now[1], now[2],
                                   0 if now[3] == TARGET_MODE_BOXCAR else 1, now[4], now[5], now[6])
            self._sent = now
            self._target_dirty = False

    def _mirror_histories(self):
        """Optional host-side mirrors for a GUI; the numbers are the device state of the call that just finished."""
        e = self.engine
        if self.separationEnabled and self.coefficientMaskHistories:
            self.coefficientMaskHistories[self.dictionarySize].set(1 - e.export(rt.EXPORT_ATOM_MASK))
        if self.inputSpectrogramHistory:
            self.inputSpectrogramHistory.set(-np.mean(np.abs(e.export(rt.EXPORT_INPUT_SPEC)), axis=0) ** (1 / 3.0))
        if self.gccPHATHistory:
            self.gccPHATHistory.set(e.export(rt.EXPORT_GCCPHAT))
        if self.tdoaHistory:
            self.targetTDOAIndex = np.float32(e.export(rt.EXPORT_TARGET)[0])      # the device took the localisation decision (:221-225)
            self.tdoaHistory.set(np.array([[self.targetTDOAIndex]]))
        if self.outputSpectrogramHistory:
            with np.errstate(all='ignore'):
                self.outputSpectrogramHistory.set(-np.nanmean(np.abs(e.export(rt.EXPORT_OUTPUT_SPEC)), axis=0) ** (1 / 3.0))

    # ------------------------------------------------------------------ :201-231
    def processFrames(self, windowedSamples, forcedAtomMask=None):
        """windowedSamples (2, N, nT) float32 -> (2, N, nT) float32.  forcedAtomMask (K, nT): use this atom mask instead of the
        one derived from the TDOA argmax (teacher-forced parity tests)."""
        if self.engine is None:
            self.buildFunctions()
        self._sync_params()
        out = self.engine.process_frames(np.asarray(windowedSamples, dtype=np.float32), forcedAtomMask).copy()
        self._mirror_histories()
        return out

    def processBlock(self, inputFrames, hopSize, blockSize, useGraph=True, forcedAtomMask=None):
        """One audio block through the device-resident overlap-add rings AND processFrames as a single CUDA graph launch:
        OverlapAddProcessor.processFrames(self.processFrames) of gccNMF/realtime/utils.py:99-116 / gccNMFProcessor.py:97.
        inputFrames (2, blockSize) float32 -> the next output block (2, blockSize) float32 (two blocks of latency, utils.py:115)."""
        if self.engine is None or self._geometry != (int(hopSize), int(blockSize)):
            self.buildFunctions(hopSize, blockSize)
        self._sync_params()
        out = self.engine.process_block(np.asarray(inputFrames, dtype=np.float32), use_graph=useGraph, forcedAtomMask=forcedAtomMask)
        self._mirror_histories()
        return out
