"""Headless real-time runner: wav file -> blocks -> GCC-NMF speech enhancement -> wav file.

Mirrors the no-GUI path of the reference (gccNMF/realtime/runRealtimeGCCNMF.py:122-179 `RealtimeGCCNMFNoGUI`,
the file-player callback of gccNMF/realtime/audioProcessor.py:106-132, the defaults of gccNMF/realtime/config.py:46-82
and the parameter messages of `initParams`, runRealtimeGCCNMF.py:141-167) without its process shell: the reference runs
the audio callback and the GCC-NMF processor in two OS processes that hand one block at a time through shared arrays and
a pair of Events, i.e. strictly synchronously -- here the same two steps are called in sequence per block:

    inputFrames <- next blockSize samples of the file          (audioProcessor.py:112-116, int16 -> float32 / 32768)
    oladProcessor.processFrames(gccNMFProcessor.processFrames) (gccNMFProcessor.py:97)
    outputFrames -> int16 with clipping                        (audioProcessor.py:123, wavfile.py:92-110)

PyAudio, Qt and the parameter queues are out of scope; the per-block processing
times the reference logs every 2 s (audioProcessor.py:98-102) are returned as min / max / mean.
"""
import argparse
import logging
import time
from collections import namedtuple

import numpy as np

from ..wavio import float2pcm, pcm2float  # noqa: F401  (re-exported: the reference keeps them in gccNMF/wavfile.py)
from .utils import OverlapAddProcessor

# gccNMF/realtime/config.py:46-82 (getDefaultConfig) -- the reference never reads a config file (:104-111)
DEFAULT_PARAMS = dict(
    numTDOAs=64, numTDOAHistory=128, numSpectrogramHistory=128, gccPHATNLAlpha=2.0, gccPHATNLEnabled=False,
    microphoneSeparationInMetres=0.1, targetTDOAEpsilon=5.0, targetTDOABeta=2.0, targetTDOANoiseFloor=0.0,
    localizationEnabled=True, localizationWindowSize=6,
    numChannels=2, sampleRate=16000, deviceIndex=None,
    windowSize=1024, hopSize=512, blockSize=512,
    dictionarySize=64, dictionarySizes=[64, 128, 256, 512, 1024], dictionaryType='Pretrained', numHUpdates=0)
HEADLESS_TARGET_TDOA_INDEX = 9.60          # runRealtimeGCCNMF.py:144


def getGCCNMFConfigParams(audioPath=None, dataDir=None, dictionariesW=None, **overrides):
    """config.py:107-120.  `dictionariesW` ({type: {size: W}}) replaces the CHiME pre-training when given; otherwise the
    dictionaries are loaded / pre-trained from `dataDir` (gccNMFPretraining.getDictionariesW)."""
    p = dict(DEFAULT_PARAMS)
    unknown = set(overrides) - set(p)
    if unknown:
        raise ValueError('unknown configuration options: %s' % sorted(unknown))
    p.update(overrides)
    p['audioPath'] = audioPath
    p['numFreq'] = p['windowSize'] // 2 + 1
    p['windowsPerBlock'] = p['blockSize'] // p['hopSize']
    if dictionariesW is None:
        if dataDir is None:
            raise ValueError('either dictionariesW or dataDir (with chimeTrainSet.npy or pretrainedW/) is required')
        from .gccNMFPretraining import getDictionariesW
        dictionariesW = getDictionariesW(p['windowSize'], p['dictionarySizes'], dataDir, ordered=True)
    p['dictionariesW'] = dictionariesW
    return namedtuple('ParamsDict', p.keys())(**p)


class RealtimeGCCNMFNoGUI(object):
    """File in, file out.  `processFramesFunction` (windowedSamples (2, N, windowsPerBlock) -> same shape) defaults to a
    GCCNMFProcessor built from `params` exactly as GCCNMFProcess.run builds it (gccNMFProcessor.py:66-75)."""

    def __init__(self, audioPath=None, params=None, processFramesFunction=None, device=0, **overrides):
        self.params = params if params is not None else getGCCNMFConfigParams(audioPath, **overrides)
        p = self.params
        self.inputFrames = np.zeros((p.numChannels, p.blockSize), np.float32)     # the reference's shared arrays (runRealtimeGCCNMF.py:64-72)
        self.outputFrames = np.zeros((p.numChannels, p.blockSize), np.float32)
        self.oladProcessor = OverlapAddProcessor(p.numChannels, p.windowSize, p.hopSize, p.blockSize, p.windowsPerBlock,
                                                 self.inputFrames, self.outputFrames)
        self.gccNMFProcessor = None
        if processFramesFunction is None:
            from .gccNMFProcessor import GCCNMFProcessor
            from .utils import CircularBuffer
            # the sliding-window localisation reads the GCC-PHAT history (gccNMFProcessor.py:216-227)
            self.gccPHATHistory = CircularBuffer((p.numTDOAs, p.numTDOAHistory)) if p.localizationEnabled else None
            self.tdoaHistory = CircularBuffer((1, p.numTDOAHistory)) if p.localizationEnabled else None
            g = GCCNMFProcessor(p.sampleRate, p.windowSize, p.windowsPerBlock, p.dictionariesW, p.dictionaryType, p.dictionarySize,
                                p.numHUpdates, p.microphoneSeparationInMetres, p.localizationEnabled, p.localizationWindowSize,
                                gccPHATHistory=self.gccPHATHistory, tdoaHistory=self.tdoaHistory, device=device)
            # the messages of initParams (runRealtimeGCCNMF.py:141-161)
            g.setTargetTDOARange(HEADLESS_TARGET_TDOA_INDEX, p.targetTDOAEpsilon, p.targetTDOABeta, p.targetTDOANoiseFloor)
            g.numTDOAs = p.numTDOAs
            g.separationEnabled = True
            g.reset()
            self.gccNMFProcessor = g
            processFramesFunction = g.processFrames
        self.processFramesFunction = processFramesFunction
        self.processingTimes = []

    @property
    def latencySamples(self):
        """The overlap-add ring emits block [-3B:-2B] (utils.py:116): the output lags the input by two blocks."""
        return 2 * self.params.blockSize

    def processBlock(self, block):
        """block (numChannels, blockSize) float32 -> the next output block (a view that is overwritten by the next call)."""
        startTime = time.time()
        self.inputFrames[:] = block
        self.oladProcessor.processFrames(self.processFramesFunction)
        self.processingTimes.append(time.time() - startTime)
        return self.outputFrames

    def processSamples(self, samples, flush=True):
        """samples (numChannels, n) float32 -> (numChannels, n_out) float32: whole blocks of the file, then (flush) two
        silent blocks so that the tail leaves the overlap-add ring; n_out = (blocks [+ 2]) * blockSize."""
        p = self.params
        samples = np.asarray(samples, dtype=np.float32)
        if samples.ndim != 2 or samples.shape[0] != p.numChannels:
            raise ValueError('expected (%d, n) samples, got %s' % (p.numChannels, samples.shape))
        B = p.blockSize
        numBlocks = (samples.shape[1] + B - 1) // B
        total = numBlocks + (2 if flush else 0)
        padded = np.zeros((p.numChannels, total * B), np.float32)
        padded[:, :samples.shape[1]] = samples
        out = np.empty_like(padded)
        for b in range(total):
            out[:, b * B:(b + 1) * B] = self.processBlock(padded[:, b * B:(b + 1) * B])
        return out

    def processingTimeStats(self):
        """audioProcessor.py:98-102: (min, max, mean) seconds per block."""
        t = np.asarray(self.processingTimes)
        return (float(t.min()), float(t.max()), float(t.mean())) if t.size else (0.0, 0.0, 0.0)

    def run(self, outputPath=None, alignOutput=True):
        """Reads params.audioPath (int16 stereo wav), enhances it block by block and, when `outputPath` is given, writes
        the int16 result.  alignOutput drops the two-block latency so that output sample i corresponds to input sample i."""
        from scipy.io import wavfile
        p = self.params
        sampleRate, data = wavfile.read(p.audioPath)
        if sampleRate != p.sampleRate:
            raise ValueError('sample rate of %s is %d, configured %d' % (p.audioPath, sampleRate, p.sampleRate))
        samples = pcm2float(data).T if data.dtype.kind in 'iu' else np.asarray(data, np.float32).T
        out = self.processSamples(samples, flush=True)
        if alignOutput:
            out = out[:, self.latencySamples:self.latencySamples + samples.shape[1]]
        logging.info('Processing times (min/max/avg): %f, %f, %f' % self.processingTimeStats())
        if outputPath is not None:
            wavfile.write(outputPath, sampleRate, float2pcm(np.ascontiguousarray(out.T)))
        return out


def parseArguments(argv=None):
    """config.py:122-127 plus the output path and the dictionary directory (the reference takes DATA_DIR from defs.py)."""
    parser = argparse.ArgumentParser(description='Headless real-time GCC-NMF speech enhancement (B200)')
    parser.add_argument('-i', '--input', help='input wav file path', required=True)
    parser.add_argument('-o', '--output', help='output wav file path', required=True)
    parser.add_argument('-d', '--data-dir', help='directory with chimeTrainSet.npy and / or pretrainedW/', required=True)
    parser.add_argument('--dictionary-size', type=int, default=DEFAULT_PARAMS['dictionarySize'])
    return parser.parse_args(argv)


if __name__ == '__main__':
    logging.getLogger().setLevel(logging.INFO)
    args = parseArguments()
    RealtimeGCCNMFNoGUI(args.input, dataDir=args.data_dir, dictionarySize=args.dictionary_size,
                        dictionarySizes=[args.dictionary_size]).run(args.output)
