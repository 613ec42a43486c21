"""Block buffering around the real-time processor (gccNMF/realtime/utils.py): the circular history buffer of the GUI
(:34-70, here a plain float64 array instead of multiprocessing shared memory) and the block overlap-add processor (:72-116).

The overlap-add rings live ON THE DEVICE whenever the function handed to `OverlapAddProcessor.processFrames` is the bound
`processFrames` of this package's GCCNMFProcessor: a block is then one H2D copy, one CUDA-graph launch (rings + FFTs + GCC-NMF
mask + overlap-add + block emit, csrc/rt.cu) and one D2H copy.  Any other callable gets a host ring with the same semantics
(in-place circular indexing; the reference moves seven blocks of memory per call instead).
"""
import numpy as np


class CircularBuffer(object):
    """History ring of gccNMF/realtime/utils.py:34-70 (`SharedMemoryCircularBuffer`): columns are written at a moving index
    that wraps; `getUnraveledArray` returns them oldest first."""

    def __init__(self, shape, initValue=0):
        self.values = np.full(shape, initValue, dtype=np.float64)
        self.numValues = self.values.shape[-1]
        self.index = 0

    def set(self, newValues, index=None):
        start = self.index if index is None else index
        newValues = np.asarray(newValues)
        count = newValues.shape[-1]
        columns = (start + np.arange(count)) % self.numValues
        self.values[..., columns] = newValues
        # a write that reaches the last column wraps the index to the start (:49-59)
        self.index = int((start + count) % self.numValues)
        return self.index

    def get(self, index=None):
        return self.values[..., (self.index - 1 if index is None else index) % self.numValues]

    def getUnraveledArray(self):
        return np.roll(self.values, -self.index, axis=-1)

    def size(self):
        return self.numValues


SharedMemoryCircularBuffer = CircularBuffer


class OverlapAddProcessor(object):
    """gccNMF/realtime/utils.py:72-116.  Each `processFrames(fn)` call pushes `inputFrames` (channels, blockSize) into an 8-block
    input ring, cuts `windowsPerBlock` windows ending at the newest sample, runs `fn` on (channels, windowSize, windowsPerBlock),
    overlap-adds the result into an 8-block output ring and writes the block that lies two blocks in the past to `outputFrames`."""
    numBlocksPerBuffer = 8                                       # utils.py:85

    def __init__(self, numChannels, windowSize, hopSize, blockSize, windowsPerBlock, inputFrames, outputFrames):
        self.numChannels, self.windowSize, self.hopSize = numChannels, windowSize, hopSize
        self.blockSize, self.windowsPerBlock = blockSize, windowsPerBlock
        self.inputFrames, self.outputFrames = inputFrames, outputFrames
        self.inputBufferSize = self.outputBufferSize = blockSize * self.numBlocksPerBuffer
        self._calls = 0
        self._hostRings = None

    # ---- device path
    @staticmethod
    def _fused_processor(processFramesFunction):
        from .gccNMFProcessor import GCCNMFProcessor
        owner = getattr(processFramesFunction, '__self__', None)
        if isinstance(owner, GCCNMFProcessor) and getattr(processFramesFunction, '__func__', None) is GCCNMFProcessor.processFrames:
            return owner
        return None

    def processFrames(self, processFramesFunction):
        processor = self._fused_processor(processFramesFunction) if self.numChannels == 2 else None
        if processor is not None and processor.numTimePerChunk == self.windowsPerBlock and processor.windowSize == self.windowSize:
            self.outputFrames[:] = processor.processBlock(self.inputFrames, self.hopSize, self.blockSize)
            return
        self._process_on_host(processFramesFunction)

    # ---- host path (any callable)
    def _process_on_host(self, processFramesFunction):
        L, B, N = self.inputBufferSize, self.blockSize, self.windowSize
        if self._hostRings is None:
            self._hostRings = (np.zeros((self.numChannels, L), np.float32), np.zeros((self.numChannels, L), np.float32))
        ring_in, ring_out = self._hostRings
        self._calls += 1
        newest = self._calls * B                                  # stream position one past the newest sample

        def span(first, count):                                   # ring indexes of stream positions first .. first + count - 1
            return (first + np.arange(count)) % L
        fresh = span(newest - B, B)
        ring_in[:, fresh] = self.inputFrames
        ring_out[:, fresh] = 0                                    # utils.py:105: the slots being recycled
        starts = [newest - N - (self.windowsPerBlock - 1 - i) * self.hopSize for i in range(self.windowsPerBlock)]   # utils.py:107
        windowed = np.stack([ring_in[:, span(s, N)] for s in starts], axis=-1).astype(np.float32)
        processed = processFramesFunction(windowed)
        for i, s in enumerate(starts):                            # frame order, float32 ring += frame (utils.py:113-114)
            idx = span(s, N)
            ring_out[:, idx] = ring_out[:, idx] + processed[..., i]
        self.outputFrames[:] = ring_out[:, span(newest - 3 * B, B)]   # utils.py:115
