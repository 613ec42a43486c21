"""Drop-ins for the numeric helpers of gccNMF/realtime/utils.py: the circular history buffer (:34-70, without
the multiprocessing shared memory) and the block overlap-add processor (:72-116)."""
import numpy as np


class CircularBuffer(object):
    """gccNMF/realtime/utils.py:34-70 (`SharedMemoryCircularBuffer`) on a plain float64 array."""

    def __init__(self, shape, initValue=0):
        self.values = np.full(shape, initValue, dtype=np.float64)
        self.numValues = self.values.shape[-1]
        self.index = 0

    def set(self, newValues, index=None):
        index = self.index if index is None else index
        newValues = np.asarray(newValues)
        n = newValues.shape[-1]
        if index + n < self.numValues:
            self.values[..., index:index + n] = newValues
            self.index = index + n
        else:
            numAtEnd = self.numValues - index
            numAtStart = n - numAtEnd
            self.values[..., index:] = newValues[..., :numAtEnd]
            self.values[..., :numAtStart] = newValues[..., numAtEnd:]
            self.index = numAtStart
        return self.index

    def get(self, index=None):
        index = (self.index - 1) % self.numValues if index is None else (index % self.numValues)
        return self.values[..., index]

    def getUnraveledArray(self):
        return np.concatenate([self.values[:, self.index:], self.values[:, :self.index]], axis=-1)

    def size(self):
        return self.values.shape[-1]


SharedMemoryCircularBuffer = CircularBuffer


class OverlapAddProcessor(object):
    """gccNMF/realtime/utils.py:72-116: 8-block input/output rings; each call shifts in one block, cuts
    `windowsPerBlock` windows, runs `processFramesFunction` on (channels, windowSize, windowsPerBlock), overlap-adds
    the result and emits block [-3B:-2B]."""

    def __init__(self, numChannels, windowSize, hopSize, blockSize, windowsPerBlock, inputFrames, outputFrames):
        self.numChannels, self.windowSize, self.hopSize = numChannels, windowSize, hopSize
        self.blockSize, self.windowsPerBlock = blockSize, windowsPerBlock
        self.inputFrames, self.outputFrames = inputFrames, outputFrames
        self.numBlocksPerBuffer = 8
        self.inputBufferSize = self.outputBufferSize = blockSize * self.numBlocksPerBuffer
        self.inputBuffer = np.zeros((numChannels, self.inputBufferSize), np.float32)
        self.outputBuffer = np.zeros((numChannels, self.outputBufferSize), np.float32)
        self.windowedSamples = np.zeros((numChannels, windowSize, windowsPerBlock), np.float32)

    def processFrames(self, processFramesFunction):
        B = self.blockSize
        self.inputBuffer[:, :-B] = self.inputBuffer[:, B:]
        self.inputBuffer[:, -B:] = self.inputFrames
        self.outputBuffer[:, :-B] = self.outputBuffer[:, B:]
        self.outputBuffer[:, -B:] = 0
        windowIndexes = np.arange(self.inputBufferSize - self.windowSize - (self.windowsPerBlock - 1) * self.hopSize,
                                  self.inputBufferSize - self.windowSize + 1, self.hopSize)
        for i, w in enumerate(windowIndexes):
            self.windowedSamples[..., i] = self.inputBuffer[:, w:w + self.windowSize]
        processedFrames = processFramesFunction(self.windowedSamples)
        for i, w in enumerate(windowIndexes):
            self.outputBuffer[:, w:w + self.windowSize] += processedFrames[..., i]
        self.outputFrames[:] = self.outputBuffer[:, -3 * B:-2 * B]
