"""Drop-in for gccNMF/realtime/gccNMFPretraining.py:43-87: dictionary pre-training on the CHiME magnitude matrix
(performKLNMF on the B200), the `W_%d.npy` cache and the spectral-centroid atom ordering."""
import logging
from collections import OrderedDict
from os import makedirs
from os.path import exists, join

import numpy as np

from ..gccNMFFunctions import performKLNMF

SPARSITY_ALPHA = 0
NUM_PRELEARNING_ITERATIONS = 100


def getOrderedDictionary(W):
    """:60-66: atoms sorted by spectral centroid."""
    numFreq, _ = W.shape
    spectralCentroids = np.squeeze(np.sum(np.arange(numFreq)[:, np.newaxis] * W, axis=0, keepdims=True) / np.sum(W, axis=0, keepdims=True))
    return np.squeeze(W[:, np.argsort(spectralCentroids)])


def loadPretrainedW(dictionarySize, dataDir, retrainW=False, trainV=None):
    """:68-87.  `dataDir` replaces the reference's DATA_DIR constant (defs.py:30-39); the training matrix is
    `dataDir/chimeTrainSet.npy` unless given."""
    pretrainedWDir = join(dataDir, 'pretrainedW')
    pretrainedWFilePath = join(pretrainedWDir, 'W_%d.npy' % dictionarySize)
    logging.info('GCCNMFPretraining: Loading pretrained W (size %d): %s' % (dictionarySize, pretrainedWFilePath))
    if exists(pretrainedWFilePath) and not retrainW:
        return np.load(pretrainedWFilePath)
    if trainV is None:
        trainV = np.load(join(dataDir, 'chimeTrainSet.npy'))
    W, _ = performKLNMF(trainV, dictionarySize, numIterations=NUM_PRELEARNING_ITERATIONS, sparsityAlpha=SPARSITY_ALPHA,
                        epsilon=1e-16, seedValue=0)
    try:
        makedirs(pretrainedWDir)
    except OSError:
        pass
    np.save(pretrainedWFilePath, W)
    return W


def getDictionariesW(windowSize, dictionarySizes, dataDir, ordered=False):
    """:43-58."""
    fftSize = windowSize // 2 + 1
    dictionariesW = OrderedDict([
        ('Pretrained', OrderedDict([(k, loadPretrainedW(k, dataDir)) for k in dictionarySizes])),
        ('Random', OrderedDict([(k, np.random.rand(fftSize, k).astype('float32')) for k in dictionarySizes]))])
    if not ordered:
        return dictionariesW
    return OrderedDict((t, OrderedDict((k, getOrderedDictionary(W)) for k, W in d.items())) for t, d in dictionariesW.items())
