"""Drop-in for gccNMF/runGCCNMF.py:30-54.

Same signature and stage order; differences, all deliberate and documented in SURVEY.md section 3.1:
the NMF parameters are arguments (the reference hard-codes dictionarySize=128, numIterations=100 at
:41 -- those are the defaults here), the output prefix is the function's own argument (the reference
reads an undefined local at :54), and no matplotlib import.
"""
import numpy as np

from . import gccNMFFunctions as fn
from .pipeline import GCCNMFPipeline


def runGCCNMF(mixtureFilePrefix, windowSize, hopSize, numTDOAs, microphoneSeparationInMetres, numTargets=None,
              windowFunction=np.hanning, dictionarySize=128, numIterations=100, sparsityAlpha=0, save=True, device=0):
    mixtureFileName = fn.getMixtureFileName(mixtureFilePrefix)
    stereoSamples, sampleRate = fn.loadMixtureSignal(mixtureFileName)
    pipe = GCCNMFPipeline(sampleRate, windowSize, hopSize, numTDOAs, microphoneSeparationInMetres, dictionarySize,
                          numIterations, sparsityAlpha, device=device, windowFunction=windowFunction)
    targetSignalEstimates = pipe.separate_host(stereoSamples, numTargets).numpy()
    if save:
        fn.saveTargetSignalEstimates(targetSignalEstimates, sampleRate, mixtureFilePrefix)
    return targetSignalEstimates


if __name__ == '__main__':
    import sys
    prefix = sys.argv[1] if len(sys.argv) > 1 else '../data/dev1_female3_liverec_130ms_1m'
    runGCCNMF(prefix, 1024, 128, 128, 1.0, 3, np.hanning)   # the reference's __main__ parameters (:56-77)
