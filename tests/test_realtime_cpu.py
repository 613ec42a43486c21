"""CPU: host-side real-time helpers (gcc-nmf_b200/realtime/utils.py) against the oracle restatement of
gccNMF/realtime/utils.py, and the asymmetric windows against the notebook fixtures."""
import numpy as np

from oracle import gccnmf_oracle as orc


def test_overlap_add_processor_matches_restatement():
    from gcc_nmf_b200.realtime.utils import OverlapAddProcessor
    rng = np.random.default_rng(0)
    C, N, hop, B = 2, 64, 16, 32
    wpb = B // hop
    inputFrames = np.zeros((C, B), np.float32)
    outputFrames = np.zeros((C, B), np.float32)
    ours = OverlapAddProcessor(C, N, hop, B, wpb, inputFrames, outputFrames)
    ref = orc.OverlapAddProcessorOracle(C, N, hop, B, wpb)
    window = np.sqrt(np.hamming(N)).astype(np.float32)[:, None]
    fn = lambda w: (w * window) * window      # noqa: E731
    for _ in range(20):
        inputFrames[:] = rng.standard_normal((C, B)).astype(np.float32)
        ours.processFrames(fn)
        expect = ref.processFrames(inputFrames.copy(), fn)
        assert np.array_equal(outputFrames, expect)


def test_circular_buffer_wraps_like_the_reference():
    from gcc_nmf_b200.realtime.utils import CircularBuffer
    buf = CircularBuffer((3, 5))
    for i in range(7):
        buf.set(np.full((3, 2), float(i)))
    un = buf.getUnraveledArray()
    assert un.shape == (3, 5)
    assert np.array_equal(un[0], [4., 5., 5., 6., 6.])
    assert buf.get()[0] == 6.


def test_asymmetric_windows(golden):
    from gcc_nmf_b200.online import getAsymmetricAnalysisWindow, getAsymmetricSynthesisWindow
    g = golden('lowlatency_mini')
    N, synth = int(g['params'][1]), int(g['params'][5])
    assert np.array_equal(getAsymmetricAnalysisWindow(N, synth // 2, 0), g['analysisWindow'])
    assert np.array_equal(getAsymmetricSynthesisWindow(N, synth // 2, 0), g['synthesisWindow'])
