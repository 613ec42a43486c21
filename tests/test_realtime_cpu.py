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


def test_headless_runner_block_plumbing(tmp_path):
    """gcc-nmf_b200/realtime/runRealtimeGCCNMF.py with an injected frame function: wav in -> blocks ->
    overlap-add ring -> wav out equals the oracle restatement of the ring driven with the same blocks."""
    from scipy.io import wavfile
    from gcc_nmf_b200.realtime.runRealtimeGCCNMF import RealtimeGCCNMFNoGUI, float2pcm, getGCCNMFConfigParams, pcm2float
    rng = np.random.default_rng(3)
    N, hop, B, n = 64, 16, 32, 1000
    x16 = float2pcm((0.3 * rng.standard_normal((n, 2))).astype(np.float32))
    src, dst = str(tmp_path / 'in.wav'), str(tmp_path / 'out.wav')
    wavfile.write(src, 16000, x16)
    window = np.sqrt(np.hamming(N)).astype(np.float32)[:, None]
    fn = lambda w: (w * window) * window      # noqa: E731
    params = getGCCNMFConfigParams(src, dictionariesW={}, windowSize=N, hopSize=hop, blockSize=B)
    assert params.windowsPerBlock == B // hop and params.numFreq == N // 2 + 1 and params.numTDOAs == 64
    runner = RealtimeGCCNMFNoGUI(params=params, processFramesFunction=fn)
    out = runner.run(dst, alignOutput=False)
    ref = orc.OverlapAddProcessorOracle(2, N, hop, B, B // hop)
    x = pcm2float(x16).T
    blocks = (n + B - 1) // B + 2
    padded = np.zeros((2, blocks * B), np.float32)
    padded[:, :n] = x
    expect = np.concatenate([ref.processFrames(padded[:, b * B:(b + 1) * B].copy(), fn) for b in range(blocks)], axis=1)
    assert out.shape == expect.shape and np.array_equal(out, expect)
    rate, written = wavfile.read(dst)
    assert rate == 16000 and np.array_equal(written, float2pcm(out.T))
    # sqrt-hamming analysis x synthesis at 75 % overlap sums to a (nearly: symmetric window) constant, so once the two-block
    # latency is removed the output is the input times that gain
    aligned = out[:, runner.latencySamples:runner.latencySamples + n]
    gain = float(np.sum(np.hamming(N)[::hop]))
    assert np.allclose(aligned[:, N:n - N], gain * x[:, N:n - N], rtol=0.03, atol=2e-3)
    assert len(runner.processingTimes) == blocks and runner.processingTimeStats()[1] >= runner.processingTimeStats()[0]
