"""GPU: online / low-latency enhancement (row a11) against fixtures produced by exec()-ing the reference's notebook
cells, the real-time chunk processor (row a13) against its (unpinned) numpy restatement, and the pre-training call site."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import gccnmf_oracle as orc  # noqa: E402


def relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300)


def _check_online(res, g, prefix):
    X, Y, out, accMax, targets, ang, masks, wiener = res
    assert ang.shape == g[prefix + 'angularSpectrogram'].shape if prefix + 'angularSpectrogram' in g else True
    tgt_ref = g[prefix + 'targetTDOAs']
    assert targets.shape == tgt_ref.shape
    agree_t = np.mean(targets == tgt_ref)
    agree_m = np.mean(masks == g[prefix + 'atomMasks'])
    # complex64 spectra instead of the notebook's complex128: decisions may differ only on float32-level near-ties
    assert agree_t > 0.99 and agree_m > 0.995, (agree_t, agree_m)
    if agree_t == 1.0 and agree_m == 1.0:
        assert relerr(wiener, g[prefix + 'wienerFilters']) < 1e-5
        assert relerr(out, g[prefix + 'output']) < 1e-4
    return agree_t, agree_m


def test_online_loop_matches_notebook(golden):
    from gcc_nmf_b200.online import performOnlineSpeechEnhancement
    g = golden('online_mini')
    sr, N, hop, D, K = [int(v) for v in g['params']]
    res = performOnlineSpeechEnhancement(g['samples'], sr, g['W'], np.hanning(N), np.hanning(N), hop, D, float(g['micSep']),
                                         0.05 * D, gainPerFrame=True)
    np.testing.assert_allclose(res[5], g['angularSpectrogram'], rtol=0, atol=2e-4)
    t, m = _check_online(res, g, '')
    print('online: target agreement %.4f, atom-mask agreement %.5f' % (t, m))


def test_low_latency_loops_match_notebook(golden):
    from gcc_nmf_b200.online import performOnlineSpeechEnhancement, getAsymmetricAnalysisWindow, getAsymmetricSynthesisWindow
    g = golden('lowlatency_mini')
    sr, N, hop, D, K, synth = [int(v) for v in g['params']]
    assert np.array_equal(getAsymmetricAnalysisWindow(N, synth // 2, 0), g['analysisWindow'])
    assert np.array_equal(getAsymmetricSynthesisWindow(N, synth // 2, 0), g['synthesisWindow'])
    for tag, win, syn in (('sym_', g['symmetricWindow'], g['symmetricWindow']), ('asym_', g['analysisWindow'], g['synthesisWindow'])):
        res = performOnlineSpeechEnhancement(g['samples'], sr, g['W'], win, syn, hop, D, float(g['micSep']), 0.05 * D, gainPerFrame=False)
        t, m = _check_online(res, g, tag)
        print('low-latency %s target agreement %.4f, atom-mask agreement %.5f' % (tag, t, m))


@pytest.mark.parametrize('nT,mode', [(1, 1), (4, 0), (4, 1)])
def test_realtime_processor_matches_restatement(nT, mode):
    from gcc_nmf_b200.realtime.gccNMFProcessor import GCCNMFProcessor
    from gcc_nmf_b200.realtime.utils import CircularBuffer
    rng = np.random.default_rng(nT * 10 + mode)
    sr, N, K, D = 16000, 256, 64, 32
    W = (rng.random((N // 2 + 1, K)) ** 3).astype(np.float32)
    proc = GCCNMFProcessor(sr, N, nT, {'Pretrained': {K: W}}, 'Pretrained', K, 0, 0.1, True, 6,
                           gccPHATHistory=CircularBuffer((D, 128)), tdoaHistory=CircularBuffer((1, 128)))
    proc.numTDOAs = D
    proc.targetMode = mode
    proc.reset()
    proc.setTargetTDOARange(10.0, 3.0, 2.0, 0.01)
    ref = orc.GCCNMFProcessorOracle(sr, N, nT, W, D, 0.1, localizationEnabled=True, localizationWindowSize=6)
    ref.targetMode = mode
    ref.setTargetTDOARange(10.0, 3.0, 2.0, 0.01)
    worst = 0.0
    for step in range(12):
        # a delayed copy in the right channel gives a well-defined TDOA peak
        s = rng.standard_normal((N + 8, nT)).astype(np.float32)
        frames = np.stack([s[4:4 + N], 0.8 * s[2:2 + N] + 0.05 * rng.standard_normal((N, nT)).astype(np.float32)])
        y = proc.processFrames(frames)
        yr = ref.processFrames(frames)
        assert y.shape == yr.shape == (2, N, nT)
        assert float(proc.targetTDOAIndex) == float(ref.targetTDOAIndex)          # localisation decision
        worst = max(worst, relerr(y, yr))
    assert worst < 2e-3, worst      # float32 restatement vs float64-argmax device path; near-ties may differ
    print('rt processor nT=%d mode=%d: worst rel error %.2e' % (nT, mode, worst))


def test_pretraining_call_site(golden, tmp_path):
    from gcc_nmf_b200.realtime import gccNMFPretraining as pre
    g = golden('pretraining_mini')
    W = pre.loadPretrainedW(12, str(tmp_path), trainV=g['trainV'])
    Wo, _ = orc.performKLNMF(g['trainV'], 12, 100, 0, 1e-16, 0)
    assert relerr(W, Wo) < 1e-4
    assert np.array_equal(pre.loadPretrainedW(12, str(tmp_path)), W)             # second call hits the .npy cache
    assert np.array_equal(pre.getOrderedDictionary(g['W']), g['orderedW'])


def test_headless_realtime_runner(tmp_path):
    """wav file -> blocks -> GCCNMFProcessor through the overlap-add ring, against the oracle
    restatements of both (gccNMF/realtime/gccNMFProcessor.py:201-231, utils.py:99-116) driven with the same blocks."""
    from scipy.io import wavfile
    from gcc_nmf_b200.realtime.runRealtimeGCCNMF import (HEADLESS_TARGET_TDOA_INDEX, RealtimeGCCNMFNoGUI, float2pcm,
                                                         getGCCNMFConfigParams, pcm2float)
    rng = np.random.default_rng(21)
    sr, N, hop, B, K, D = 16000, 256, 128, 256, 64, 32
    n = 24 * B                              # whole blocks: no silent (0 / 0 in the PHAT normalisation) frames
    s = rng.standard_normal(n + 8).astype(np.float32)
    x = 0.2 * np.stack([s[4:4 + n], 0.8 * s[2:2 + n] + 0.05 * rng.standard_normal(n).astype(np.float32)])
    src = str(tmp_path / 'in.wav')
    wavfile.write(src, sr, float2pcm(np.ascontiguousarray(x.T)))
    W = (rng.random((N // 2 + 1, K)) ** 3).astype(np.float32)
    params = getGCCNMFConfigParams(src, dictionariesW={'Pretrained': {K: W}}, windowSize=N, hopSize=hop, blockSize=B, numTDOAs=D,
                                   dictionarySize=K, dictionarySizes=[K], sampleRate=sr)
    runner = RealtimeGCCNMFNoGUI(params=params)
    xq = pcm2float(wavfile.read(src)[1]).T
    out = runner.processSamples(xq, flush=False)
    nT = B // hop
    ref = orc.GCCNMFProcessorOracle(sr, N, nT, W, D, 0.1, localizationEnabled=True, localizationWindowSize=6)
    ref.setTargetTDOARange(HEADLESS_TARGET_TDOA_INDEX, params.targetTDOAEpsilon, params.targetTDOABeta, params.targetTDOANoiseFloor)
    ring = orc.OverlapAddProcessorOracle(2, N, hop, B, nT)
    expect = np.concatenate([ring.processFrames(xq[:, b * B:(b + 1) * B].copy(), ref.processFrames) for b in range(n // B)], axis=1)
    assert out.shape == expect.shape == (2, n)
    assert float(runner.gccNMFProcessor.targetTDOAIndex) == float(ref.targetTDOAIndex)
    assert np.isfinite(out).all()
    assert relerr(out, expect) < 5e-3          # float32 restatement vs float64-argmax device path; near-ties may differ
    assert len(runner.processingTimes) == n // B
    print('headless runner: %d blocks, rel error %.2e, processing times (min/max/avg) %s' % (n // B, relerr(out, expect), runner.processingTimeStats()))
