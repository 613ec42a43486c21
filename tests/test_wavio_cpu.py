"""wav conventions of gccNMF/wavfile.py (clip protection, float2pcm / pcm2float) against files written by the reference
itself (tests/golden/wavfile_mini.npz, oracle/make_golden.py:golden_wavfile)."""
import os

import numpy as np


def test_wavwrite_and_wavread_match_reference_bytes(golden, tmp_path):
    from gcc_nmf_b200 import wavio
    g = golden('wavfile_mini')
    for name in ('quiet', 'loud'):                      # loud: peak >= 1 -> rescaled to 0.99 (wavfile.py:41-44)
        path = str(tmp_path / (name + '.wav'))
        wavio.wavwrite(g[name], path, 16000)
        assert np.array_equal(np.frombuffer(open(path, 'rb').read(), dtype=np.uint8), g[name + '_bytes'])
        x, sr = wavio.wavread(path)
        assert sr == 16000 and x.dtype == np.float32 and np.array_equal(x, g[name + '_read'])
    assert np.array_equal(wavio.pcm2float(g['u8']), g['u8_float'])      # unsigned PCM is re-centred (wavfile.py:88-90)


def test_drop_in_save_and_load_use_the_reference_mapping(golden, tmp_path):
    import gcc_nmf_b200.gccNMFFunctions as fn       # imports without a GPU (the library is only loaded on first use)
    g = golden('wavfile_mini')
    prefix = str(tmp_path / 'clip')
    fn.saveTargetSignalEstimates(np.stack([g['loud'], g['quiet']]), 16000, prefix)
    for i, name in enumerate(('loud', 'quiet')):
        path = fn.getSourceEstimateFileName(prefix, i)
        assert os.path.basename(path) == 'clip_sim_%d.wav' % (i + 1)
        assert np.array_equal(np.frombuffer(open(path, 'rb').read(), dtype=np.uint8), g[name + '_bytes'])
        x, sr = fn.loadMixtureSignal(path)
        assert np.array_equal(x, g[name + '_read']) and x.flags['C_CONTIGUOUS']
