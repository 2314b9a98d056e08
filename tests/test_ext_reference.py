"""Sanity of the extension's CPU restatement itself (tests/ext_reference.py): filter bank shape, a pure tone lands in the
right mel filter / pitch class, silence.  CPU only."""
import numpy as np

import ext_reference as ext


def test_weight_tables():
    wm, wc = ext.weights()
    assert wm.shape == (40, 1024) and wc.shape == (12, 1024)
    assert np.all(wm >= 0) and np.all(wm <= 1) and np.all(wm.max(axis=1) == 1.0)          # equal gain: every filter peaks at 1
    peaks = wm.argmax(axis=1)
    assert np.all(np.diff(peaks) > 0) and peaks[0] >= 1 and peaks[-1] < 738
    used = (wc.sum(axis=0) > 0)
    assert np.allclose(wc.sum(axis=0)[used], 1.0, atol=1e-6)                              # a bin's weight is shared, not multiplied
    assert not used[:3].any() and not used[400:].any()                                    # below C2 / above C9


def test_pure_tone_lands_on_its_pitch_class():
    t = np.arange(44100) / 44100.0
    for f, cls in ((440.0, 9), (261.6256, 0), (1318.51, 4)):                              # A4, C4, E6
        x = 0.5 * np.sin(2 * np.pi * f * t)
        mfcc, chroma, idx, E, C = ext.analyze(x, 1024)
        assert np.all(idx[2:-2] == cls), (f, idx)
        assert np.all(chroma.max(axis=1)[2:-2] == 1.0) and mfcc.shape == (len(idx), 13)


def test_silence():
    mfcc, chroma, idx, E, C = ext.analyze(np.zeros(4096), 1024)
    assert np.all(chroma == 0) and np.all(idx == 0)
    assert np.allclose(mfcc[:, 0], 40 * np.log(2e-42)) and np.allclose(mfcc[:, 1:], 0.0, atol=1e-9)
