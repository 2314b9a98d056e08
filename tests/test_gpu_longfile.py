"""Long files conditioned in parts (BASELINE config 5) on the GPU, through the C ABI: the result of the three-phase
part protocol must equal afx_analyze() on the whole file bit for bit (the RMS, a sum of squares added in a
different order, to one float32 ulp), for resampled stereo, 44.1 kHz mono and degenerate inputs, at several part counts.
The whole-file path itself is checked against the oracle in test_gpu_parity.py."""
import numpy as np
import pytest

from afec_b200 import api, layout, longfile, synth

pytestmark = pytest.mark.gpu

@pytest.fixture(scope="module")
def an():
    a = api.SampleAnalyser(44100, 2048, 1024, features=api.FEAT_ALL)
    yield a
    a.close()


def same(a: layout.FileResult, b: layout.FileResult):
    assert (a.status, a.F, a.Fr) == (b.status, b.F, b.Fr)
    ha, hb = a.header.copy(), b.header.copy()
    rms_slots = [i for i in range(len(ha)) if ha[i] != hb[i]]
    for i in rms_slots:                                        # only the RMS slot (a float32) may differ, by one ulp
        assert abs(ha[i] - hb[i]) <= 2e-7 * max(1e-30, abs(hb[i])), (i, ha[i], hb[i])
    assert len(rms_slots) <= 1
    for x, y in zip(a.fs, b.fs):
        assert np.array_equal(x, y)
    for x, y in zip(a.fv, b.fv):
        assert np.array_equal(x, y)
    assert np.array_equal(a.stats, b.stats)


def cases():
    clip = synth.one_shot(31, 8.0, rate=96000, channels=2)
    yield "96k-stereo-32s", np.ascontiguousarray(np.tile(clip, (4, 1))), 96000
    m = synth.one_shot(5, 9.0)
    pad = np.zeros(50000, dtype=np.int16)
    yield "44k-mono-padded", np.concatenate([pad, m, m[::-1].copy(), pad, m, pad]), 44100
    yield "22k-mono", synth.one_shot(9, 4.0, rate=22050), 22050
    yield "silent", np.zeros(200000, dtype=np.int16), 44100
    yield "48k-short", synth.one_shot(11, 0.4, rate=48000), 48000


@pytest.mark.parametrize("name,pcm,rate", list(cases()), ids=[c[0] for c in cases()])
def test_parts_equal_whole_file(an, name, pcm, rate):
    whole = an.batch([pcm], [rate]).run()
    want = whole.result(0)
    whole.free()
    for n_parts in (1, 2, 5):
        b = longfile.analyze_in_parts([an], pcm, rate, n_parts=n_parts)
        got = b.result(0)
        b.free()
        same(got, want)


def test_part_window_is_the_conditioned_signal(an):
    """The window handed to the analysis, scaled, is mData (SampleAnalyser.cpp:712-718) -- checked through the
    debug read-back of the batch made by afx_analyze_conditioned."""
    clip = synth.one_shot(31, 8.0, rate=96000, channels=2)
    pcm = np.ascontiguousarray(np.tile(clip, (2, 1)))            # 16 s: shorter than the cap, so the window is all of mData
    whole = an.batch([pcm], [96000]).run()
    want = whole.conditioned(0)
    whole.free()
    b = longfile.analyze_in_parts([an], pcm, 96000, n_parts=3)
    got = b.conditioned(0)
    b.free()
    assert len(got) == len(want) < 882000 and np.array_equal(got, want)
