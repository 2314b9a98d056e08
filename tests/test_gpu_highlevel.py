"""The high-level stage of the CUDA path (AFX_FEAT_HIGHLEVEL, afx_highlevel.cu) through the C ABI: against the reference's
own vectors end to end, and against the oracle's restatement applied to the low-level values the CUDA path itself produced
(which isolates the stage from tolerance-sized differences of its inputs)."""
import numpy as np
import pytest

import highlevel_io
from afec_b200 import api, layout, synth

pytestmark = pytest.mark.gpu

CASES = highlevel_io.load()
FEATS = api.FEAT_ALL | api.FEAT_HIGHLEVEL


@pytest.fixture(scope="module")
def analyser():
    an = api.SampleAnalyser(44100, 2048, 1024, features=FEATS)
    yield an
    an.close()


def test_highlevel_matches_reference_golden(analyser):
    pcms = [c["pcm"] for c in CASES]
    b = analyser.batch(pcms, [c["rate"] for c in CASES]).run()
    for i, c in enumerate(CASES):
        errs = highlevel_io.compare(b.highlevel(i), c["ref"])
        assert not errs, c["name"] + ":\n" + "\n".join(errs[:20])
    b.free()


@pytest.mark.parametrize("hop", [1024, 512])
def test_highlevel_stage_vs_oracle_on_its_own_inputs(oracle_lib, hop):
    pcms = [synth.one_shot(1400 + i, 0.2 + 0.7 * i) for i in range(8)]
    pcms += [synth.one_shot(1410, 21.5), synth.one_shot(1411, 0.02), np.zeros(30000, dtype=np.int16),
             synth.one_shot(1412, 1.5, channels=2), np.zeros((0,), dtype=np.int16)]
    t = np.arange(44100 * 3) / 44100.0
    pcms.append(np.round(9000 * np.sin(2 * np.pi * 330.0 * t)).astype(np.int16))          # steady tone: the high-confidence branch
    an = api.SampleAnalyser(44100, 2048, hop, features=FEATS)
    b = an.batch(pcms, [44100] * len(pcms)).run()
    for i, p in enumerate(pcms):
        ll = b.result(i)
        got = b.highlevel(i)
        if ll.status != 0:
            assert got.status == ll.status
            continue
        hdr = layout.HEADER_NAMES
        want = oracle_lib.highlevel(ll, ll.header[hdr.index("peak_value")], ll.header[hdr.index("rms_value")])
        errs = highlevel_io.compare(got, want)
        assert not errs, "file %d:\n" % i + "\n".join(errs[:20])
    b.free()
    an.close()


def test_highlevel_needs_the_full_low_level_set():
    with pytest.raises(api.AfxError):
        api.SampleAnalyser(44100, 2048, 1024, features=api.FEAT_SPECTRAL | api.FEAT_HIGHLEVEL)


def test_low_level_results_do_not_change_with_the_highlevel_stage(analyser):
    pcms = [synth.one_shot(1420 + i, 0.4 + 0.5 * i) for i in range(4)]
    plain = api.SampleAnalyser(44100, 2048, 1024, features=api.FEAT_ALL)
    a = plain.analyze_pcm(pcms, [44100] * 4)
    plain.close()
    g = analyser.analyze_pcm(pcms, [44100] * 4)
    for x, y in zip(a, g):
        assert np.array_equal(x.header, y.header) and np.array_equal(x.stats, y.stats)
        for u, v in zip(x.fs + x.fv, y.fs + y.fv):
            assert np.array_equal(u, v)
