"""The oracle's restatement of the model-free high-level derivations and the classification feature vector
(oracle/afec_oracle.c afxo_highlevel; SampleAnalyser.cpp:1232-1606, SampleClassificationDescriptors.cpp:404-560) against
vectors produced by the unmodified reference (tests/golden/make_golden_highlevel.py), end to end from the PCM."""
import numpy as np
import pytest

import highlevel_io
from afec_b200 import layout

CASES = highlevel_io.load()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_highlevel_matches_reference_golden(oracle_lib, case):
    pcm = case["pcm"]
    ll = oracle_lib.analyze(pcm, src_rate=case["rate"], hop=case["hop"], file_size=44 + pcm.size * 2)
    _, _, pk, rms = oracle_lib.condition(pcm, src_rate=case["rate"])
    got = oracle_lib.highlevel(ll, pk, rms)
    errs = highlevel_io.compare(got, case["ref"])
    assert not errs, "\n".join(errs[:20])


def test_feature_vector_shape_and_padding(oracle_lib):
    """1680 = 35 rows of the 48-entry time series; frames past the end of a short file take the silent sample's values."""
    ll = oracle_lib.analyze(CASES[0]["pcm"], src_rate=CASES[0]["rate"], hop=1024)
    hl = oracle_lib.highlevel(ll, 0.5, 0.1)
    assert hl.features.shape == (layout.HL_N_FEATURES,) and layout.HL_N_FEATURES % 48 == 0
    pad = oracle_lib.silence_pad()
    assert pad[15] == 1.0 and pad[17] == -1.0 and np.count_nonzero(pad) == 2        # flatness 1, contrast -1, the rest 0
    F = ll.F
    sig = hl.features[:14 * 48].reshape(14, 48)
    assert F < 44 and np.all(sig[:, F:] == 0.0) and np.all(sig[:, :F] > 0.0)
