"""The CPU oracle against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  This is what pins the oracle: the reference's own tests hold
no descriptor values."""
import pytest

import golden_io
import parity

CASES = golden_io.load()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_reference_golden(oracle_lib, case):
    pcm = case["pcm"]
    got = oracle_lib.analyze(pcm, src_rate=case["rate"], hop=case["hop"], file_size=44 + pcm.size * 2)
    errs = parity.compare(got, case["ref"])
    assert not errs, "\n".join(errs[:20])


def test_rejects_bad_input(oracle_lib):
    """SampleAnalyser.cpp:472-482: channels outside 1..8 and empty files are load errors."""
    import numpy as np
    assert oracle_lib.analyze(np.zeros((16, 9), dtype=np.int16)).status != 0
    assert oracle_lib.analyze(np.zeros((0,), dtype=np.int16)).status != 0
