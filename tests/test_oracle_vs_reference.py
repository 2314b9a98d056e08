"""Wider randomized check of the oracle against the live reference binary (oracle/_ref/afec_ref).
Skipped where the reference was not built (e.g. a checkout without /root/reference)."""
import numpy as np
import pytest

import parity
from afec_b200 import synth
from oracle import oracle

pytestmark = pytest.mark.skipif(not oracle.have_reference(), reason="oracle/_ref not built")


@pytest.mark.parametrize("hop", [1024, 512])
def test_mixed_corpus(hop):
    cases = [(synth.one_shot(200 + i, 0.4 + 0.35 * i), 44100) for i in range(6)]
    cases.append((synth.one_shot(300, 0.8, channels=2), 44100))
    cases.append((synth.one_shot(301, 0.6, rate=48000, channels=2), 48000))
    cases.append((synth.one_shot(302, 0.5, rate=32000), 32000))
    cases.append((synth.one_shot(303, 21.0), 44100))          # crosses the 20 s analysis cap
    refs = oracle.reference_analyze([c[0] for c in cases], [c[1] for c in cases], hop=hop)
    for (pcm, rate), ref in zip(cases, refs):
        got = oracle.analyze(pcm, src_rate=rate, hop=hop, file_size=44 + pcm.size * 2)
        errs = parity.compare(got, ref)
        assert not errs, "\n".join(errs[:20])
