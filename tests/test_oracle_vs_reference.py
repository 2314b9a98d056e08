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


def test_sample_formats_as_the_reference_decodes_them(tmp_path):
    """WAV 8 / 16 / 24 / 32-bit, float 32 / 64 and AIFF / AIFC files through the reference's own decoders: its low-level
    values must equal the oracle's on the float32 samples tests/audio_files.py derives for each format -- this pins both the
    test writers and the restated sample conversion (SampleConverter.h:392-518) that the device-side conversion is checked against."""
    import os
    import subprocess
    import audio_files
    from afec_b200 import layout
    pcm = synth.one_shot(61, 0.45, rate=48000, channels=2)
    names, expect = [], []
    for name, (writer, kind, kw, _) in sorted(audio_files.format_cases().items()):
        values = audio_files.quantise(pcm, kind)
        path = str(tmp_path / name)
        writer(path, values, kind, 48000, **kw)
        names.append(path)
        expect.append((audio_files.to_float16range(values, kind), os.path.getsize(path),
                       {"u8": 8, "i8": 8, "i16": 16, "i24": 24, "i32": 32, "f32": 32, "f64": 64}[kind]))
    out = str(tmp_path / "dump.bin")
    subprocess.run([oracle.REF_BIN, "dump", "1024", out] + names, check=True, env=dict(os.environ, HOME=str(tmp_path)),
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    refs = layout.load_dump(out)
    assert len(refs) == len(names)
    for path, ref, (x, size, bits) in zip(names, refs, expect):
        assert ref.status == 0, path
        got = oracle.analyze(x, src_rate=48000, hop=1024, file_size=size, bit_depth=bits)
        errs = parity.compare(got, ref)
        assert not errs, os.path.basename(path) + ":\n" + "\n".join(errs[:10])
