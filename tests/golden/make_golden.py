"""Generates tests/golden/ref_golden.npz by running the UNMODIFIED reference
(oracle/_ref/afec_ref, built by oracle/build_ref.sh from /root/reference) on small seeded
synthetic PCM.  Run in the build container only (needs /root/reference to build _ref):

    python tests/golden/make_golden.py

The npz holds, per case, the int16 PCM, sample rate, hop and the reference's raw AFXD record.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from afec_b200 import synth  # noqa: E402
from oracle import oracle  # noqa: E402


def cases():
    out = []
    out.append(("mono_0.6s_h1024", synth.one_shot(1, 0.6), 44100, 1024))
    out.append(("mono_0.6s_h512", synth.one_shot(2, 0.6), 44100, 512))
    out.append(("mono_1.2s_h1024", synth.one_shot(3, 1.2), 44100, 1024))
    out.append(("stereo_0.5s_h1024", synth.one_shot(4, 0.5, channels=2), 44100, 1024))
    out.append(("stereo96k_0.4s_h1024", synth.one_shot(5, 0.4, rate=96000, channels=2), 96000, 1024))
    out.append(("mono22k_0.4s_h1024", synth.one_shot(6, 0.4, rate=22050), 22050, 1024))
    out.append(("tiny_0.02s_h1024", synth.one_shot(7, 0.02), 44100, 1024))
    out.append(("silence_0.3s_h1024", np.zeros(13230, dtype=np.int16), 44100, 1024))
    # 4.5 s of clicks at 150 bpm: enough onsets for the beat tracker branch
    x = np.zeros(int(44100 * 4.5))
    for k in range(11):
        p = int(k * 0.4 * 44100)
        rng = np.random.default_rng(k)
        x[p:p + 1500] += rng.standard_normal(1500) * np.exp(-np.arange(1500) / 250.0)
    x += np.random.default_rng(99).standard_normal(x.size) * 0.003
    out.append(("clicks_4.5s_h1024", np.round(x / np.abs(x).max() * 25000).astype(np.int16), 44100, 1024))
    return out


def main():
    assert oracle.have_reference(), "build oracle/_ref first (oracle/build_ref.sh)"
    store = {}
    names = []
    for name, pcm, rate, hop in cases():
        ref = oracle.reference_analyze([pcm], [rate], hop=hop)[0]
        assert ref.status == 0
        body = np.concatenate([ref.header] + [np.ravel(a) for a in ref.fs] + [np.ravel(a) for a in ref.fv]
                              + [np.ravel(ref.stats)])
        store[name + "/pcm"] = pcm
        store[name + "/meta"] = np.array([rate, hop, ref.F, ref.Fr], dtype=np.int64)
        store[name + "/record"] = body
        names.append(name)
    store["names"] = np.array(names)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_golden.npz"), **store)
    print("wrote", len(names), "cases")


if __name__ == "__main__":
    main()
