"""Generates tests/golden/ref_golden_highlevel.npz: the high-level descriptors that need no classification model and the
classification feature vector, produced by the UNMODIFIED reference (`afec_ref dumphl`: TSampleAnalyser::Analyze with
kHighLevelDescriptors and no models loaded, then TSampleClassificationDescriptors) on the cases of make_golden.py plus a
few longer ones (more than 64 frames: the spectrum signature is then a real down-sampling; pitched material for the
base-note branch).  Run in the build container only:

    python tests/golden/make_golden_highlevel.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from afec_b200 import synth  # noqa: E402
from oracle import oracle  # noqa: E402
import make_golden  # noqa: E402


def cases():
    out = [(n, p, r) for n, p, r, h in make_golden.cases() if h == 1024]
    out.append(("mono_8s", synth.one_shot(12, 8.0), 44100))
    out.append(("mono_3.1s", synth.one_shot(15, 3.1), 44100))
    t = np.arange(int(44100 * 2.5)) / 44100.0                       # a steady tone: high pitch confidence, base note = A3
    out.append(("tone_220hz_2.5s", np.round(12000 * np.sin(2 * np.pi * 220.0 * t) * np.exp(-t / 2.0)).astype(np.int16), 44100))
    return out


def main():
    assert oracle.have_reference(), "build oracle/_ref first (oracle/build_ref.sh)"
    store, names = {}, []
    cs = cases()
    ref = oracle.reference_analyze_highlevel([c[1] for c in cs], [c[2] for c in cs], hop=1024)
    for (name, pcm, rate), (ll, hl) in zip(cs, ref):
        assert ll.status == 0 and hl.status == 0
        store[name + "/pcm"] = pcm
        store[name + "/meta"] = np.array([rate, 1024, ll.F, ll.Fr], dtype=np.int64)
        store[name + "/scalars"] = hl.scalars
        store[name + "/pitch"] = hl.pitch
        store[name + "/peak"] = hl.peak
        store[name + "/signature"] = hl.signature
        store[name + "/features"] = hl.features
        names.append(name)
    store["names"] = np.array(names)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_golden_highlevel.npz"), **store)
    print("wrote", len(names), "cases")


if __name__ == "__main__":
    main()
