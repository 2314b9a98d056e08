"""Loader for tests/golden/ref_golden.npz (see tests/golden/make_golden.py)."""
import os

import numpy as np

from afec_b200 import layout

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_golden.npz")


def load():
    z = np.load(PATH)
    out = []
    for name in z["names"]:
        name = str(name)
        rate, hop, F, Fr = [int(v) for v in z[name + "/meta"]]
        hdr = b"AFXD" + np.array([0, F, Fr], dtype="<i4").tobytes()
        rec, _ = layout.parse_record(memoryview(hdr + z[name + "/record"].astype("<f8").tobytes()), 0)
        out.append(dict(name=name, pcm=z[name + "/pcm"], rate=rate, hop=hop, ref=rec))
    return out
