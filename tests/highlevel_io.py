"""Loader for tests/golden/ref_golden_highlevel.npz (tests/golden/make_golden_highlevel.py) and the comparison of two
high-level results under the parity tolerance."""
import os

import numpy as np

import parity
from afec_b200 import layout

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_golden_highlevel.npz")


def load():
    z = np.load(PATH)
    out = []
    for name in z["names"]:
        name = str(name)
        rate, hop, F, Fr = [int(v) for v in z[name + "/meta"]]
        hl = layout.HighLevelResult(F=F, scalars=z[name + "/scalars"], pitch=z[name + "/pitch"], peak=z[name + "/peak"],
                                    signature=z[name + "/signature"], features=z[name + "/features"])
        out.append(dict(name=name, pcm=z[name + "/pcm"], rate=rate, hop=hop, F=F, ref=hl))
    return out


def compare(got: layout.HighLevelResult, want: layout.HighLevelResult):
    """-> list of mismatch strings (1e-4 relative / 1e-6 absolute on every value; frame counts exact)."""
    errs = []
    if got.status != want.status:
        return ["status %d != %d" % (got.status, want.status)]
    if got.F != want.F:
        return ["F %d != %d" % (got.F, want.F)]
    for i, n in enumerate(layout.HL_SCALARS):
        if not parity.close(got.scalars[i], want.scalars[i]):
            errs.append("%s: %r != %r" % (n, got.scalars[i], want.scalars[i]))
    for n, a, b in (("pitch", got.pitch, want.pitch), ("peak", got.peak, want.peak), ("spectrum_signature", got.signature, want.signature),
                    ("classification features", got.features, want.features)):
        a, b = np.asarray(a), np.asarray(b)
        if a.shape != b.shape:
            errs.append("%s: shape %s != %s" % (n, a.shape, b.shape))
            continue
        bad = ~parity.close(a, b)
        if bad.any():
            i = np.argwhere(bad)[0]
            errs.append("%s: %d values differ, first at %s: %r != %r" % (n, int(bad.sum()), i.tolist(), a[tuple(i)], b[tuple(i)]))
    return errs
