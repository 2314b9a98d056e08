"""Column-by-column comparison of two afec-ll.db rows (tests only): types and integers exact, REALs and
msgpack BLOBs within the parity tolerance (1e-4 relative / 1e-6 absolute), BLOB byte lengths equal."""
import sqlite3

import msgpack
import numpy as np

import parity

# index-weighted / ill-conditioned statistics are compared with parity.compare's exclusions on the arrays
# themselves (tests/test_gpu_parity.py); here they get the plain tolerance unless listed
LOOSE_SUFFIXES = ("_centroid", "_spread", "_skewness", "_kurtosis", "_flatness")


def rows(path):
    c = sqlite3.connect(path)
    c.row_factory = sqlite3.Row
    out = {}
    for r in c.execute("select * from assets"):
        out[r["filename"].split("/")[-1]] = dict(r)
    sql = c.execute("select sql from sqlite_master where name='assets'").fetchone()[0]
    pragmas = {k: c.execute("pragma " + k).fetchone()[0] for k in ("user_version", "encoding", "journal_mode")}
    c.close()
    return out, sql, pragmas


def compare_row(got: dict, want: dict, skip=("filename", "modtime")):
    errs = []
    for k, w in want.items():
        if k in skip:
            continue
        g = got[k]
        if type(g) is not type(w):
            errs.append("%s: type %s != %s" % (k, type(g).__name__, type(w).__name__))
            continue
        if w is None or isinstance(w, (str, int)):
            if g != w:
                errs.append("%s: %r != %r" % (k, g, w))
        elif isinstance(w, float):
            loose = any(k.endswith(s + "_R") for s in LOOSE_SUFFIXES)
            if not parity.close(g, w) and not (loose and abs(g - w) <= 1e-3 * max(1.0, abs(w))):
                errs.append("%s: %r != %r" % (k, g, w))
        else:
            if len(g) != len(w):
                errs.append("%s: blob length %d != %d" % (k, len(g), len(w)))
                continue
            a = np.array(msgpack.unpackb(g), dtype=np.float64)
            b = np.array(msgpack.unpackb(w), dtype=np.float64)
            ok = parity.close(a, b)
            if any(k.endswith(s + "_VR") for s in LOOSE_SUFFIXES):
                ok |= np.abs(a - b) <= 1e-3 * np.maximum(1.0, np.abs(b))
            if not ok.all():
                i = np.argwhere(~ok)[0]
                errs.append("%s: %d values differ, first at %s: %r != %r" % (k, int((~ok).sum()), i.tolist(), a[tuple(i)], b[tuple(i)]))
    return errs
