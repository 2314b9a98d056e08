"""Comparison of two afec-ll.db rows (tests only).  Column types, TEXT / INTEGER values and BLOB byte lengths must be
equal; the numeric content is rebuilt into layout.FileResult records and compared by parity.compare -- the same rules
(1e-4 relative / 1e-6 absolute, integer series exact, the documented ill-conditioned statistics) as every other
parity test, nothing looser."""
import sqlite3

import msgpack
import numpy as np

import parity
from afec_b200 import layout


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


def _unpack(blob):
    return np.array(msgpack.unpackb(blob), dtype=np.float64)


def row_to_result(row: dict) -> layout.FileResult:
    """The numeric columns of a succeeded row as a FileResult (header scalars that are DB columns, every series, every statistic)."""
    r = layout.FileResult()
    for i, n in enumerate(layout.HEADER_NAMES[:23]):
        r.header[i] = float(row[n + "_R"])
    for n in layout.FRAMED_SCALARS:
        r.fs.append(_unpack(row[n + "_VR"]).reshape(-1))
    for n, nb in layout.FRAMED_VECTORS:
        v = _unpack(row[n + "_VVR"])
        r.fv.append(v.reshape(-1, nb) if v.size else np.zeros((0, nb)))
    r.F = len(r.fs[0]); r.Fr = len(r.fs[layout.N_FS_MAIN])
    si = 0
    for n in layout.FRAMED_SCALARS:
        for k, st in enumerate(layout.STAT_NAMES):
            r.stats[si, k] = float(row["%s_%s_R" % (n, st)])
        si += 1
    for n, nb in layout.FRAMED_VECTORS:
        for k, st in enumerate(layout.STAT_NAMES):
            v = _unpack(row["%s_%s_VR" % (n, st)])
            assert v.shape == (nb,), (n, st, v.shape)
            r.stats[si:si + nb, k] = v
        si += nb
    return r


def compare_row(got: dict, want: dict, skip=("filename", "modtime")):
    errs = []
    for k, w in want.items():
        if k in skip:
            continue
        g = got[k]
        if type(g) is not type(w):
            errs.append("%s: type %s != %s" % (k, type(g).__name__, type(w).__name__))
        elif w is None or isinstance(w, (str, int)):
            if g != w:
                errs.append("%s: %r != %r" % (k, g, w))
        elif isinstance(w, bytes) and len(g) != len(w):
            errs.append("%s: blob length %d != %d" % (k, len(g), len(w)))
    if errs or want["status"] != "succeeded":
        return errs
    return parity.compare(row_to_result(got), row_to_result(want))
