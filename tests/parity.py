"""Shared comparison rules for descriptor parity (tests only).

Tolerance is the one BASELINE.json's north_star states: 1e-4 relative / 1e-6 absolute per
float descriptor; integer-valued outputs (frame counts, rolloff counts, complexity counts,
silence flags, onset counts) bit-exact.

Two documented exclusions, both ill-conditioned *by construction* in the reference itself
(they amplify 1-ulp FFT differences without bound, so even the reference built with IPP
instead of Ooura would disagree with itself):
  * index-weighted statistics (centroid/spread/skewness/kurtosis) and flatness (gmean/mean)
    of a series whose sum cancels (sum|x| / |sum x| > 1e6): Statistics.cpp:459-574 divide by
    that sum;
  * skewness/kurtosis whose spread is within 1e-6 of the 1e-12 cut-off.
"""
from __future__ import annotations

import numpy as np

from afec_b200 import layout

RTOL = 1e-4
ATOL = 1e-6


def close(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    with np.errstate(invalid="ignore"):
        ok = np.abs(a - b) <= ATOL + RTOL * np.abs(b)
    return ok | both_nan | ((a == b))


def _series_iter(r: layout.FileResult):
    for i, n in enumerate(layout.FRAMED_SCALARS):
        yield n, r.fs[i]
    for i, (n, nb) in enumerate(layout.FRAMED_VECTORS):
        for b in range(nb):
            yield "%s[%d]" % (n, b), r.fv[i][:, b]


def compare(got: layout.FileResult, want: layout.FileResult, skip_series=(), only_series=None,
            check_stats=True, check_header=True, max_flip_frac=0.0):
    """Return a list of human-readable mismatch strings (empty == parity)."""
    errs = []
    if got.status != want.status:
        return ["status %d != %d" % (got.status, want.status)]
    if want.status != 0:
        return errs
    if (got.F, got.Fr) != (want.F, want.Fr):
        return ["frame counts (%d, %d) != (%d, %d)" % (got.F, got.Fr, want.F, want.Fr)]
    if check_header and only_series is None:
        for i, n in enumerate(layout.HEADER_NAMES[:23]):
            if not close(got.header[i], want.header[i]):
                errs.append("header %s: %r != %r" % (n, got.header[i], want.header[i]))
    names = list(layout.FRAMED_SCALARS) + [n for n, _ in layout.FRAMED_VECTORS]
    for n in names:
        if n in skip_series or (only_series is not None and n not in only_series):
            continue
        a, b = got.series(n), want.series(n)
        if n in layout.INTEGER_SERIES:
            bad = (a != b)
        else:
            bad = ~close(a, b)
        nb = int(bad.sum())
        if nb > max_flip_frac * bad.size:
            idx = np.argwhere(bad)[0]
            errs.append("%s: %d/%d values differ, first at %s: %r != %r" % (
                n, nb, bad.size, idx.tolist(), a[tuple(idx)], b[tuple(idx)]))
    if check_stats:
        for si, (n, x) in enumerate(_series_iter(want)):
            base = n.split("[")[0]
            if base in skip_series or (only_series is not None and base not in only_series):
                continue
            a, b = got.stats[si], want.stats[si]
            ok = close(a, b)
            sx = np.sum(np.abs(x))
            cancel = sx > 0 and abs(np.sum(x)) * 1e6 < sx
            if cancel:
                ok[6:11] = True          # centroid..kurtosis and flatness (= gmean / mean)
            if abs(abs(b[7]) - 1e-12) < 1e-6 * 1e-12 or abs(b[7]) < 1e-9:
                ok[8:10] = True
            if not ok.all():
                k = int(np.argwhere(~ok)[0][0])
                errs.append("stat %s_%s: %r != %r" % (n, layout.STAT_NAMES[k], a[k], b[k]))
    return errs
