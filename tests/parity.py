"""Shared comparison rules for descriptor parity (tests only).

Tolerance is the one BASELINE.json's north_star states: 1e-4 relative / 1e-6 absolute per
float descriptor; integer-valued outputs (frame counts, rolloff counts, complexity counts,
silence flags, onset counts) bit-exact.

Documented exclusions, all ill-conditioned *by construction* in the reference itself
(they amplify 1-ulp FFT differences without bound, so even the reference built with IPP
instead of Ooura would disagree with itself):
  * index-weighted statistics (centroid/spread/skewness/kurtosis) and flatness (gmean/mean)
    of a series whose sum cancels (sum|x| / |sum x| > 1e6): Statistics.cpp:459-574 divide by
    that sum;
  * skewness/kurtosis whose spread is within 1e-6 of the 1e-12 cut-off;
  * the pitch triple (f0, f0_confidence, failsafe_f0) of a frame whose first 1024 samples are silent next to the rest
    (digital silence, or the last LSB ticks of a decayed tail: less than 1e-6 of the frame's energy): aubio's yinfast
    forms yin[tau] = sq[tau] - r[tau] with r from FFTs (pitchyinfast.c:110-137).  With a silent first half sq[tau] is
    exactly 0 for small tau and r[tau] is FFT rounding noise; with a few ticks x[j] in it r[tau] = sum x[j] x[j + tau]
    vanishes wherever the signal does, sq[tau] is the same 2 sum x[j]^2 for every small tau, the normalised function is
    exactly 1 there and the arg-min fall-back (mathutils.c:250-258) picks among ties.  Either way the FFT's rounding
    (1e-16 of the FRAME's energy, i.e. >= 1e-10 of these values) decides the cumulative-mean normalised function, the
    first-dip search and the confidence (callers pass the conditioned signal so those frames can be found).  Found by
    profiles/parity_sweep.py (silent half: round 1; one tick: seed 15064, frame 230 -- the oracle with its two FFT
    variants returns 3163.9 and 4026.0 Hz, the CUDA path 2845.2; two ticks: seed 16138, frame 350);
    tests/test_fft_rounding_rules.py demonstrates both.  The same ties arise from one impulse alone in a frame (yin[tau] =
    x0^2 for every tau) and from a constant frame (yin[tau] = W c^2); instead of one structural rule per shape, a pitch frame
    that differs is put to the general test pitch_noise_sensitive(): the reference's own decision procedure, restated here,
    is run on the exact difference function with r[tau] perturbed at the level of an FFT's rounding (1e-15 of the frame's
    energy); if its f0 or confidence moves beyond the tolerance, the frame has no reference value.  Ordinary frames do not
    react to it (tests/test_parity_rules.py); profiles/stress_corpus.py holds the material that does.
  * peak counts of a frame whose windowed signal holds exactly ONE non-zero sample (the last LSB tick of a decayed tail;
    the window's end points are zero): its magnitude
    spectrum is |x w[n]| / N in every bin, so which bins are "strict local maxima above 0.25 max" (spectral_complexity,
    spectral_complexity_bands; Statistics.cpp:140-232, SampleAnalyser.cpp:2150-2187) is decided by the last bit of the
    FFT's rounding; the frame's flatness is then 1e-16-sized noise around 0, which the geometric-mean statistics of the
    flatness series (gmean, flatness = gmean / mean) amplify.  Found by profiles/parity_sweep.py; the reference built
    with another FFT disagrees with itself on these frames in the same way.
  * the geometric mean (and flatness = gmean / mean) of a series that holds values at the FP-noise level: TStatistics::
    GeometricMean sums log(|x| + 1e-20) (Statistics.cpp:417-455), so a frame value of 0 against 1e-17 -- both "zero" to
    eleven orders below the absolute tolerance, e.g. f0_confidence = (1 - yin') / 0.25 where yin' rounds to 1, or the flatness
    of a two-bin band whose bins are equal -- moves the log sum by 7 / n.  When the two series differ at an element below
    1e-9, these two statistics of the CUDA path are checked against the oracle's TStatistics restatement applied to the
    series the CUDA path itself produced (as for the noise-determined frames above), not against the reference's.
    Found by profiles/parity_sweep.py (round 2, seeds 7114 and 8030).  The same FP-noise-level element can carry the
    index-weighted statistics: an f0_confidence series with ONE frame of 1.4e-5 and, in one of the two results, one ulp
    (4.4e-16) in two other frames has spread 0 against 2.6e-9 -- on either side of the reference's 1e-12 cut-off -- and so
    skewness 0 against -1.1e32 (profiles/edge_self_sweep.py, seed 91764: the oracle's two FFT variants).  So every
    statistic that disagrees under this condition is checked on the produced series.
  * every statistic of a series that lies below the absolute tolerance as a whole in both results (f0_confidence of a tone
    next to Nyquist: one ulp in some frames, 0 in the others -- which frames is rounding): the temporal centroid, spread, ...
    divide by the sum of the values.  The CUDA statistics are checked against the oracle's TStatistics restatement applied to
    the series the CUDA path produced.  Found by profiles/stress_corpus.py (sine_21000).
Derived tolerance: the statistic flatness = gmean / mean (Statistics.cpp:69-72) is compared with the tolerance
its two inputs carry, |fl| * (tol(gmean) / |gmean| + tol(mean) / |mean|), on top of its own -- a gmean that
agrees to the absolute tolerance (series with many FP-noise values around 0, e.g. DCT rows of silent frames)
cannot give a quotient that agrees to 1e-4 relative.
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


PITCH_SERIES = ("f0", "f0_confidence", "failsafe_f0")


PITCH_SILENT_HALF = 1e-6      # first-half energy below this fraction of the frame's: -60 dB


def ill_conditioned_pitch_frames(mdata, hop, F, N=2048):
    """Frames whose first N/2 samples are silent next to the second half: exactly zero, or a few LSB ticks of a decayed
    tail -- less than PITCH_SILENT_HALF of the frame's energy (see the module docstring)."""
    x = np.asarray(mdata, dtype=np.float64)
    out = np.zeros(F, dtype=bool)
    for t in range(F):
        a, m, e = t * hop, t * hop + N // 2, min(t * hop + N, len(x))
        if m <= len(x):
            s0 = float(np.dot(x[a:m], x[a:m])); s1 = float(np.dot(x[m:e], x[m:e]))
            out[t] = s1 > 0.0 and s0 <= PITCH_SILENT_HALF * (s0 + s1)
    return out


PITCH_FFT_NOISE = 1e-15       # rounding of an FFT-made correlation, as a fraction of the frame's energy (eps x log2 N x a few)


def _yin_decision(yin, sr):
    """aubio's yinfast decision on a difference function (pitchyinfast.c:150-176, mathutils.c:250-258, 494-506;
    SampleAnalyser.cpp:887-889): cumulative-mean normalisation, first dip below 0.75, arg-min fall-back (the last minimum
    wins ties), parabolic refinement -> (f0 before the silence gate, confidence)."""
    W = len(yin)
    y = np.array(yin, dtype=np.float64)
    y[0] = 1.0
    c = np.cumsum(y[1:])
    with np.errstate(divide="ignore", invalid="ignore"):
        y[1:] = np.where(c != 0, y[1:] * np.arange(1, W) / c, 1.0)
    pos = -1
    # first p >= 2 with y[p] < 0.75 and y[p] < y[p + 1], seen at t = p + 3 <= W - 1
    cand = np.nonzero((y[2:W - 3] < 0.75) & (y[2:W - 3] < y[3:W - 2]))[0]
    if cand.size:
        pos = int(cand[0]) + 2
    else:
        m = y.min()
        pos = int(np.nonzero(y == m)[0][-1])
    if pos == 0 or pos == W - 1:
        period = float(pos)
    else:
        s0, s1, s2 = y[pos - 1], y[pos], y[pos + 1]
        den = s0 - 2.0 * s1 + s2
        period = pos + 0.5 * (s0 - s2) / den if den != 0 else float("nan")
    peak = int(period) if (period == period and period >= 0) else 0
    f0 = sr / period if period > 0 else 0.0
    conf = min(1.0, max(0.0, (1.0 - y[min(peak, W - 1)]) / 0.25))
    return f0, conf


def _yin_exact(frame):
    """The difference function of one frame in extended precision with the correlation as a direct sum; returns (yin, energy)."""
    N = 2048
    x = np.zeros(N, dtype=np.longdouble)
    f = np.asarray(frame, dtype=np.float64)[:N]
    x[:len(f)] = f
    W = N // 2
    E = float(np.dot(x, x))
    sq = np.empty(W, dtype=np.longdouble)
    s0 = np.dot(x[:W], x[:W])
    x2 = x * x
    sq[0] = s0
    sq[1:] = s0 - np.cumsum(x2[:W - 1]) + np.cumsum(x2[W:2 * W - 1])
    sq += s0
    r = np.correlate(x[:2 * W - 1], x[:W], mode="valid")           # r[tau] = sum_{m < W} x[m] x[m + tau]
    return (sq - r).astype(np.float64), E                          # aubio's halved correlation: sq - 2 (r / 2)


def pitch_noise_sensitive(frame, sr=44100.0, draws=6, seed=1) -> bool:
    """The general form of the pitch rule: True when the reference's OWN pitch decision for this frame (see _yin_decision)
    changes beyond the tolerance once its correlation r[tau] -- which the reference computes with FFTs -- is perturbed at the
    level of an FFT's rounding, PITCH_FFT_NOISE x the frame's energy.  Such a frame has no reference value to compare with:
    exact ties of the normalised difference function (a silent or tick-only first half, one impulse, a constant frame) or
    near-ties far below what double precision resolves."""
    yin, E = _yin_exact(frame)
    if E == 0.0:
        return False
    W = len(yin)
    base = _yin_decision(yin, sr)
    rng = np.random.default_rng(seed)
    for _ in range(draws):
        got = _yin_decision(yin + PITCH_FFT_NOISE * E * rng.uniform(-1.0, 1.0, W), sr)
        if not (close(got[0], base[0]) and close(got[1], base[1])):
            return True
    return False


FLAT_COUNT_SERIES = ("spectral_complexity", "spectral_complexity_bands")
FLAT_GMEAN_SERIES = ("spectral_flatness", "spectral_flatness_bands")


def impulse_frames(mdata, hop, F, N=2048):
    """Frames whose WINDOWED signal holds exactly one non-zero sample (see the module docstring).  The Hann window
    (SampleAnalyser.cpp:178-181) is zero at both ends, so a tick on the frame's first or last sample does not count."""
    x = np.asarray(mdata, dtype=np.float64)
    w = 1.0 - np.cos(2.0 * np.pi * np.arange(N) / (N - 1))
    out = np.zeros(F, dtype=bool)
    for t in range(F):
        seg = x[t * hop:t * hop + N]
        out[t] = np.count_nonzero(seg * w[:len(seg)]) == 1
    return out


def _series_iter(r: layout.FileResult):
    for i, n in enumerate(layout.FRAMED_SCALARS):
        yield n, r.fs[i]
    for i, (n, nb) in enumerate(layout.FRAMED_VECTORS):
        for b in range(nb):
            yield "%s[%d]" % (n, b), r.fv[i][:, b]


def compare(got: layout.FileResult, want: layout.FileResult, skip_series=(), only_series=None,
            check_stats=True, check_header=True, max_flip_frac=0.0, mdata=None, hop=1024):
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
    ill_pitch = ill_conditioned_pitch_frames(mdata, hop, want.F) if mdata is not None else np.zeros(want.F, dtype=bool)
    ill_flat = impulse_frames(mdata, hop, want.F) if mdata is not None else np.zeros(want.F, dtype=bool)
    if mdata is not None:
        # pitch frames that differ and that no structural rule names: is the reference's own decision determined there?
        cand = np.zeros(want.F, dtype=bool)
        for n in PITCH_SERIES:
            if not (n in skip_series or (only_series is not None and n not in only_series)):
                cand |= ~close(got.series(n), want.series(n))
        x = np.asarray(mdata, dtype=np.float64)
        for t in np.nonzero(cand & ~ill_pitch)[0]:
            if pitch_noise_sensitive(x[t * hop:t * hop + 2048]):
                ill_pitch[t] = True
    for n in names:
        if n in skip_series or (only_series is not None and n not in only_series):
            continue
        a, b = got.series(n), want.series(n)
        if n in layout.INTEGER_SERIES:
            bad = (a != b)
        else:
            bad = ~close(a, b)
        if n in PITCH_SERIES:
            bad &= ~ill_pitch
        if n in FLAT_COUNT_SERIES:
            bad &= ~(ill_flat if bad.ndim == 1 else ill_flat[:, None])
        nb = int(bad.sum())
        if nb > max_flip_frac * bad.size:
            idx = np.argwhere(bad)[0]
            errs.append("%s: %d/%d values differ, first at %s: %r != %r" % (
                n, nb, bad.size, idx.tolist(), a[tuple(idx)], b[tuple(idx)]))
    if check_stats:
        errs += compare_stats(got, want, skip_series, only_series, ill_pitch, ill_flat)
    return errs


def stat_rules(ok, x, b):
    """Apply the documented ill-conditioning rules to the 13 statistics of ONE series x whose expected row is b:
    only the statistics a rule names are released, the rest of the row stays under the plain tolerance."""
    sx = np.sum(np.abs(x))
    if sx > 0 and abs(np.sum(x)) * 1e6 < sx:             # cancelling sum: centroid..kurtosis and flatness (= gmean / mean)
        ok[6:11] = True
    if abs(abs(b[7]) - 1e-12) < 1e-6 * 1e-12 or abs(b[7]) < 1e-9:   # spread on the 1e-12 cut-off: skewness, kurtosis
        ok[8:10] = True
    return ok


def log_domain_noise(got_series, want_series) -> bool:
    """True when the two series differ at an element that is below 1e-9 in either of them (see the module docstring)."""
    a, b = np.asarray(got_series, dtype=np.float64), np.asarray(want_series, dtype=np.float64)
    tiny = (np.abs(a) < 1e-9) | (np.abs(b) < 1e-9)
    return bool(np.any(tiny & (a != b)))


def sub_tolerance_series(got_series, want_series) -> bool:
    """True when both series lie below the absolute tolerance in every element (and are not identical)."""
    a, b = np.asarray(got_series, dtype=np.float64), np.asarray(want_series, dtype=np.float64)
    return bool(a.size and np.max(np.abs(a)) <= ATOL and np.max(np.abs(b)) <= ATOL and not np.array_equal(a, b))


def flatness_tol(a, b, ok):
    if b[3] != 0.0 and b[4] != 0.0:          # flatness = gmean / mean with its inputs' tolerances
        tol = ATOL + RTOL * abs(b[10]) + abs(b[10]) * ((ATOL + RTOL * abs(b[4])) / abs(b[4]) +
                                                       (ATOL + RTOL * abs(b[3])) / abs(b[3]))
        ok[10] = ok[10] or abs(a[10] - b[10]) <= tol
    return ok


def compare_stats(got, want, skip_series=(), only_series=None, ill_pitch=None, ill_flat=None):
    """The 13 statistics of every series.  A series that holds noise-determined FRAME values (the pitch triple of
    half-silent frames, the peak counts of impulse frames) cannot have its statistics compared with the reference's --
    every statistic is a function of those frames -- so for such a series the statistics pass itself is checked
    instead: the CUDA statistics must equal the oracle's TStatistics restatement (oracle.stats13, pinned by the
    reference's own KATs) applied to the series the CUDA path produced.  Nothing is skipped."""
    errs = []
    gs = dict(_series_iter(got))
    for si, (n, x) in enumerate(_series_iter(want)):
        base = n.split("[")[0]
        if base in skip_series or (only_series is not None and base not in only_series):
            continue
        a, b = got.stats[si], want.stats[si]
        noise = (base in PITCH_SERIES and ill_pitch is not None and ill_pitch.any()) or \
                (base in FLAT_COUNT_SERIES + FLAT_GMEAN_SERIES and ill_flat is not None and ill_flat.any())
        # a series that lies below the absolute tolerance as a whole (e.g. f0_confidence = (1 - yin') / 0.25 of a tone next
        # to Nyquist: one ulp, 4.4e-16, in some frames and 0 in the others): both series ARE equal under the tolerance, and
        # every statistic that weighs or divides by the values is a function of which frames carry the ulp
        noise = noise or sub_tolerance_series(gs[n], x)
        if noise:
            from oracle import oracle
            x = np.ascontiguousarray(gs[n], dtype=np.float64)
            b = oracle.stats13(x)
        ok = close(a, b)
        ok = flatness_tol(a, b, ok)
        ok = stat_rules(ok, x, b)
        if not noise and not ok.all() and log_domain_noise(gs[n], x):
            from oracle import oracle
            mine = np.ascontiguousarray(gs[n], dtype=np.float64)
            own = oracle.stats13(mine)
            ok |= stat_rules(flatness_tol(a, own, close(a, own)), mine, own)
        if not ok.all():
            k = int(np.argwhere(~ok)[0][0])
            errs.append("stat %s_%s%s: %r != %r" % (n, layout.STAT_NAMES[k], " (of the produced series)" if noise else "", a[k], b[k]))
    return errs
