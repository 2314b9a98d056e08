"""The frame classifiers behind the documented parity exclusions (tests/parity.py), on hand-made signals."""
import numpy as np

import parity


def test_impulse_frames_are_exactly_the_single_sample_frames():
    hop, N = 1024, 2048
    x = np.zeros(hop * 6 + N)
    x[:N] = 0.1                       # frame 0 (and the first half of frame 1) full of signal
    x[hop * 4 + 7] = 3.0e-5           # one LSB tick: frames 3 and 4 see it alone
    F = (len(x) - N) // hop + 1
    got = parity.impulse_frames(x, hop, F)
    want = np.zeros(F, dtype=bool)
    want[[3, 4]] = True
    assert np.array_equal(got, want)
    # two ticks in the same frame are not an impulse frame ...
    x[hop * 4 + 900] = -3.0e-5
    assert not parity.impulse_frames(x, hop, F)[3:5].any()
    # ... unless one of them sits on the window's zero end point
    y = np.zeros(hop * 3 + N)
    y[hop + 100] = 1.0e-4; y[hop + N - 1] = -2.0e-4           # frame 1: sample 100 and the last sample
    assert parity.impulse_frames(y, hop, (len(y) - N) // hop + 1)[1]


def test_half_silent_pitch_frames():
    hop, N = 1024, 2048
    x = np.zeros(hop * 4 + N)
    x[hop * 2 + N // 2:] = 0.2        # frame 2: first half silent, second half not
    F = (len(x) - N) // hop + 1
    got = parity.ill_conditioned_pitch_frames(x, hop, F)
    assert got[2] and not got[0] and not got[3]


def test_close_uses_the_stated_tolerance():
    assert parity.close(1.0 + 0.9e-4, 1.0) and not parity.close(1.0 + 1.2e-4, 1.0)
    assert parity.close(5e-7, 0.0) and not parity.close(2e-6, 0.0)
    assert parity.close(np.nan, np.nan)


def test_log_domain_noise_rule_only_fires_on_sub_tolerance_differences():
    a = np.array([0.5, 0.25, 0.0, 0.75])
    assert not parity.log_domain_noise(a, a.copy())                       # identical, zeros included
    b = a.copy(); b[2] = 1e-17
    assert parity.log_domain_noise(a, b) and parity.log_domain_noise(b, a)
    c = a.copy(); c[1] = 0.2500001                                         # an ordinary difference is not "noise"
    assert not parity.log_domain_noise(a, c)
    # and the geometric mean really is that sensitive: 0 against 1e-17 in one of 200 frames moves it by percents
    from oracle import oracle
    x = np.full(200, 0.5); y = x.copy(); x[7] = 0.0; y[7] = 1e-17
    gx, gy = oracle.stats13(x)[4], oracle.stats13(y)[4]
    assert abs(gx - gy) > 1e-2 * gx and parity.close(x, y).all()


def test_yin_decision_restates_the_oracle():
    """The decision procedure behind pitch_noise_sensitive() is a restatement of the oracle's (and so of aubio's): on
    ordinary material, fed the exact difference function, it returns the oracle's f0 and confidence frame for frame."""
    from afec_b200 import synth
    from oracle import oracle
    n_checked = 0
    for seed in (21, 22, 23):
        pcm = synth.one_shot(seed, 1.5)
        want = oracle.analyze(pcm, file_size=44 + pcm.size * 2)
        data = np.asarray(oracle.condition(pcm)[0], dtype=np.float64)
        ill = parity.ill_conditioned_pitch_frames(data, 1024, want.F)
        f0, conf = want.series("f0"), want.series("f0_confidence")
        for t in range(want.F):
            if ill[t]:
                continue
            yin, E = parity._yin_exact(data[t * 1024:t * 1024 + 2048])
            g0, gc = parity._yin_decision(yin, 44100.0)
            assert parity.close(gc, conf[t]), (seed, t, gc, conf[t])
            if f0[t] != 0.0:                                  # the silence gate (-48 dB) sits outside the decision
                assert parity.close(g0, f0[t]), (seed, t, g0, f0[t])
            n_checked += 1
    assert n_checked > 100


def test_pitch_noise_sensitivity_names_the_tie_frames_only():
    from afec_b200 import synth
    from oracle import oracle
    N = 2048
    impulse = np.zeros(N); impulse[0] = 1.0                   # yin[tau] = x0^2 for every tau
    assert parity.pitch_noise_sensitive(impulse)
    assert parity.pitch_noise_sensitive(np.full(N, 0.03))     # constant: yin[tau] = W c^2
    ticks = np.zeros(N); ticks[370] = ticks[961] = 3.8e-5     # the sweep's seed 16138, frame 350
    ticks[1256:1286] = 0.4 * np.hanning(30)
    assert parity.pitch_noise_sensitive(ticks)
    assert not parity.pitch_noise_sensitive(np.zeros(N))      # silence: nothing to decide
    # ordinary material does not react to rounding-level noise: the test is not a blanket
    flagged = total = 0
    for seed in (31, 32, 33, 34):
        pcm = synth.one_shot(seed, 1.2)
        data = np.asarray(oracle.condition(pcm)[0], dtype=np.float64)
        F = (len(data) - N) // 1024 + 1
        ill = parity.ill_conditioned_pitch_frames(data, 1024, F)
        for t in range(F):
            if not ill[t]:
                total += 1
                flagged += bool(parity.pitch_noise_sensitive(data[t * 1024:t * 1024 + N]))
    assert total > 100 and flagged == 0


def test_sub_tolerance_series_rule():
    """A series of rounding-level values (one ulp here, zero there) equals its counterpart under the tolerance element by
    element, but its index-weighted statistics are a function of WHERE the ulps sit (profiles/stress_corpus.py, sine_21000)."""
    from oracle import oracle
    a = np.zeros(64); b = np.zeros(64)
    a[[5, 7, 15, 40]] = 4.44e-16
    b[[6, 30, 50]] = 4.44e-16
    assert parity.close(a, b).all()
    sa, sb = oracle.stats13(a), oracle.stats13(b)
    assert not parity.close(sa[6], sb[6])                    # the temporal centroids differ by far more than the tolerance
    assert parity.sub_tolerance_series(a, b)
    assert not parity.sub_tolerance_series(a, a)             # identical series: nothing to release
    c = b.copy(); c[3] = 1e-3
    assert not parity.sub_tolerance_series(a, c)             # one element above the tolerance: the plain rules apply


def test_noise_level_elements_can_carry_the_index_weighted_statistics():
    """profiles/edge_self_sweep.py, seed 91764: one frame of 1.4e-5 and -- in one result only -- one ulp in two other frames put
    the spread on either side of the reference's 1e-12 cut-off: skewness 0 against -1e32.  The series agree element by element;
    the rule checks such statistics on the produced series and still catches a wrong statistic."""
    from afec_b200 import layout
    from oracle import oracle
    a = np.zeros(160); b = np.zeros(160)
    a[126] = b[126] = 1.377e-5
    a[[40, 90]] = 4.44e-16
    sa, sb = oracle.stats13(a), oracle.stats13(b)
    k = layout.STAT_NAMES.index("skewness")
    assert sb[k] == 0.0 and abs(sa[k]) > 1e20
    assert parity.close(a, b).all() and parity.log_domain_noise(a, b)
    ok = parity.close(sb, sa)
    assert not ok[k]
    own = oracle.stats13(b)                                   # what compare_stats does under the rule
    assert parity.close(sb, own).all()
    wrong = sb.copy(); wrong[k] = 3.0
    assert not parity.close(wrong, own)[k]                    # a wrong statistic of the produced series is still caught
