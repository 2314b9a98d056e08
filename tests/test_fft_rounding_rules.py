"""Demonstrates -- rather than asserts -- the claim behind the rule-based exclusions of tests/parity.py: the PATH itself,
run with two mathematically identical FFTs that round differently (the situation of the reference built with another
FFT back end: Fourier.cpp picks IPP / vDSP / Ooura per platform), disagrees with itself on exactly the frames the rules
name and nowhere else.  CPU only: both runs are the oracle (oracle.set_fft_variant)."""
import numpy as np
import pytest

import parity
from afec_b200 import synth
from oracle import oracle


def both_ffts(pcm, hop=1024):
    a = oracle.analyze(pcm, hop=hop, file_size=44 + pcm.size * 2)
    oracle.set_fft_variant(1)
    try:
        b = oracle.analyze(pcm, hop=hop, file_size=44 + pcm.size * 2)
    finally:
        oracle.set_fft_variant(0)
    return a, b


def test_fft_variant_is_the_same_transform():
    rng = np.random.default_rng(3)
    for n in (512, 2048):
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        outs = []
        for v in (0, 1):
            oracle.set_fft_variant(v)
            re, im = np.ascontiguousarray(x.real), np.ascontiguousarray(x.imag)
            oracle.lib().afxo_fft(re.ctypes.data, im.ctypes.data, n, 1)
            outs.append(re + 1j * im)
        oracle.set_fft_variant(0)
        assert np.max(np.abs(outs[0] - outs[1])) < 1e-11
        assert np.max(np.abs(outs[0] - np.conj(np.fft.fft(np.conj(x))))) < 1e-10
        assert not np.array_equal(outs[0], outs[1])          # same transform, different last bits


def test_impulse_frame_peak_counts_flip_with_the_fft():
    """The sweep's file 85 of seed 5000 (tests/test_gpu_parity.py::test_impulse_tail_frames)."""
    x = synth.one_shot(5085, 3.041115350932388, channels=2)
    pcm = np.ascontiguousarray((x.astype(np.float64) * 0.03786796017229539).astype(np.int16))
    data = oracle.condition(pcm)[0]
    a, b = both_ffts(pcm)
    ill = parity.impulse_frames(data, 1024, a.F)
    assert ill.any()
    diff = a.series("spectral_complexity") != b.series("spectral_complexity")
    assert diff.any(), "expected the two FFTs to disagree on the peak count of an impulse frame"
    assert not (diff & ~ill).any()                            # ... and only there
    assert parity.compare(a, b), "the plain tolerance must flag the disagreement"
    assert parity.compare(a, b, mdata=data) == []             # the documented rules account for all of it


def test_half_silent_frame_pitch_flips_with_the_fft():
    """A frame whose first 1024 samples are digital silence: yin[tau] = sq[tau] - r[tau] is FFT rounding noise for small tau."""
    pcm = synth.one_shot(77, 1.2).copy()
    g0 = 20000
    pcm[g0:g0 + 5000] = 0                                      # a gap of digital silence inside the file
    data = oracle.condition(pcm)[0]
    a, b = both_ffts(pcm)
    ill = parity.ill_conditioned_pitch_frames(data, 1024, a.F)
    assert ill.any()
    d = np.zeros(a.F, dtype=bool)
    for n in parity.PITCH_SERIES:
        d |= ~parity.close(a.series(n), b.series(n))
    assert not (d & ~ill).any()                               # outside the rule's frames the two FFTs agree
    assert parity.compare(a, b, mdata=data) == []


@pytest.mark.parametrize("ticks", [(4976,), (4400, 4976)])
def test_ticks_in_a_silent_half_pitch_flips_with_the_fft(ticks):
    """The same with one or two least-significant-bit ticks inside the silent half (the sweep's files of seeds 15064, frame
    230, and 16138, frame 350): r[tau] is the ticks times the signal tau samples on -- zero across the gap -- so the
    normalised function ties at exactly 1 over the small lags and the arg-min fall-back picks among FFT rounding noise."""
    pcm = synth.one_shot(77, 1.2).copy()
    g0 = 20000
    pcm[g0:g0 + 6000] = 0
    for k in ticks:                                            # the last one 1024 samples before the gap ends: whatever the frame
        pcm[g0 + k] = 1                                        # grid's offset, one frame has the ticks alone in its first half and signal in its second
    data = oracle.condition(pcm)[0]
    a, b = both_ffts(pcm)
    ill = parity.ill_conditioned_pitch_frames(data, 1024, a.F)
    only_zero_half = np.zeros(a.F, dtype=bool)
    x = np.asarray(data)
    for t in range(a.F):
        only_zero_half[t] = not x[t * 1024:t * 1024 + 1024].any()
    assert (ill & ~only_zero_half).any(), "expected a frame whose first half holds nothing but the ticks"
    d = np.zeros(a.F, dtype=bool)
    for n in parity.PITCH_SERIES:
        d |= ~parity.close(a.series(n), b.series(n))
    assert (d & ill & ~only_zero_half).any(), "expected the two FFTs to disagree on a tick frame"
    assert not (d & ~ill).any()                               # outside the rule's frames the two FFTs agree
    assert parity.compare(a, b, mdata=data) == []


@pytest.mark.parametrize("kind", ["one_impulse", "dc"])
def test_tie_frames_of_other_shapes_pitch_flips_with_the_fft(kind):
    """One impulse alone in a frame (yin[tau] = x0^2 for every tau) and a constant signal (yin[tau] = W c^2) tie the
    normalised difference function exactly as well; no structural rule names them -- the general test of
    parity.pitch_noise_sensitive() does (profiles/stress_corpus.py found both)."""
    n = 66150
    if kind == "one_impulse":
        pcm = np.zeros(n, dtype=np.int16); pcm[n // 2] = 30000
    else:
        pcm = np.full(n, 1000, dtype=np.int16)
    data = oracle.condition(pcm)[0]
    a, b = both_ffts(pcm)
    d = np.zeros(a.F, dtype=bool)
    for name in parity.PITCH_SERIES:
        d |= ~parity.close(a.series(name), b.series(name))
    assert d.any(), "expected the two FFTs to disagree on the pitch of a tie frame"
    assert not parity.ill_conditioned_pitch_frames(data, 1024, a.F)[d].all()      # not (all) covered by the structural rule
    x = np.asarray(data, dtype=np.float64)
    assert all(parity.pitch_noise_sensitive(x[t * 1024:t * 1024 + 2048]) for t in np.nonzero(d)[0])
    assert parity.compare(a, b), "the plain tolerance must flag the disagreement"
    assert parity.compare(a, b, mdata=data) == []


@pytest.mark.parametrize("seed", [11, 12, 13, 14])
def test_ordinary_files_do_not_depend_on_the_fft(seed):
    """On ordinary material every output agrees between the two FFTs under the rules: the exclusions are not a blanket."""
    pcm = synth.one_shot(seed, 0.8 + 0.3 * (seed % 3))
    data = oracle.condition(pcm)[0]
    a, b = both_ffts(pcm, hop=512 if seed % 2 else 1024)
    assert parity.compare(a, b, mdata=data, hop=512 if seed % 2 else 1024) == []
