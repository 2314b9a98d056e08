"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the reference goldens.
Run on the B200 box: python -m pytest tests -m gpu."""
import numpy as np
import pytest

import golden_io
import parity
from afec_b200 import api, layout, synth

pytestmark = pytest.mark.gpu

CASES = golden_io.load()

SPECTRAL = ["spectral_rms", "spectral_centroid", "spectral_rolloff", "spectral_spread", "spectral_skewness",
            "spectral_kurtosis", "spectral_flatness", "spectral_flux", "spectral_inharmonicity",
            "tristimulus1", "tristimulus2", "tristimulus3"]
AMPLITUDE = ["amplitude_silence", "amplitude_peak", "amplitude_rms", "amplitude_envelope"]
GROUPS = {
    api.FEAT_SPECTRAL: SPECTRAL,
    api.FEAT_AMPLITUDE: AMPLITUDE,
    api.FEAT_PEAKS: ["spectral_complexity"],
    api.FEAT_BANDS: ["spectral_contrast", "spectral_rms_bands", "spectral_flatness_bands", "spectral_flux_bands",
                     "spectral_complexity_bands", "spectral_contrast_bands", "frequency_bands", "cepstrum_bands"],
    api.FEAT_PITCH: ["f0", "f0_confidence", "failsafe_f0"],
    api.FEAT_AUTOCORR: ["auto_correlation"],
    api.FEAT_RHYTHM: ["rhythm_complex_onsets", "rhythm_percussive_onsets"],
}


def implemented_features() -> int:
    """Feature groups compiled into the library (all of them once the path is complete)."""
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", api.LIB_PATH], capture_output=True, text=True).stdout
    feats = api.FEAT_SPECTRAL | api.FEAT_AMPLITUDE
    for flag, sym in ((api.FEAT_PEAKS, "afx_launch_peaks"), (api.FEAT_BANDS, "afx_launch_bands"),
                      (api.FEAT_PITCH, "afx_launch_pitch"), (api.FEAT_AUTOCORR, "afx_launch_autocorr"),
                      (api.FEAT_RHYTHM, "afx_launch_rhythm"), (api.FEAT_STATS, "afx_launch_stats")):
        if sym in out:
            feats |= flag
    return feats


def series_for(features: int):
    names = []
    for flag, ns in GROUPS.items():
        if features & flag:
            names += ns
    return names


@pytest.fixture(scope="module")
def feats():
    return implemented_features()


@pytest.fixture(scope="module")
def analysers(feats):
    cache = {}

    def get(hop):
        if hop not in cache:
            cache[hop] = api.SampleAnalyser(44100, 2048, hop, features=feats)
        return cache[hop]
    yield get
    for a in cache.values():
        a.close()


def check(got, want, feats, **kw):
    full = (feats & api.FEAT_ALL) == api.FEAT_ALL
    errs = parity.compare(got, want, only_series=None if full else series_for(feats),
                          check_stats=bool(feats & api.FEAT_STATS), check_header=full, **kw)
    assert not errs, "\n".join(errs[:25])


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_golden_reference_vectors(analysers, feats, case):
    """CUDA path vs vectors produced by the unmodified reference (tests/golden/make_golden.py)."""
    if case["rate"] != 44100 and not hasattr(api, "HAVE_RESAMPLE"):
        pytest.skip("resampler not built yet")
    got = analysers(case["hop"]).analyze_pcm([case["pcm"]], [case["rate"]])[0]
    check(got, case["ref"], feats)


@pytest.mark.parametrize("hop", [1024, 512, 768, 256, 2048])      # every samples-per-thread class of the amplitude features
def test_batch_vs_oracle(analysers, feats, oracle_lib, hop):
    """A mixed batch (ragged lengths, stereo, tiny, silent) in ONE call vs the oracle file by file."""
    pcms = [synth.one_shot(400 + i, 0.3 + 0.4 * i) for i in range(6)]
    pcms.append(synth.one_shot(410, 0.9, channels=2))
    pcms.append(synth.one_shot(411, 0.02))
    pcms.append(np.zeros(30000, dtype=np.int16))
    pcms.append(synth.one_shot(412, 21.5))                 # crosses the 20 s cap
    pcms.append(synth.one_shot(413, 0.7, channels=5).astype(np.float32))   # float32 input path
    rates = [44100] * len(pcms)
    got = analysers(hop).analyze_pcm(pcms, rates)
    for g, p in zip(got, pcms):
        want = oracle_lib.analyze(p, hop=hop, file_size=44 + p.size * p.itemsize)
        check(g, want, feats)


# every k_spectrum instantiation the launcher can pick (afx_spectrum.cu: <10, false> without the pitch / amplitude code
# for the spectral-only subset of BASELINE configs[1], <10, true> otherwise) meets the oracle
SUBSETS = [api.FEAT_SPECTRAL, api.FEAT_SPECTRAL | api.FEAT_STATS, api.FEAT_SPECTRAL | api.FEAT_AMPLITUDE | api.FEAT_STATS,
           api.FEAT_SPECTRAL | api.FEAT_PEAKS | api.FEAT_BANDS | api.FEAT_STATS]


@pytest.mark.parametrize("features", SUBSETS, ids=["spectral", "spectral+stats", "spectral+amplitude+stats", "spectral+peaks+bands+stats"])
@pytest.mark.parametrize("hop", [512, 1024])
def test_feature_subsets_vs_oracle(oracle_lib, hop, features):
    """The feature-mask subsets (include/afec_b200.h AFX_FEAT_*) against the oracle: BASELINE configs[1] runs
    FEAT_SPECTRAL alone at hop 512."""
    pcms = [synth.one_shot(400 + i, 0.3 + 0.4 * i) for i in range(6)]
    pcms.append(synth.one_shot(410, 0.9, channels=2))
    pcms.append(synth.one_shot(411, 0.02))
    pcms.append(np.zeros(30000, dtype=np.int16))
    pcms.append(synth.one_shot(412, 21.5))
    an = api.SampleAnalyser(44100, 2048, hop, features=features)
    got = an.analyze_pcm(pcms, [44100] * len(pcms))
    an.close()
    for g, p in zip(got, pcms):
        want = oracle_lib.analyze(p, hop=hop, file_size=44 + p.size * p.itemsize)
        errs = parity.compare(g, want, only_series=series_for(features), check_stats=bool(features & api.FEAT_STATS),
                              check_header=False, mdata=oracle_lib.condition(p)[0], hop=hop)
        assert not errs, "\n".join(errs[:25])


@pytest.mark.parametrize("case", [c for c in CASES if c["rate"] == 44100], ids=[c["name"] for c in CASES if c["rate"] == 44100])
def test_golden_reference_vectors_spectral_subset(case):
    """The reference's own vectors against the spectral-only instantiation (the kernel BASELINE configs[1] times)."""
    an = api.SampleAnalyser(44100, 2048, case["hop"], features=api.FEAT_SPECTRAL)
    got = an.analyze_pcm([case["pcm"]], [case["rate"]])[0]
    an.close()
    errs = parity.compare(got, case["ref"], only_series=SPECTRAL, check_stats=False, check_header=False)
    assert not errs, "\n".join(errs[:25])


def test_unsupported_description_fails_that_file_only(analysers, feats, oracle_lib):
    """A file description the kernels cannot take (rate <= 0, unknown format) gets AFX_FILE_UNSUPPORTED; its
    neighbours are analysed (the reference fails files one by one, SampleAnalyser.cpp:372-408)."""
    good = synth.one_shot(601, 0.5)
    an = analysers(1024)
    f_good, keep = an.describe(good, 44100)
    f_rate, _ = an.describe(good, 44100); f_rate.src_rate = -5
    f_fmt, _ = an.describe(good, 44100); f_fmt.format = 99
    b = an.batch_from_descriptors([f_good, f_rate, f_fmt, f_good], keep).run()
    got = [b.result(i) for i in range(4)]
    b.free()
    assert [g.status for g in got] == [0, 3, 3, 0]
    want = oracle_lib.analyze(good, file_size=44 + good.size * 2)
    check(got[0], want, feats)
    check(got[3], want, feats)


def test_many_resampler_shapes_in_one_batch(analysers, oracle_lib):
    """More distinct (rate, length) resampler shapes in ONE batch than the process-wide shape cache holds: every file
    must still get its own libresample time stamps (a stale cache entry once gave a file another file's)."""
    rng = np.random.default_rng(5)
    base = synth.one_shot(950, 0.4, rate=48000)
    pcms = [np.ascontiguousarray(base[: 6000 + 17 * i]) for i in range(600)]
    an = analysers(1024)
    b = an.batch(pcms, [48000] * len(pcms)).run()
    for i in rng.choice(len(pcms), 24, replace=False).tolist() + [0, 1, 255, 256, 257, 511, 512, 513, 599]:
        data = oracle_lib.condition(pcms[i], src_rate=48000)[0]
        got = b.conditioned(i)
        assert got.shape == data.shape and np.array_equal(got, data), "file %d" % i
    b.free()


def test_one_live_batch_per_context(analysers):
    """include/afec_b200.h: a second batch may not be uploaded on a context while another one is alive."""
    an = analysers(1024)
    p = synth.one_shot(602, 0.3)
    b1 = an.batch([p], [44100]).run()
    b2 = an.batch([p], [44100])
    with pytest.raises(api.AfxError):
        b2.upload()
    b1.free()
    b2.run()
    b2.free()


def test_conditioning_bit_exact(analysers, oracle_lib):
    """mData (SampleAnalyser.cpp:698-718) must be bit-identical: integer trim / pad + one multiply."""
    pcms = [synth.one_shot(500 + i, 0.2 + 0.3 * i, channels=1 + (i % 3)) for i in range(5)]
    pcms.append(np.zeros(5000, dtype=np.int16))
    an = analysers(1024)
    b = an.batch(pcms, [44100] * len(pcms)).run()
    for i, p in enumerate(pcms):
        data, off, pk, rms = oracle_lib.condition(p)
        got = b.conditioned(i)
        assert got.shape == data.shape
        assert np.array_equal(got, data)
        r = b.result(i)
        assert r.header[layout.HEADER_NAMES.index("data_offset")] == off
        assert r.header[layout.HEADER_NAMES.index("peak_value")] == pk
        assert abs(r.header[layout.HEADER_NAMES.index("rms_value")] - rms) <= 1e-6 * max(rms, 1e-9)
    b.free()


def test_rejected_files_do_not_poison_batch(analysers, feats, oracle_lib):
    """SampleAnalyser.cpp:472-482: bad channel count / empty file fail alone; neighbours still analyse."""
    good = synth.one_shot(600, 0.5)
    pcms = [good, np.zeros((100, 9), dtype=np.int16), np.zeros((0,), dtype=np.int16), good]
    got = analysers(1024).analyze_pcm(pcms, [44100] * 4)
    assert got[1].status == 1 and got[2].status == 2
    want = oracle_lib.analyze(good, file_size=44 + good.size * 2)
    check(got[0], want, feats)
    check(got[3], want, feats)


def test_empty_batch(analysers):
    assert analysers(1024).analyze_pcm([], []) == []


def test_impulse_tail_frames(analysers, feats, oracle_lib):
    """A decayed tail whose frames hold a single LSB tick: flat spectra, peak counts decided by FFT rounding (parity.py
    rule, found by profiles/parity_sweep.py with this very file); everything else of the file must still agree."""
    rate = 44100                                               # file 85 of the sweep's seed-5000 corpus: a quiet stereo one-shot
    x = synth.one_shot(5085, 3.041115350932388, rate=rate, channels=2)
    pcm = np.ascontiguousarray((x.astype(np.float64) * 0.03786796017229539).astype(np.int16))
    data = oracle_lib.condition(pcm, src_rate=rate)[0]
    assert parity.impulse_frames(data, 1024, (len(data) - 2048) // 1024 + 1).any()
    got = analysers(1024).analyze_pcm([pcm], [rate])[0]
    want = oracle_lib.analyze(pcm, src_rate=rate, file_size=44 + pcm.size * 2)
    check(got, want, feats, mdata=data)


@pytest.mark.parametrize("ticks", [(24976,), (24400, 24976)])
def test_ticks_in_a_silent_half_frame(analysers, feats, oracle_lib, ticks):
    """A frame with one or two LSB ticks in its first half and signal in its second: the pitch arg-min falls among exact ties
    that FFT rounding breaks (parity.py rule, found by profiles/parity_sweep.py seeds 15064 and 16138;
    tests/test_fft_rounding_rules.py shows the oracle disagreeing with itself there); everything else of the file must
    still agree."""
    pcm = synth.one_shot(77, 1.2).copy()
    pcm[20000:26000] = 0
    for k in ticks:
        pcm[k] = 1
    data = oracle_lib.condition(pcm)[0]
    got = analysers(1024).analyze_pcm([pcm], [44100])[0]
    want = oracle_lib.analyze(pcm, file_size=44 + pcm.size * 2)
    assert parity.ill_conditioned_pitch_frames(data, 1024, want.F).any()
    check(got, want, feats, mdata=data)


def test_edge_material_corpus(analysers, feats, oracle_lib):
    """Steady tones on and off bin centres, square / saw / impulse trains, DC, clipped noise, a Nyquist tone, one impulse in
    silence, phase-inverted stereo, few-LSB material ... (tests/edge_corpus.py) in ONE batch against the oracle, under the
    rules of tests/parity.py -- the frames of this material whose pitch the reference's own arithmetic does not determine
    (exact ties: impulse, constant) are the ones tests/test_fft_rounding_rules.py demonstrates."""
    import edge_corpus
    files = edge_corpus.build()
    got = analysers(1024).analyze_pcm([p for _, p, _ in files], [r for _, _, r in files])
    bad = []
    for g, (name, p, r) in zip(got, files):
        want = oracle_lib.analyze(p, src_rate=r, file_size=44 + p.size * 2)
        data = oracle_lib.condition(p, src_rate=r)[0] if want.status == 0 else None
        full = (feats & api.FEAT_ALL) == api.FEAT_ALL
        errs = parity.compare(g, want, only_series=None if full else series_for(feats), check_stats=bool(feats & api.FEAT_STATS),
                              check_header=full, mdata=data)
        if errs:
            bad.append("%s: %s" % (name, errs[:3]))
    assert not bad, "\n".join(bad)


def test_trim_releases_and_regrows(feats):
    """afx_trim gives the device buffers back; the next batch grows them again and computes the same bits."""
    import torch
    an = api.SampleAnalyser(44100, 2048, 1024, features=feats)
    pcms = [synth.one_shot(600 + i, 1.0 + i) for i in range(4)]
    first = an.analyze_pcm(pcms, [44100] * 4)
    used = torch.cuda.mem_get_info()[0]
    an.trim()
    assert torch.cuda.mem_get_info()[0] > used          # free memory went up
    again = an.analyze_pcm(pcms, [44100] * 4)
    for a, b in zip(first, again):
        assert (a.F, a.Fr) == (b.F, b.Fr)
        for x, y in zip(a.fs, b.fs):
            assert np.array_equal(x, y)
    an.close()


def test_large_batch_properties(analysers, feats):
    """Size-independent properties at bench scale: identical files give identical rows; results do not
    depend on batch composition or order."""
    base = [synth.one_shot(700 + i, 1.0 + 0.1 * i) for i in range(8)]
    pcms = [base[i % 8] for i in range(512)]
    an = analysers(512)
    got = an.analyze_pcm(pcms, [44100] * len(pcms))
    solo = an.analyze_pcm(base, [44100] * 8)
    for i, g in enumerate(got):
        s = solo[i % 8]
        assert (g.F, g.Fr) == (s.F, s.Fr)
        for a, b in zip(g.fs, s.fs):
            assert np.array_equal(a, b)
        for a, b in zip(g.fv, s.fv):
            assert np.array_equal(a, b)


def test_launch_groups_do_not_change_results(feats, monkeypatch):
    """The frame-level kernels run group by group over a bounded scratch (afx_api.cu); results must not
    depend on where the group boundaries fall."""
    pcms = [synth.one_shot(800 + i, 0.4 + 0.35 * i) for i in range(7)]
    pcms.insert(3, np.zeros((0,), dtype=np.int16))           # a rejected file inside a group
    rates = [44100] * len(pcms)
    whole = api.SampleAnalyser(44100, 2048, 1024, features=feats)
    want = whole.analyze_pcm(pcms, rates)
    whole.close()
    monkeypatch.setenv("AFX_GROUP_FRAMES", "60")
    monkeypatch.setenv("AFX_GROUP_RFRAMES", "500")
    split = api.SampleAnalyser(44100, 2048, 1024, features=feats)
    got = split.analyze_pcm(pcms, rates)
    split.close()
    for g, w in zip(got, want):
        assert g.status == w.status and (g.F, g.Fr) == (w.F, w.Fr)
        assert np.array_equal(g.header, w.header)
        for a, b in zip(g.fs + g.fv, w.fs + w.fv):
            assert np.array_equal(a, b)
        assert np.array_equal(g.stats, w.stats)


def test_raw_sample_formats_convert_on_the_device(analysers, feats, oracle_lib):
    """Every AFX_PCM_* raw format (include/afec_b200.h): the bytes as they sit in a WAV / AIFF file go up unconverted and
    k_downmix applies the decoders' sample conversion (SampleConverter.h:392-518) -- the conditioned signal must equal the
    oracle's on the float32 samples the reference's decoders produce for those bytes (pinned against the live reference
    in tests/test_oracle_vs_reference.py), bit for bit."""
    import audio_files
    pcm = synth.one_shot(62, 0.4, rate=48000, channels=2)
    mono = synth.one_shot(63, 0.3)
    an = analysers(1024)
    files, keep, want = [], [], []
    codes = {("u8", False): api.AFX_PCM_U8, ("i8", True): api.AFX_PCM_I8, ("i16", False): api.AFX_PCM_I16, ("i16", True): api.AFX_PCM_I16BE,
             ("i24", False): api.AFX_PCM_I24, ("i24", True): api.AFX_PCM_I24BE, ("i32", False): api.AFX_PCM_I32, ("i32", True): api.AFX_PCM_I32BE,
             ("f32", False): api.AFX_PCM_F32U, ("f32", True): api.AFX_PCM_F32UBE}
    for src, rate in ((pcm, 48000), (mono, 44100)):
        for (kind, big), code in codes.items():
            values = audio_files.quantise(src, kind)
            raw = np.frombuffer(audio_files._sample_bytes(values, kind, big), dtype=np.uint8).copy()
            keep.append(raw)
            files.append(api.AfxFile(raw.ctypes.data, values.shape[0], values.shape[1], rate, code, 16, 1234))
            want.append((audio_files.to_float16range(values, kind), rate))
    b = an.batch_from_descriptors(files, keep).run()
    for i, (x, rate) in enumerate(want):
        data = oracle_lib.condition(x, src_rate=rate)[0]
        got = b.conditioned(i)
        assert got.shape == data.shape and np.array_equal(got, data), "format case %d" % i
        check(b.result(i), oracle_lib.analyze(x, src_rate=rate, file_size=1234), feats, mdata=data)
    b.free()


def test_rhythm_front_end_fused_equals_split(feats, monkeypatch, oracle_lib):
    """The rhythm front end has two schedules (afx_rhythm.cu): ONE fused kernel with a CTA per file (launch groups with
    many files) and the split polar / whiten / odf / power kernels (few files).  Both must give the same bits, and both
    must meet the oracle."""
    pcms = [synth.one_shot(820 + i, 0.3 + 0.45 * i) for i in range(7)]
    pcms += [synth.one_shot(830, 21.5), synth.one_shot(831, 0.02), np.zeros(30000, dtype=np.int16),
             synth.one_shot(832, 0.9, channels=2), np.zeros((0,), dtype=np.int16), synth.one_shot(833, 0.1)]
    rates = [44100] * len(pcms)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("AFX_RHYTHM_FUSED", mode)
        an = api.SampleAnalyser(44100, 2048, 1024, features=feats)
        out[mode] = an.analyze_pcm(pcms, rates)
        an.close()
    for g, w, p in zip(out["1"], out["0"], pcms):
        assert g.status == w.status and (g.F, g.Fr) == (w.F, w.Fr)
        assert np.array_equal(g.header, w.header)
        for a, b in zip(g.fs + g.fv, w.fs + w.fv):
            assert np.array_equal(a, b)
        assert np.array_equal(g.stats, w.stats)
        if g.status == 0:
            check(g, oracle_lib.analyze(p, file_size=44 + p.size * p.itemsize), feats, mdata=oracle_lib.condition(p)[0])


def test_rhythm_pipeline_equals_split(feats, monkeypatch, oracle_lib):
    """Whitening + the two onset functions run either as three kernels over the polar rows or as one producer / consumer
    pipeline per file (k_rhythm_pipe, large launch groups; forced here with AFX_RHYTHM_PIPE=1).  Same bits, both meet the
    oracle.  File lengths cover 0, 1 .. 3 rhythm frames, one ring half exactly, and many hand-overs."""
    pcms = [synth.one_shot(1820 + i, 0.2 + 0.41 * i) for i in range(7)]
    pcms += [synth.one_shot(1830, 17.5), synth.one_shot(1831, 0.02), np.zeros(30000, dtype=np.int16), synth.one_shot(1834, 0.012),
             synth.one_shot(1832, 0.9, channels=2), np.zeros((0,), dtype=np.int16), synth.one_shot(1833, 0.1), synth.one_shot(1835, 0.031)]
    rates = [44100] * len(pcms)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("AFX_RHYTHM_PIPE", mode)
        an = api.SampleAnalyser(44100, 2048, 1024, features=feats)
        out[mode] = an.analyze_pcm(pcms, rates)
        an.close()
    for g, w, p in zip(out["1"], out["0"], pcms):
        assert g.status == w.status and (g.F, g.Fr) == (w.F, w.Fr)
        assert np.array_equal(g.header, w.header)
        for a, b in zip(g.fs + g.fv, w.fs + w.fv):
            assert np.array_equal(a, b)
        assert np.array_equal(g.stats, w.stats)
        if g.status == 0:
            check(g, oracle_lib.analyze(p, file_size=44 + p.size * p.itemsize), feats, mdata=oracle_lib.condition(p)[0])


def test_pitch_block_sharing_form_vs_general(feats, monkeypatch, oracle_lib):
    """At hop 1024 the pitch kernel shares one block transform between consecutive frames (k_pitch_hop, afx_pitch.cu); the
    general 2048-point kernel (AFX_PITCH_GENERIC=1, every other hop) computes the same correlation another way.  Both
    must meet the oracle; against each other they agree to rounding on every frame the parity rules do not release
    (runs start at chunk and file boundaries: long files, short files, dead slots and padding are all in the batch)."""
    pcms = [synth.one_shot(860 + i, 0.25 + 0.5 * i) for i in range(8)]
    pcms += [synth.one_shot(870, 23.0), synth.one_shot(871, 0.02), np.zeros(40000, dtype=np.int16),
             synth.one_shot(872, 1.3, channels=2), np.zeros((0,), dtype=np.int16), synth.one_shot(873, 0.1),
             synth.one_shot(874, 6.0, rate=48000)]
    rates = [44100] * (len(pcms) - 1) + [48000]
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("AFX_PITCH_GENERIC", mode)
        an = api.SampleAnalyser(44100, 2048, 1024, features=feats)
        out[mode] = an.analyze_pcm(pcms, rates)
        an.close()
    n_frames = 0
    for g, w, p, r in zip(out["0"], out["1"], pcms, rates):
        assert g.status == w.status and (g.F, g.Fr) == (w.F, w.Fr)
        if g.status != 0:
            continue
        want = oracle_lib.analyze(p, src_rate=r, file_size=44 + p.size * p.itemsize)
        mdata = oracle_lib.condition(p, src_rate=r)[0]
        check(g, want, feats, mdata=mdata)
        check(w, want, feats, mdata=mdata)
        n_frames += g.F
    assert n_frames > 1500


def test_peaks_fused_equals_split(feats):
    """The whitening + peak count has three schedules (afx_peaks.cu): the split pair, a fused per-file kernel
    (AFX_PEAKS_FUSED=1) and the producer / consumer pipeline that large launch groups get (forced here with
    AFX_PEAKS_PIPE=1); all must give the same counts."""
    import subprocess
    import sys
    code = ("import sys, numpy as np; sys.path.insert(0, '.'); from afec_b200 import api, synth\n"
            "pcms = [synth.one_shot(840 + i, 0.3 + 0.9 * i) for i in range(6)] + [np.zeros(30000, dtype=np.int16), synth.one_shot(850, 0.02),"
            " synth.one_shot(851, 19.0), synth.one_shot(852, 0.05)]\n"
            "an = api.SampleAnalyser(44100, 2048, 1024, features=%d)\n"
            "r = an.analyze_pcm(pcms, [44100] * len(pcms))\n"
            "print(';'.join(','.join('%%d' %% v for v in x.series('spectral_complexity')) for x in r))" % feats)
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    envs = [dict(AFX_PEAKS_FUSED="0", AFX_PEAKS_PIPE="0"), dict(AFX_PEAKS_FUSED="1", AFX_PEAKS_PIPE="0"), dict(AFX_PEAKS_FUSED="0", AFX_PEAKS_PIPE="1")]
    outs = [subprocess.run([sys.executable, "-c", code], cwd=root, env=dict(os.environ, **e), capture_output=True, text=True,
                           check=True).stdout for e in envs]
    assert outs[0] == outs[1] == outs[2] and outs[0].count(",") > 1000


def test_resampled_batch_vs_oracle(analysers, feats, oracle_lib):
    """Files at other rates go through the libresample restatement (SampleAnalyser.cpp:563-607)."""
    cases = [(synth.one_shot(900, 0.5, rate=96000, channels=2), 96000), (synth.one_shot(901, 0.6, rate=22050), 22050),
             (synth.one_shot(902, 0.7, rate=48000), 48000), (synth.one_shot(903, 0.5), 44100),
             (synth.one_shot(904, 0.3, rate=8000), 8000), (synth.one_shot(900, 0.5, rate=96000, channels=2), 96000)]
    an = analysers(1024)
    b = an.batch([c[0] for c in cases], [c[1] for c in cases]).run()
    for i, (p, r) in enumerate(cases):
        data, off, pk, rms = oracle_lib.condition(p, src_rate=r)
        got = b.conditioned(i)
        assert got.shape == data.shape
        assert np.array_equal(got, data), "resampled + conditioned signal of case %d differs" % i
        want = oracle_lib.analyze(p, src_rate=r, file_size=44 + p.size * 2)
        check(b.result(i), want, feats)
    b.free()


@pytest.mark.parametrize("n", [256, 1024, 2048])
def test_fft_core_known_answers(analysers, n):
    """The register-blocked FFT core vs numpy: random data, an impulse, a pure tone, and the round trip
    conj(FFT(conj(FFT(x)))) / n == x (the properties TestFourier.cpp:16-83 checks on the reference FFT)."""
    rng = np.random.default_rng(n)
    x = rng.standard_normal((5, n)) + 1j * rng.standard_normal((5, n))
    x[1] = 0; x[1, 3] = 1.0
    x[2] = np.exp(2j * np.pi * 17 * np.arange(n) / n)
    x[3] = rng.standard_normal(n)                      # real input
    an = analysers(1024)
    X = an.debug_fft(x)
    ref = np.fft.fft(x, axis=1)
    assert np.max(np.abs(X - ref)) <= 1e-11 * max(1.0, np.max(np.abs(ref)))
    back = np.conj(an.debug_fft(np.conj(X))) / n
    assert np.max(np.abs(back - x)) <= 1e-12 * max(1.0, np.max(np.abs(x)))


def test_long_resampled_stereo_file_vs_oracle(analysers, feats, oracle_lib):
    """BASELINE config 5 in miniature: a 96 kHz stereo file longer than the 20 s analysis cap -- downmix,
    libresample-exact resampling of ~2.3 M frames (hundreds of resampler blocks), trim over the whole file,
    capped analysis (F = 860, Fr = 6887)."""
    clip = synth.one_shot(31, 8.0, rate=96000, channels=2)
    pcm = np.ascontiguousarray(np.tile(clip, (4, 1)))                   # 32 s
    an = analysers(1024)
    b = an.batch([pcm], [96000]).run()
    data, off, pk, rms = oracle_lib.condition(pcm, src_rate=96000)
    got = b.conditioned(0)
    assert got.shape == data.shape and np.array_equal(got, data)
    r = b.result(0)
    assert (r.F, r.Fr) == (860, 6887)
    want = oracle_lib.analyze(pcm, src_rate=96000, file_size=44 + pcm.size * 2)
    check(r, want, feats, mdata=data)
    b.free()


@pytest.mark.parametrize("rate,seconds", [(48000, 6.0), (88200, 3.0), (32000, 4.0), (11025, 5.0), (192000, 2.0), (44101, 2.0),
                                          (96000, 12.0), (22050, 9.0), (16000, 3.0), (47999, 1.0)])
def test_resampler_bit_exact_across_rates(analysers, oracle_lib, rate, seconds):
    """k_resample against the oracle's libresample restatement, bit for bit, over rate ratios p / q that exercise the
    shared-memory coefficient rows (q = 147, 441, 1, 2, 4), several blocks per file, and ratios whose rows do not fit
    (44101 / 44100, 47999 / 44100: the literal look-up)."""
    pcm = synth.one_shot(1200 + rate % 97, seconds, rate=rate, channels=1 + (rate % 2))
    an = analysers(1024)
    b = an.batch([pcm], [rate]).run()
    data, off, pk, rms = oracle_lib.condition(pcm, src_rate=rate)
    got = b.conditioned(0)
    b.free()
    assert got.shape == data.shape
    assert np.array_equal(got, data)
