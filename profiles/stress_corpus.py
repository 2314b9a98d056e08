"""Edge-material corpus (tests / sweeps only): the signals a sample library really holds next to decaying one-shots -- steady
tones on and off bin centres, square / saw / impulse trains, DC, full-scale and clipped noise, Nyquist tones, chirps, one
impulse in silence, silence next to bursts, phase-inverted stereo (a mono sum of zero), few-LSB material.

    python profiles/stress_corpus.py self [hop]     CPU: the oracle with its two FFT variants, under tests/parity.py's rules
    python profiles/stress_corpus.py gpu [hop]      CUDA path against the oracle, same rules"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np


def i16(x):
    return np.ascontiguousarray(np.clip(np.round(x), -32768, 32767).astype(np.int16))


def build(rate=44100):
    rng = np.random.default_rng(424242)
    out = []
    def add(name, x, r=rate):
        out.append((name, i16(x) if x.dtype != np.int16 else np.ascontiguousarray(x), r))
    n = int(1.5 * rate); t = np.arange(n) / rate
    for k in (1, 4, 21, 100, 511, 738, 1000):                      # steady tones on bin centres of the 2048-point transform
        add("sine_bin%d" % k, 20000 * np.sin(2 * np.pi * (k * rate / 2048.0) * t))
    for f in (55.0, 440.0, 997.3, 7919.1, 15000.0, 21000.0):      # ... and off them
        add("sine_%g" % f, 12000 * np.sin(2 * np.pi * f * t))
    add("sine_quiet", 20 * np.sin(2 * np.pi * 440.0 * t))
    add("sine_3lsb", 3 * np.sin(2 * np.pi * 300.0 * t))
    for p in (32, 100, 441, 1024):                                # square / saw / impulse trains with integer periods
        add("square_%d" % p, 15000 * np.where((np.arange(n) % p) < p // 2, 1.0, -1.0))
        add("saw_%d" % p, 15000 * ((np.arange(n) % p) / p * 2 - 1))
        x = np.zeros(n); x[::p] = 25000; add("impulses_%d" % p, x)
    x = np.zeros(n); x[n // 2] = 30000; add("one_impulse", x)
    x = np.zeros(n); x[5000] = 1; add("one_lsb_tick", x)
    add("dc", np.full(n, 1000.0))
    add("dc_noise", 1000.0 + rng.standard_normal(n) * 2)
    add("noise_full", rng.standard_normal(n) * 9000)
    add("noise_clipped", rng.standard_normal(n) * 60000)
    add("noise_lsb", rng.integers(-1, 2, n).astype(np.float64))
    add("nyquist", 10000 * np.where(np.arange(n) % 2 == 0, 1.0, -1.0))
    add("chirp", 15000 * np.sin(2 * np.pi * (50 * t + 0.5 * (18000 - 50) / t[-1] * t * t)))
    x = np.zeros(n); x[n // 2:] = rng.standard_normal(n - n // 2) * 8000; add("silence_then_noise", x)
    x = np.zeros(n); x[:n // 3] = 14000 * np.sin(2 * np.pi * 330 * t[:n // 3]); add("tone_then_silence", x)
    x = np.zeros(n); x[3000:3030] = 20000 * np.hanning(30); x[40000:40030] = -20000 * np.hanning(30); add("two_clicks", x)
    add("ramp", np.linspace(-30000, 30000, n))
    s = 12000 * np.sin(2 * np.pi * 220 * t) * np.exp(-t / 0.3)
    add("stereo_inverted", np.stack([s, -s], axis=1))            # mono sum exactly zero
    add("stereo_one_side", np.stack([s, np.zeros(n)], axis=1))
    add("full_scale_square", np.where((np.arange(n) % 64) < 32, 32767.0, -32768.0))
    add("decay_to_lsb", 30000 * np.sin(2 * np.pi * 500 * t) * np.exp(-t / 0.05))
    add("tone_48k", 12000 * np.sin(2 * np.pi * 1000.0 * np.arange(int(1.2 * 48000)) / 48000), 48000)
    add("square_22k", 12000 * np.where((np.arange(int(1.2 * 22050)) % 50) < 25, 1.0, -1.0), 22050)
    add("short_2049", 9000 * np.sin(2 * np.pi * 800 * np.arange(2049) / rate))
    add("short_700", 9000 * np.sin(2 * np.pi * 800 * np.arange(700) / rate))
    add("beats_120bpm", np.concatenate([np.concatenate([16000 * rng.standard_normal(800) * np.exp(-np.arange(800) / 150.0), np.zeros(22050 - 800)]) for _ in range(8)]))
    return out


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "self"
    hop = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    import parity
    from oracle import oracle
    oracle.build()
    files = build()
    bad = 0
    got = None
    if mode == "gpu":
        from afec_b200 import api
        an = api.SampleAnalyser(44100, 2048, hop, features=api.FEAT_ALL)
        got = an.analyze_pcm([p for _, p, _ in files], [r for _, _, r in files])
    t0 = time.time()
    for i, (name, p, r) in enumerate(files):
        want = oracle.analyze(p, src_rate=r, hop=hop, file_size=44 + p.size * 2)
        data = oracle.condition(p, src_rate=r)[0] if want.status == 0 else None
        if mode == "gpu":
            other = got[i]
        else:
            oracle.set_fft_variant(1)
            try:
                other = oracle.analyze(p, src_rate=r, hop=hop, file_size=44 + p.size * 2)
            finally:
                oracle.set_fft_variant(0)
        errs = parity.compare(other, want, mdata=data, hop=hop)
        if errs:
            bad += 1
            print("%-20s status %d F %d: %d mismatches; first: %s" % (name, want.status, want.F, len(errs), errs[:3]), flush=True)
    print("stress corpus (%s): %d files, hop %d, %.0f s; files with mismatches: %d" % (mode, len(files), hop, time.time() - t0, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
