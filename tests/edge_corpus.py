"""Edge-material corpus (tests / sweeps only): the signals a sample library really holds next to decaying one-shots -- steady
tones on and off bin centres, square / saw / impulse trains, DC, full-scale and clipped noise, Nyquist tones, chirps, one
impulse in silence, silence next to bursts, phase-inverted stereo (a mono sum of zero), few-LSB material.  build() returns
[(name, int16 PCM, rate)]; profiles/stress_corpus.py runs it through the oracle's two FFT variants (CPU) or the CUDA path,
tests/test_gpu_parity.py::test_edge_material_corpus through the CUDA path."""
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
