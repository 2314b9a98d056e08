"""CPU only: randomised edge material (steady tones, square / saw / impulse trains with integer periods, DC + noise, LSB
noise, gated bursts, chirps; amplitudes from 2 LSB to clipping; 22.05 - 96 kHz; mono / stereo) through the oracle with its two
FFT variants, under the rules of tests/parity.py -- looks for frames the reference's own arithmetic does not determine and
that no rule names.    python profiles/edge_self_sweep.py [n_files] [hop] [seed0] [procs]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import multiprocessing as mp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
hop = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
seed0 = int(sys.argv[3]) if len(sys.argv) > 3 else 70000
procs = int(sys.argv[4]) if len(sys.argv) > 4 else (os.cpu_count() or 1)


def make(i):
    rng = np.random.default_rng(seed0 + i)
    rate = int(rng.choice([44100, 44100, 44100, 48000, 22050, 96000]))
    m = int(rate * float(np.exp(rng.uniform(np.log(0.1), np.log(4.0)))))
    t = np.arange(m) / rate
    amp = float(np.exp(rng.uniform(np.log(2.0), np.log(60000.0))))
    kind = str(rng.choice(["sine", "sine_bin", "square", "saw", "impulses", "dc_noise", "noise", "lsb_noise", "chirp", "gated", "two_tones"]))
    if kind == "sine":
        x = np.sin(2 * np.pi * float(np.exp(rng.uniform(np.log(20.0), np.log(0.49 * rate)))) * t)
    elif kind == "sine_bin":
        x = np.sin(2 * np.pi * (int(rng.integers(1, 1000)) * 44100 / 2048.0) * t + rng.uniform(0, 6.28))
    elif kind == "square":
        p = int(rng.integers(2, 2000)); x = np.where((np.arange(m) % p) < max(1, p // 2), 1.0, -1.0)
    elif kind == "saw":
        p = int(rng.integers(2, 2000)); x = (np.arange(m) % p) / p * 2 - 1
    elif kind == "impulses":
        p = int(rng.integers(8, 5000)); x = np.zeros(m); x[int(rng.integers(0, p))::p] = 1.0
    elif kind == "dc_noise":
        x = rng.uniform(-1, 1) + rng.standard_normal(m) * float(np.exp(rng.uniform(np.log(1e-4), np.log(0.3))))
    elif kind == "noise":
        x = rng.standard_normal(m) * 0.3
    elif kind == "lsb_noise":
        x = rng.integers(-2, 3, m).astype(np.float64); amp = 1.0
    elif kind == "chirp":
        f0, f1 = sorted(np.exp(rng.uniform(np.log(30.0), np.log(0.45 * rate), 2)))
        x = np.sin(2 * np.pi * (f0 * t + 0.5 * (f1 - f0) / max(t[-1], 1e-3) * t * t))
    elif kind == "gated":
        x = rng.standard_normal(m) * 0.3 * (np.sin(2 * np.pi * rng.uniform(0.5, 8.0) * t) > rng.uniform(-0.5, 0.9))
    else:
        x = 0.6 * np.sin(2 * np.pi * rng.uniform(50, 5000) * t) + 0.4 * np.sin(2 * np.pi * rng.uniform(50, 15000) * t)
    pcm = np.clip(np.round(x * amp), -32768, 32767).astype(np.int16)
    if rng.random() < 0.25:
        pcm = np.stack([pcm, np.roll(pcm, int(rng.integers(0, 50))) if rng.random() < 0.7 else -pcm], axis=1)
    return kind, rate, np.ascontiguousarray(pcm)


def work(i):
    import parity
    from oracle import oracle
    kind, rate, x = make(i)
    a = oracle.analyze(x, src_rate=rate, hop=hop, file_size=44 + x.size * 2)
    oracle.set_fft_variant(1)
    try:
        b = oracle.analyze(x, src_rate=rate, hop=hop, file_size=44 + x.size * 2)
    finally:
        oracle.set_fft_variant(0)
    data = oracle.condition(x, src_rate=rate)[0] if a.status == 0 else None
    errs = parity.compare(b, a, mdata=data, hop=hop)
    return i, kind, rate, x.shape, errs[:3], len(errs)


if __name__ == "__main__":
    from oracle import oracle
    oracle.build()
    t0 = time.time()
    bad = 0
    with mp.Pool(procs) as pool:
        for i, kind, rate, shape, errs, ne in pool.imap_unordered(work, range(n), chunksize=2):
            if ne:
                bad += 1
                print("file %d (seed %d, %s, rate %d, shape %s): %d mismatches; first: %s" % (i, seed0 + i, kind, rate, shape, ne, errs), flush=True)
    print("edge self-sweep: %d files, hop %d, seed0 %d, %.0f s; files where the two FFT variants disagree outside the rules: %d"
          % (n, hop, seed0, time.time() - t0, bad))
