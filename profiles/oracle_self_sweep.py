"""CPU-only companion of parity_sweep.py: the oracle against ITSELF with its second FFT variant (same transform, different
rounding -- the situation of the reference built with another FFT back end, and of the CUDA path's FFTs) on the sweep's random
corpus, under the rules of tests/parity.py.  A file that fails here holds a frame the reference's own arithmetic does not
determine and that no rule names yet.    python profiles/oracle_self_sweep.py [n_files] [hop] [seed0] [procs]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import multiprocessing as mp

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
hop = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
seed0 = int(sys.argv[3]) if len(sys.argv) > 3 else 5000
procs = int(sys.argv[4]) if len(sys.argv) > 4 else (os.cpu_count() or 1)


def corpus():
    from afec_b200 import synth
    rng = np.random.default_rng(seed0)
    out = []
    for i in range(n):                       # the same draws as profiles/parity_sweep.py
        rate = int(rng.choice([44100, 44100, 44100, 48000, 22050, 96000]))
        ch = int(rng.choice([1, 1, 2]))
        sec = float(np.exp(rng.uniform(np.log(0.06), np.log(24.0))))
        pad = rng.random() < 0.15
        pad_s = rng.uniform(0.05, 1.5) if pad else 0.0
        both = (rng.random() < 0.5) if pad else False
        quiet = rng.random() < 0.1
        scale = rng.uniform(0.001, 0.05) if quiet else 1.0
        out.append((i, rate, ch, sec, pad_s, both, scale))
    return out


def work(spec):
    import parity
    from afec_b200 import synth
    from oracle import oracle
    i, rate, ch, sec, pad_s, both, scale = spec
    x = synth.one_shot(seed0 + i, sec, rate=rate, channels=ch)
    if pad_s:
        pad = np.zeros((int(rate * pad_s),) + x.shape[1:], dtype=x.dtype)
        x = np.concatenate([pad, x, pad]) if both else np.concatenate([x, pad])
    if scale != 1.0:
        x = (x.astype(np.float64) * scale).astype(np.int16)
    x = np.ascontiguousarray(x)
    a = oracle.analyze(x, src_rate=rate, hop=hop, file_size=44 + x.size * x.itemsize)
    oracle.set_fft_variant(1)
    try:
        b = oracle.analyze(x, src_rate=rate, hop=hop, file_size=44 + x.size * x.itemsize)
    finally:
        oracle.set_fft_variant(0)
    data = oracle.condition(x, src_rate=rate)[0] if a.status == 0 else None
    errs = parity.compare(b, a, mdata=data, hop=hop)
    return i, rate, x.shape, errs[:3], len(errs)


if __name__ == "__main__":
    from oracle import oracle
    oracle.build()
    t0 = time.time()
    bad = 0
    with mp.Pool(procs) as pool:
        for i, rate, shape, errs, ne in pool.imap_unordered(work, corpus(), chunksize=4):
            if ne:
                bad += 1
                print("file %d (seed %d, rate %d, shape %s): %d mismatches; first: %s" % (i, seed0 + i, rate, shape, ne, errs), flush=True)
    print("oracle self-sweep: %d files, hop %d, seed0 %d, %.0f s; files where the two FFT variants disagree outside the rules: %d"
          % (n, hop, seed0, time.time() - t0, bad))
