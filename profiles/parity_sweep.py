"""Wide parity sweep (not part of the test suite: minutes of CPU oracle time): N random files of mixed length, channel
count and sample rate through the CUDA path in ONE batch, every file against the CPU oracle with the rules of
tests/parity.py.  Prints the files that break a rule.   python profiles/parity_sweep.py [n_files] [hop] [seed0]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import parity
from afec_b200 import api, synth
from oracle import oracle

n = int(sys.argv[1]) if len(sys.argv) > 1 else 120
hop = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
seed0 = int(sys.argv[3]) if len(sys.argv) > 3 else 5000
rng = np.random.default_rng(seed0)
pcms, rates = [], []
for i in range(n):
    rate = int(rng.choice([44100, 44100, 44100, 48000, 22050, 96000]))
    ch = int(rng.choice([1, 1, 2]))
    sec = float(np.exp(rng.uniform(np.log(0.06), np.log(24.0))))
    x = synth.one_shot(seed0 + i, sec, rate=rate, channels=ch)
    if rng.random() < 0.15:                       # leading / trailing digital silence, quiet files
        pad = np.zeros((int(rate * rng.uniform(0.05, 1.5)),) + x.shape[1:], dtype=x.dtype)
        x = np.concatenate([pad, x, pad]) if rng.random() < 0.5 else np.concatenate([x, pad])
    if rng.random() < 0.1:
        x = (x.astype(np.float64) * rng.uniform(0.001, 0.05)).astype(np.int16)
    pcms.append(np.ascontiguousarray(x)); rates.append(rate)
oracle.build()
an = api.SampleAnalyser(44100, 2048, hop, features=api.FEAT_ALL)
t0 = time.time()
got = an.analyze_pcm(pcms, rates)
t1 = time.time()
bad = 0
for i, (g, p, r) in enumerate(zip(got, pcms, rates)):
    want = oracle.analyze(p, src_rate=r, hop=hop, file_size=44 + p.size * p.itemsize)
    data = oracle.condition(p, src_rate=r)[0] if want.status == 0 else None       # finds the frames the pitch rule excludes
    errs = parity.compare(g, want, mdata=data, hop=hop)
    if errs:
        bad += 1
        print("file %d (seed %d, rate %d, shape %s): %d mismatches; first: %s" % (i, seed0 + i, r, p.shape, len(errs), errs[:3]))
print("sweep: %d files, hop %d, %.1f s of audio; GPU %.2f s, oracle %.1f s; files with mismatches: %d"
      % (n, hop, sum(p.shape[0] / r for p, r in zip(pcms, rates)), t1 - t0, time.time() - t1, bad))
an.close()
sys.exit(1 if bad else 0)
