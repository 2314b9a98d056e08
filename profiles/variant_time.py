"""Per-kernel-group device times of one batch (AFX_DEBUG_KERNEL_TIMES=1: single stream, events around each group),
for A/B runs of kernel variants selected by environment variables.  Usage:
    AFX_DEBUG_KERNEL_TIMES=1 [AFX_SPEC_VARIANT=..] python profiles/variant_time.py [files] [hop] [spectral|all]"""
import os, sys
sys.path.insert(0, '.')
os.environ.setdefault("AFX_DEBUG_KERNEL_TIMES", "1")
import numpy as np
from afec_b200 import api, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
hop = int(sys.argv[2]) if len(sys.argv) > 2 else 512
feats = api.FEAT_SPECTRAL if (len(sys.argv) > 3 and sys.argv[3] == "spectral") else api.FEAT_ALL
if os.environ.get("VT_BENCH"):     # exactly bench.py's default corpus (N = 1)
    base = synth.corpus(64, 30.0, seed0=1000, min_seconds=0.5)
    pcms = [base[i % 64] for i in range(n)]
elif os.environ.get("VT_MIXED"):     # mixed-length (0.5-30 s) corpus, the shape of bench.py --workload full
    pcms = synth.tiled_corpus(n, 64, seconds=30.0, seed0=0, min_seconds=0.5)
else:
    pcms = synth.tiled_corpus(n, 16, seconds=3.0, seed0=0)
an = api.SampleAnalyser(44100, 2048, hop, features=feats)
b = an.batch(pcms, [44100] * len(pcms))
b.upload()
for _ in range(3):
    b.compute(); b.sync()
acc = {}
reps = 5
for _ in range(reps):
    b.compute(); b.sync()
    for k, ms in b.kernel_times():
        acc[k] = acc.get(k, 0.0) + ms
b.download(); b.sync()
frames = max(1, b.counters()["main_frames"])
print("files=%d hop=%d frames=%d variant=%s" % (n, hop, frames, {k: v for k, v in os.environ.items() if k.startswith("AFX_") and k != "AFX_DEBUG_KERNEL_TIMES"}))
print("  " + "  ".join("%s=%.3fms (%.2f ns/frame)" % (k, v / reps, v / reps * 1e6 / frames) for k, v in acc.items()))
b.free(); an.close()
