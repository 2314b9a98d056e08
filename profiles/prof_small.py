"""Small full-set batch for ncu: two warm computes, then ONE compute inside the NVTX range "prof"
(ncu --nvtx --nvtx-include "prof/" captures just that one)."""
import os, sys, numpy as np
sys.path.insert(0, '.')
import torch
from afec_b200 import api, synth
hop = int(os.environ.get("PROF_HOP", "1024"))
feats = api.FEAT_SPECTRAL if os.environ.get("PROF_FEATS") == "spectral" else api.FEAT_ALL
rate = 44100
if os.environ.get("PROF_LONG"):          # one 10-minute 96 kHz stereo file: the conditioning kernels of BASELINE configs[4]
    rate = 96000
    clip = synth.one_shot(7, 30.0, rate=rate, channels=2)
    pcms = [np.ascontiguousarray(np.tile(clip, (20, 1)))]
elif os.environ.get("PROF_MIXED"):       # the mixed-length (0.5-30 s) corpus of BASELINE configs[3], as in bench.py --workload full
    pcms = synth.tiled_corpus(int(os.environ.get("PROF_FILES", "128")), 64, seconds=30.0, seed0=0, min_seconds=0.5)
else:
    pcms = synth.tiled_corpus(int(os.environ.get("PROF_FILES", "400")), 16, seconds=3.0, seed0=0)
an = api.SampleAnalyser(44100, 2048, hop, features=feats)
b = an.batch(pcms, [rate]*len(pcms))
b.upload(); b.compute(); b.compute(); b.sync()
torch.cuda.nvtx.range_push("prof")
b.compute(); b.sync()
torch.cuda.nvtx.range_pop()
print(b.timings())
b.free(); an.close()
