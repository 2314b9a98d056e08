#!/bin/bash
# Round 2, call S: compute-sanitizer (memcheck, racecheck on shared memory) over a small mixed batch that exercises every kernel
# incl. the packed-offset vector downmix, stereo, resampling, the peaks pipeline (forced), high-level / pack / extension stages.
mkdir -p gpurun_out
cat > /tmp/san.py <<'P'
import sys, os, numpy as np
sys.path.insert(0, '.')
from afec_b200 import api, synth
pcms = [synth.one_shot(100 + i, 0.2 + 0.37 * i) for i in range(10)]
pcms += [synth.one_shot(120, 1.1, channels=2), synth.one_shot(121, 0.7, rate=48000), np.zeros(30001, dtype=np.int16), synth.one_shot(122, 0.03),
         synth.one_shot(123, 2.3, rate=22050, channels=2), np.zeros((0,), dtype=np.int16), synth.one_shot(124, 21.0)]
rates = [44100] * 10 + [44100, 48000, 44100, 44100, 22050, 44100, 44100]
feats = api.FEAT_ALL | api.FEAT_HIGHLEVEL | api.FEAT_PACK | api.FEAT_EXT_MELCHROMA
for hop in (1024, 512):
    an = api.SampleAnalyser(44100, 2048, hop, features=feats)
    r = an.analyze_pcm(pcms, rates)
    print(hop, [x.status for x in r], sum(x.F for x in r))
    an.close()
# files packed back to back at odd sample offsets in one host buffer (the adapter's / the bench's layout)
arena = np.concatenate([np.zeros(3, dtype=np.int16)] + [p.reshape(-1) for p in pcms[:10]])
an = api.SampleAnalyser(44100, 2048, 1024, features=api.FEAT_ALL)
off = 3; files = []
for p in pcms[:10]:
    files.append(api.AfxFile(arena.ctypes.data + 2 * off, p.shape[0], 1, 44100, api.AFX_PCM_I16, 16, 44 + 2 * p.size)); off += p.size
b = an.batch_from_descriptors(files, keepalive=[arena])
b.run(); print("packed", [b.result(i).status for i in range(len(files))]); b.free()
an.close()
P
AFX_PEAKS_PIPE=1 timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san.py > gpurun_out/r2s_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2s_memcheck.log
tail -12 gpurun_out/r2s_memcheck.log
AFX_PEAKS_PIPE=1 timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 python /tmp/san.py > gpurun_out/r2s_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r2s_racecheck.log
grep -E "Race reported|ERROR SUMMARY|RACECHECK SUMMARY|rc=" gpurun_out/r2s_racecheck.log | sort | uniq -c | head -20
