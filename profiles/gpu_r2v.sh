#!/bin/bash
# Round 2, call V: k_rhythm_pipe with the t - 2 phases preloaded: A/B times, ncu --set full with source counters, sanitizers.
mkdir -p gpurun_out
for v in 0 1; do VT_MIXED=1 AFX_RHYTHM_PIPE=$v timeout 300 python profiles/variant_time.py 4000 1024 all > gpurun_out/r2v_variant_pipe_$v.log 2>&1; tail -1 gpurun_out/r2v_variant_pipe_$v.log; done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "rhythm" 2>&1 | tail -3
bash profiles/gpu_r2l.sh 'k_rhythm_(pipe|polar)' > /dev/null 2>&1
for f in gpurun_out/r2l*; do mv $f ${f/r2l/r2v_ncu}; done
cat gpurun_out/r2v_ncu_summary.txt | head -64
cat > /tmp/san.py <<'P'
import sys, numpy as np
sys.path.insert(0, '.')
from afec_b200 import api, synth
pcms = [synth.one_shot(100 + i, 0.05 + 0.23 * i) for i in range(10)] + [np.zeros(30001, dtype=np.int16), synth.one_shot(122, 0.03), synth.one_shot(124, 9.0)]
an = api.SampleAnalyser(44100, 2048, 1024, features=api.FEAT_ALL)
r = an.analyze_pcm(pcms, [44100] * len(pcms))
print([x.status for x in r], sum(x.Fr for x in r))
an.close()
P
AFX_RHYTHM_PIPE=1 timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --kernel-name kns=k_rhythm_pipe --error-exitcode 9 python /tmp/san.py > gpurun_out/r2v_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r2v_racecheck.log
grep -E "Race reported|ERROR SUMMARY|RACECHECK SUMMARY|rc=" gpurun_out/r2v_racecheck.log | sort | uniq -c | head -8
AFX_RHYTHM_PIPE=1 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san.py > gpurun_out/r2v_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2v_memcheck.log
tail -3 gpurun_out/r2v_memcheck.log
