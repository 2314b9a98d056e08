#!/bin/bash
# Round 2, call U: k_rhythm_pipe (whitening + onset functions as a per-file pipeline): bit-equality / parity tests, A/B times
# against the three split kernels on the mixed corpus, racecheck + memcheck of the new kernel, hop-1024 sweep.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "rhythm or batch_vs_oracle or golden" 2>&1 | tail -6 > gpurun_out/r2u_tests.log; cat gpurun_out/r2u_tests.log
for v in 0 1; do VT_MIXED=1 AFX_RHYTHM_PIPE=$v timeout 300 python profiles/variant_time.py 4000 1024 all > gpurun_out/r2u_variant_pipe_$v.log 2>&1; tail -1 gpurun_out/r2u_variant_pipe_$v.log; done
VT_BENCH=1 timeout 300 python profiles/variant_time.py 12500 1024 all > gpurun_out/r2u_variant_bench.log 2>&1; tail -1 gpurun_out/r2u_variant_bench.log
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
AFX_RHYTHM_PIPE=1 timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --kernel-name regex:k_rhythm_pipe --error-exitcode 9 python /tmp/san.py > gpurun_out/r2u_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r2u_racecheck.log
grep -E "Race reported|ERROR SUMMARY|RACECHECK SUMMARY|rc=" gpurun_out/r2u_racecheck.log | sort | uniq -c | head -8
AFX_RHYTHM_PIPE=1 timeout 900 compute-sanitizer --tool memcheck --kernel-name regex:"k_rhythm_pipe|k_autocorr" --error-exitcode 9 python /tmp/san.py > gpurun_out/r2u_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2u_memcheck.log
tail -3 gpurun_out/r2u_memcheck.log
(timeout 900 python profiles/parity_sweep.py 320 1024 13000 2>&1 | tail -6) > gpurun_out/r2u_sweep_1024.log; cat gpurun_out/r2u_sweep_1024.log
