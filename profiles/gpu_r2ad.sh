#!/bin/bash
# Round 2, call AD: the single-tick pitch rule -- its GPU regression test, the sweep that found it (seed 15000) and two fresh seeds.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "single_tick or impulse or batch_vs_oracle" 2>&1 | tail -3 > gpurun_out/r2ad_tests.log; cat gpurun_out/r2ad_tests.log
for spec in "160 1024 15000" "320 1024 16000" "160 512 17000"; do set -- $spec
  (timeout 900 python profiles/parity_sweep.py $1 $2 $3 2>&1 | tail -4) > gpurun_out/r2ad_sweep_$2_$3.log; cat gpurun_out/r2ad_sweep_$2_$3.log
done
