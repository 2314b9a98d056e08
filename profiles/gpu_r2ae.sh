#!/bin/bash
# Round 2, call AE: the tick-frame pitch rule on the GPU: regression tests, the sweep that found the two-tick frame, one fresh seed.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ticks or impulse or large_batch" 2>&1 | tail -3 > gpurun_out/r2ae_tests.log; cat gpurun_out/r2ae_tests.log
for spec in "320 1024 16000" "200 1024 18000"; do set -- $spec
  (timeout 900 python profiles/parity_sweep.py $1 $2 $3 2>&1 | tail -4) > gpurun_out/r2ae_sweep_$2_$3.log; cat gpurun_out/r2ae_sweep_$2_$3.log
done
