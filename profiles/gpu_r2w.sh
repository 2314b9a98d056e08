#!/bin/bash
# Round 2, call W: atan2_phase with the degree-6 interpolant and one Newton step: rhythm / golden / batch parity tests, A/B time
# on the mixed corpus, hop-1024 sweep.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "rhythm or batch_vs_oracle or golden" 2>&1 | tail -4 > gpurun_out/r2w_tests.log; cat gpurun_out/r2w_tests.log
VT_MIXED=1 timeout 300 python profiles/variant_time.py 4000 1024 all > gpurun_out/r2w_variant.log 2>&1; tail -1 gpurun_out/r2w_variant.log
(timeout 900 python profiles/parity_sweep.py 320 1024 14000 2>&1 | tail -4) > gpurun_out/r2w_sweep_1024.log; cat gpurun_out/r2w_sweep_1024.log
