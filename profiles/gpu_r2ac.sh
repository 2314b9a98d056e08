#!/bin/bash
# Round 2, call AC: k_bands_select histogram counts aggregated per digit (match.any): bands parity tests, A/B time, sweep.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bands or batch_vs_oracle or golden or subset" 2>&1 | tail -3 > gpurun_out/r2ac_tests.log; cat gpurun_out/r2ac_tests.log
VT_MIXED=1 timeout 300 python profiles/variant_time.py 4000 1024 all > gpurun_out/r2ac_variant.log 2>&1; tail -1 gpurun_out/r2ac_variant.log
(timeout 900 python profiles/parity_sweep.py 160 1024 15000 2>&1 | tail -3) > gpurun_out/r2ac_sweep_1024.log; cat gpurun_out/r2ac_sweep_1024.log
