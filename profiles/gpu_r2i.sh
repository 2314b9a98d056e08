#!/bin/bash
# Round 2, call I: the block-sharing pitch kernel (k_pitch_hop) -- parity, then A/B timing against the general kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pitch or mixed or golden or group" 2>&1 | tail -15 > gpurun_out/r2i_tests.log; cat gpurun_out/r2i_tests.log
for v in "AFX_PITCH_NG=4"; do
  env VT_MIXED=1 $v timeout 600 python profiles/variant_time.py 4000 1024 all 2>&1 | tail -2
done > gpurun_out/r2i_variants.log 2>&1; cat gpurun_out/r2i_variants.log
