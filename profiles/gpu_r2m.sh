#!/bin/bash
# Round 2, call M: quick parity subset + per-group timing on the mixed corpus (A/B runs of kernel variants).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mixed or golden or group or subset or pitch or peaks" 2>&1 | tail -5 > gpurun_out/r2m_tests.log; cat gpurun_out/r2m_tests.log
for v in "AFX_PEAKS_PIPE=1" "AFX_PEAKS_PIPE=0"; do
env VT_MIXED=1 $v timeout 600 python profiles/variant_time.py 4000 1024 all 2>&1 | tail -2
done > gpurun_out/r2m_variants.log; cat gpurun_out/r2m_variants.log
