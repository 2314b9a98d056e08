#!/bin/bash
# Round 2, call B: all GPU tests (no -x: every failure is listed), the fused-vs-split rhythm timing, the crawler end to end.
TAG=${1:-r02b}
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -60) > gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_pytest_gpu.log
for m in 0 1; do VT_MIXED=1 AFX_RHYTHM_FUSED=$m timeout 300 python profiles/variant_time.py 4000 1024 all > gpurun_out/${TAG}_variant_fused$m.log 2>&1; tail -12 gpurun_out/${TAG}_variant_fused$m.log; done
