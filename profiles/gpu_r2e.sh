#!/bin/bash
# Round 2, call E: parity tests with the segment-walk bands kernel, the radix select and the move-free FP32
# autocorrelation; per-group and per-launch times; random-corpus parity sweep; N = 2 bench (needs --gpus 2).
TAG=${1:-r02e}
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -25) > gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_pytest_gpu.log
VT_MIXED=1 timeout 300 python profiles/variant_time.py 4000 1024 all > gpurun_out/${TAG}_variant.log 2>&1; tail -2 gpurun_out/${TAG}_variant.log
AFX_SINGLE_STREAM=1 PROF_MIXED=1 PROF_FILES=2000 timeout 600 ncu --nvtx --nvtx-include "prof/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_mixed2000.csv python profiles/prof_small.py > /dev/null 2>&1
(timeout 900 python profiles/parity_sweep.py 320 1024 7000 2>&1 | tail -8) > gpurun_out/${TAG}_sweep_1024.log; cat gpurun_out/${TAG}_sweep_1024.log
(timeout 600 python profiles/parity_sweep.py 160 512 8000 2>&1 | tail -8) > gpurun_out/${TAG}_sweep_512.log; cat gpurun_out/${TAG}_sweep_512.log
(timeout 600 python profiles/parity_sweep.py 160 768 9000 2>&1 | tail -8) > gpurun_out/${TAG}_sweep_768.log; cat gpurun_out/${TAG}_sweep_768.log
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/${TAG}_bench_full.json 2> gpurun_out/${TAG}_bench_full.err; tail -c 1500 gpurun_out/${TAG}_bench_full.json; tail -3 gpurun_out/${TAG}_bench_full.err
