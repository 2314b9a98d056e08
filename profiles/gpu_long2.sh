#!/bin/bash
TAG=${1:-x}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15) > gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py --workload long --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_long_whole.json 2> gpurun_out/${TAG}_bench_long_whole.err; tail -c 1200 gpurun_out/${TAG}_bench_long_whole.json; tail -3 gpurun_out/${TAG}_bench_long_whole.err
timeout 600 python bench.py --workload long --parts 4 --files 4 --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_long_parts.json 2> gpurun_out/${TAG}_bench_long_parts.err; tail -c 600 gpurun_out/${TAG}_bench_long_parts.json; tail -3 gpurun_out/${TAG}_bench_long_parts.err
PROF_LONG=1 AFX_SINGLE_STREAM=1 timeout 600 ncu --nvtx --nvtx-include "prof/" --set full --clock-control none --import-source on -k regex:'k_(downmix|resample)' -f -o gpurun_out/${TAG}_prof_long python profiles/prof_small.py > gpurun_out/${TAG}_prof_long.log 2>&1
ncu -i gpurun_out/${TAG}_prof_long.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_long_raw.csv 2>/dev/null
ls -la gpurun_out; du -sh gpurun_out
