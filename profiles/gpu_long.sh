#!/bin/bash
# long-file workload (BASELINE configs[4]): GPU parity tests, then the whole-file path and the part path on one GPU,
# and one `ncu --set full` of the default bench workload's kernels (hop 512, spectral subset) for the traffic figure
TAG=${1:-x}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15) > gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py --workload long --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_long_whole.json 2> gpurun_out/${TAG}_bench_long_whole.err; tail -c 2500 gpurun_out/${TAG}_bench_long_whole.json; tail -3 gpurun_out/${TAG}_bench_long_whole.err
timeout 600 python bench.py --workload long --parts 4 --files 4 --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_long_parts.json 2> gpurun_out/${TAG}_bench_long_parts.err; tail -c 2500 gpurun_out/${TAG}_bench_long_parts.json; tail -3 gpurun_out/${TAG}_bench_long_parts.err
PROF_HOP=512 PROF_FEATS=spectral AFX_SINGLE_STREAM=1 timeout 600 ncu --nvtx --nvtx-include "prof/" --set full --clock-control none --import-source on -k regex:'k_(spectrum|flux)' -f -o gpurun_out/${TAG}_prof_config2 python profiles/prof_small.py > gpurun_out/${TAG}_prof_config2.log 2>&1
ncu -i gpurun_out/${TAG}_prof_config2.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_config2_raw.csv 2>/dev/null
PROF_LONG=1 AFX_SINGLE_STREAM=1 timeout 600 ncu --nvtx --nvtx-include "prof/" --set full --clock-control none --import-source on -k regex:'k_(downmix|resample|reduce|trim|eff)' -f -o gpurun_out/${TAG}_prof_long python profiles/prof_small.py > gpurun_out/${TAG}_prof_long.log 2>&1
ncu -i gpurun_out/${TAG}_prof_long.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_long_raw.csv 2>/dev/null
ls -la gpurun_out; du -sh gpurun_out
