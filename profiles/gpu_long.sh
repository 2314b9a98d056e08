#!/bin/bash
# long-file workload (BASELINE configs[4]): GPU parity tests, then the whole-file path and the part path on one GPU
TAG=${1:-x}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15) > gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py --workload long --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_long_whole.json 2> gpurun_out/${TAG}_bench_long_whole.err; tail -c 2500 gpurun_out/${TAG}_bench_long_whole.json; tail -3 gpurun_out/${TAG}_bench_long_whole.err
timeout 600 python bench.py --workload long --parts 4 --files 2 --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_long_parts.json 2> gpurun_out/${TAG}_bench_long_parts.err; tail -c 2500 gpurun_out/${TAG}_bench_long_parts.json; tail -3 gpurun_out/${TAG}_bench_long_parts.err
