#!/bin/bash
# Round 2, call A: GPU parity tests (new subset / shape-cache / live-batch tests), the default bench line (full workload)
# and the config-2 line.  Outputs land in gpurun_out/ ($1 = tag).
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
(timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -25) > gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/${TAG}_bench_full.json 2> gpurun_out/${TAG}_bench_full.err; tail -c 6000 gpurun_out/${TAG}_bench_full.json; tail -5 gpurun_out/${TAG}_bench_full.err
timeout 600 python bench.py --workload config2 --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_config2.json 2> gpurun_out/${TAG}_bench_config2.err; tail -c 3000 gpurun_out/${TAG}_bench_config2.json
