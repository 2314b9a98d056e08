#!/bin/bash
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload long --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_long_n2.json 2> gpurun_out/${TAG}_bench_long_n2.err; tail -c 1500 gpurun_out/${TAG}_bench_long_n2.json; tail -5 gpurun_out/${TAG}_bench_long_n2.err
timeout 600 python bench.py --workload long --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_bench_long_whole.json 2> gpurun_out/${TAG}_bench_long_whole.err; tail -c 1200 gpurun_out/${TAG}_bench_long_whole.json
