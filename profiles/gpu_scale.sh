#!/bin/bash
# bench.py under torchrun on N GPUs of one box ($1 = N, $2 = tag): the default (full) workload; also records host memory,
# CPU count and the PCIe / NUMA topology for the end-to-end scaling analysis.
N=${1:-2}; TAG=${2:-r02f_scale}
mkdir -p gpurun_out
(free -g; nproc; lscpu | grep -E "Model name|Socket|NUMA|Thread|Core"; nvidia-smi topo -m) > gpurun_out/${TAG}_host_n$N.txt 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/${TAG}_bench_full_n$N.json 2> gpurun_out/${TAG}_bench_full_n$N.err
tail -c 1200 gpurun_out/${TAG}_bench_full_n$N.json; tail -5 gpurun_out/${TAG}_bench_full_n$N.err; head -12 gpurun_out/${TAG}_host_n$N.txt
