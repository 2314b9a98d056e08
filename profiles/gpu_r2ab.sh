#!/bin/bash
# Round 2, call AB: k_rhythm_back at 4 CTAs per SM (64 registers): rhythm parity tests, A/B time, sweep; then the per-launch time
# list of bench.py itself (the default command, 2 steps) under ncu -- shares only, a number printed under ncu is not a bench value.
mkdir -p gpurun_out
bash profiles/gpu_r2w.sh
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02z_launches_bench_py.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sink --no-parity-check > gpurun_out/r2ab_bench_under_ncu.log 2>&1
tail -c 200 gpurun_out/r2ab_bench_under_ncu.log; wc -l gpurun_out/r02z_launches_bench_py.csv
