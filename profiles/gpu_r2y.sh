#!/bin/bash
# Round 2, call Y ($1 = N GPUs): final code on N GPUs of one box -- the two-device adapter test, then bench.py under torchrun.
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_host.py -x -q -m gpu -k "two_devices" 2>&1 | tail -3 > gpurun_out/r02z_pytest_two_devices_n$N.log; cat gpurun_out/r02z_pytest_two_devices_n$N.log
bash profiles/gpu_scale.sh $N r02z_scale
