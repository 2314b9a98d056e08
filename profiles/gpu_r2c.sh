#!/bin/bash
# Round 2, call C: the extension tests first (tcgen05 kernel: bounded waits, per-test timeout), then all GPU tests, the
# per-group times with the lane-per-frame bands kernels, the default bench line.
TAG=${1:-r02c}
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_ext.py -m gpu -q --timeout 120 2>&1 | tail -40) > gpurun_out/${TAG}_pytest_ext.log; cat gpurun_out/${TAG}_pytest_ext.log
(timeout 1500 python -m pytest tests -m gpu -q --timeout 300 --deselect tests/test_gpu_ext.py 2>&1 | tail -60) > gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_pytest_gpu.log
VT_MIXED=1 timeout 300 python profiles/variant_time.py 4000 1024 all > gpurun_out/${TAG}_variant.log 2>&1; tail -4 gpurun_out/${TAG}_variant.log
timeout 900 python bench.py --steps 8 --warmup 3 > gpurun_out/${TAG}_bench_full.json 2> gpurun_out/${TAG}_bench_full.err; tail -c 5000 gpurun_out/${TAG}_bench_full.json; tail -5 gpurun_out/${TAG}_bench_full.err
