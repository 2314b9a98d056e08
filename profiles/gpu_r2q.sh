#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_longfile.py -x -q -m gpu -k "condition or resampl or long or golden or formats or raw" 2>&1 | tail -3
VT_BENCH=1 timeout 900 python profiles/variant_time.py 12500 1024 all 2>&1 | tail -1
