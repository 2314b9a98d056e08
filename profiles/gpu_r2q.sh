#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mixed or golden or group or subset or autocorr" 2>&1 | tail -3
VT_MIXED=1 timeout 900 python profiles/variant_time.py 4000 1024 all 2>&1 | tail -1
