#!/bin/bash
# Round 2, call K: parity sweeps (hop 1024, two seeds) with the block-sharing pitch kernel + the whole GPU suite.
mkdir -p gpurun_out
(timeout 900 python profiles/parity_sweep.py 320 1024 7000 2>&1 | tail -6) > gpurun_out/r2k_sweep_1024a.log; cat gpurun_out/r2k_sweep_1024a.log
(timeout 900 python profiles/parity_sweep.py 320 1024 11000 2>&1 | tail -6) > gpurun_out/r2k_sweep_1024b.log; cat gpurun_out/r2k_sweep_1024b.log
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r2k_tests.log; cat gpurun_out/r2k_tests.log
