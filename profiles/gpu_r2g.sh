#!/bin/bash
# Round 2, call G: the fused whitening + peak-count kernel (A/B against the split pair), all GPU tests, sweep, bench.
TAG=${1:-r02g}
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -12) > gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_pytest_gpu.log
for v in 1 0; do VT_MIXED=1 AFX_PEAKS_SPLIT=$v timeout 300 python profiles/variant_time.py 4000 1024 all > gpurun_out/${TAG}_variant_peaks_split_$v.log 2>&1; tail -1 gpurun_out/${TAG}_variant_peaks_split_$v.log; done
(timeout 900 python profiles/parity_sweep.py 320 1024 7000 2>&1 | tail -4) > gpurun_out/${TAG}_sweep_1024.log; cat gpurun_out/${TAG}_sweep_1024.log
(timeout 600 python profiles/parity_sweep.py 160 512 8000 2>&1 | tail -4) > gpurun_out/${TAG}_sweep_512.log; cat gpurun_out/${TAG}_sweep_512.log
timeout 900 python bench.py --steps 8 --warmup 3 --no-sink > gpurun_out/${TAG}_bench_full.json 2> gpurun_out/${TAG}_bench_full.err; tail -c 400 gpurun_out/${TAG}_bench_full.json; tail -3 gpurun_out/${TAG}_bench_full.err
