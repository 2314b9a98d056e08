#!/bin/bash
# Round 2, call F: FP32 autocorrelation with 17 lags per lane (A/B against FP64), parity tests + sweep, ncu of the bands
# kernels (segment walk / radix select) and the autocorrelation, bench line.
TAG=${1:-r02f}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 2>&1 | tail -8) > gpurun_out/${TAG}_pytest_parity.log; cat gpurun_out/${TAG}_pytest_parity.log
for v in 1 0; do VT_MIXED=1 AFX_AUTOCORR_FP64=$v timeout 300 python profiles/variant_time.py 4000 1024 all > gpurun_out/${TAG}_variant_acfp64_$v.log 2>&1; tail -1 gpurun_out/${TAG}_variant_acfp64_$v.log; done
(timeout 900 python profiles/parity_sweep.py 320 1024 7000 2>&1 | tail -4) > gpurun_out/${TAG}_sweep_1024.log; cat gpurun_out/${TAG}_sweep_1024.log
export AFX_SINGLE_STREAM=1
PROF_MIXED=1 PROF_FILES=1000 timeout 900 ncu --nvtx --nvtx-include "prof/" --set full --clock-control none --import-source on -k regex:'k_(bands|autocorr)' -f -o gpurun_out/${TAG}_bands python profiles/prof_small.py > gpurun_out/${TAG}_bands.log 2>&1
ncu -i gpurun_out/${TAG}_bands.ncu-rep --page raw --csv > gpurun_out/${TAG}_bands_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_bands.ncu-rep --page source --csv -k regex:k_bands_lane > gpurun_out/${TAG}_bands_lane_source.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_bands.ncu-rep --page source --csv -k regex:k_bands_select > gpurun_out/${TAG}_bands_select_source.csv 2>/dev/null
rm -f gpurun_out/${TAG}_bands.ncu-rep
unset AFX_SINGLE_STREAM
timeout 900 python bench.py --steps 8 --warmup 3 --no-sink > gpurun_out/${TAG}_bench_full.json 2> gpurun_out/${TAG}_bench_full.err; tail -c 600 gpurun_out/${TAG}_bench_full.json; tail -3 gpurun_out/${TAG}_bench_full.err
