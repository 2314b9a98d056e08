#!/bin/bash
# Round 2, call D: extension tests (tcgen05 with 8 accumulator sets), FP32 vs FP64 autocorrelation, per-launch times of the
# mixed corpus, `ncu --set full` of the new bands / autocorrelation kernels, random-corpus parity sweep.
TAG=${1:-r02d}
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_ext.py -m gpu -q -s --timeout 120 2>&1 | tail -30) > gpurun_out/${TAG}_pytest_ext.log; cat gpurun_out/${TAG}_pytest_ext.log
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 2>&1 | tail -15) > gpurun_out/${TAG}_pytest_parity.log; cat gpurun_out/${TAG}_pytest_parity.log
for v in 1 0; do VT_MIXED=1 AFX_AUTOCORR_FP64=$v timeout 300 python profiles/variant_time.py 4000 1024 all > gpurun_out/${TAG}_variant_acfp64_$v.log 2>&1; tail -2 gpurun_out/${TAG}_variant_acfp64_$v.log; done
export AFX_SINGLE_STREAM=1
PROF_MIXED=1 PROF_FILES=2000 timeout 600 ncu --nvtx --nvtx-include "prof/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_mixed2000.csv python profiles/prof_small.py > /dev/null 2>&1
timeout 900 ncu --nvtx --nvtx-include "prof/" --set full --clock-control none --import-source on -k regex:'k_(bands|autocorr)' -f -o gpurun_out/${TAG}_bands python profiles/prof_small.py > gpurun_out/${TAG}_bands.log 2>&1
ncu -i gpurun_out/${TAG}_bands.ncu-rep --page raw --csv > gpurun_out/${TAG}_bands_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_bands.ncu-rep --page source --csv -k regex:k_bands_lane > gpurun_out/${TAG}_bands_lane_source.csv 2>/dev/null
if [ $(stat -c %s gpurun_out/${TAG}_bands.ncu-rep) -gt 30000000 ]; then rm gpurun_out/${TAG}_bands.ncu-rep; fi
unset AFX_SINGLE_STREAM
(timeout 900 python profiles/parity_sweep.py 320 1024 7000 2>&1 | tail -12) > gpurun_out/${TAG}_sweep_1024.log; cat gpurun_out/${TAG}_sweep_1024.log
(timeout 600 python profiles/parity_sweep.py 160 512 8000 2>&1 | tail -12) > gpurun_out/${TAG}_sweep_512.log; cat gpurun_out/${TAG}_sweep_512.log
ls -la gpurun_out | tail -12
