#!/bin/bash
# Profile-only gpurun call: GPU parity tests + launch list + `ncu --set full` of one compute (NVTX range "prof").
TAG=${1:-x}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15) > gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_pytest_gpu.log
AFX_SINGLE_STREAM=1 timeout 600 ncu --nvtx --nvtx-include "prof/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python profiles/prof_small.py > gpurun_out/${TAG}_launches.log 2>&1
AFX_SINGLE_STREAM=1 timeout 900 ncu --nvtx --nvtx-include "prof/" --set full --clock-control none --import-source on -k regex:'k_(spectrum|bands|pitch|autocorr|rhythm|peaks|whiten|stats|flux)' -f -o gpurun_out/${TAG}_prof python profiles/prof_small.py > gpurun_out/${TAG}_prof.log 2>&1
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_raw.csv 2>/dev/null
if [ $(stat -c %s gpurun_out/${TAG}_prof.ncu-rep) -gt 45000000 ]; then rm gpurun_out/${TAG}_prof.ncu-rep; fi
ls -la gpurun_out; du -sh gpurun_out
