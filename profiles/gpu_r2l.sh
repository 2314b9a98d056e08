#!/bin/bash
# Round 2, call L: ncu --set full with source counters for the spectrum, bands and rhythm front-end kernels (mixed corpus).
mkdir -p gpurun_out
export AFX_SINGLE_STREAM=1
PROF_MIXED=1 PROF_FILES=400 timeout 900 ncu --nvtx --nvtx-include "prof/" --set full --clock-control none --import-source on -k regex:'k_(spectrum|bands_lane|bands_select|rhythm_polar|rhythm_odf|rhythm_back|stats)' -f -o gpurun_out/r2l python profiles/prof_small.py > gpurun_out/r2l.log 2>&1
ncu -i gpurun_out/r2l.ncu-rep --page raw --csv > gpurun_out/r2l_raw.csv 2>/dev/null
for k in k_spectrum k_bands_lane k_bands_select k_rhythm_polar k_rhythm_odf k_rhythm_back k_stats; do
  ncu -i gpurun_out/r2l.ncu-rep --page source --csv -k regex:$k > gpurun_out/r2l_source_$k.csv 2>/dev/null
done
python profiles/ncu_summary.py gpurun_out/r2l_raw.csv > gpurun_out/r2l_summary.txt
rm -f gpurun_out/r2l.ncu-rep
ls -la gpurun_out | grep r2l
