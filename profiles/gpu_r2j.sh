#!/bin/bash
# Round 2, call J: ncu --set full of the block-sharing pitch kernel with source counters.
mkdir -p gpurun_out
export AFX_SINGLE_STREAM=1
for v in 4; do
  AFX_PITCH_NG=$v PROF_MIXED=1 PROF_FILES=400 timeout 600 ncu --nvtx --nvtx-include "prof/" --set full --clock-control none --import-source on -k regex:'k_pitch' -f -o gpurun_out/r2j_pitch$v python profiles/prof_small.py > gpurun_out/r2j_pitch$v.log 2>&1
  ncu -i gpurun_out/r2j_pitch$v.ncu-rep --page raw --csv > gpurun_out/r2j_pitch${v}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r2j_pitch$v.ncu-rep --page source --csv > gpurun_out/r2j_pitch${v}_source.csv 2>/dev/null
  python profiles/ncu_summary.py gpurun_out/r2j_pitch${v}_raw.csv > gpurun_out/r2j_pitch${v}_summary.txt
done
rm -f gpurun_out/r2j_pitch*.ncu-rep
cat gpurun_out/r2j_pitch4_summary.txt
