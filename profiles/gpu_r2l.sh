#!/bin/bash
# Round 2, call L: ncu --set full with source counters for selected kernels (mixed corpus).  $1 = kernel regex
mkdir -p gpurun_out
K=${1:-'k_(bands_lane|bands_select)'}
export AFX_SINGLE_STREAM=1
PROF_MIXED=1 PROF_FILES=400 timeout 900 ncu --nvtx --nvtx-include "prof/" --set full --clock-control none --import-source on -k regex:"$K" -f -o gpurun_out/r2l python profiles/prof_small.py > gpurun_out/r2l.log 2>&1
ncu -i gpurun_out/r2l.ncu-rep --page raw --csv > gpurun_out/r2l_raw.csv 2>/dev/null
for k in $(python - <<'P'
import csv
rows=list(csv.reader(open('gpurun_out/r2l_raw.csv')))
i=rows[0].index('Kernel Name')
print(' '.join(sorted({r[i].split('(')[0].split('<')[0].replace('void ','') for r in rows[2:]})))
P
); do
  ncu -i gpurun_out/r2l.ncu-rep --page source --csv -k regex:$k > gpurun_out/r2l_source_$k.csv 2>/dev/null
done
python profiles/ncu_summary.py gpurun_out/r2l_raw.csv > gpurun_out/r2l_summary.txt
rm -f gpurun_out/r2l.ncu-rep
ls gpurun_out | grep r2l
