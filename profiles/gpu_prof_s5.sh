#!/bin/bash
# Round-1 session-5 profile set (one gpurun call): per-launch time lists and `ncu --set full` of one compute of
# profiles/prof_small.py for (a) config 2's kernels (hop 512, spectral subset) and (b) the full set (hop 1024), plus a
# launch list at bench scale on the mixed-length corpus.  Raw CSV exports land in gpurun_out/ ($1 = tag).
TAG=${1:-s5}
mkdir -p gpurun_out
export AFX_SINGLE_STREAM=1
K='k_(spectrum|bands|pitch|autocorr|rhythm|peaks|whiten|stats|flux)'
PROF_FEATS=spectral PROF_HOP=512 timeout 600 ncu --nvtx --nvtx-include "prof/" --set full --clock-control none --import-source on -k regex:"$K" -f -o gpurun_out/${TAG}_config2 python profiles/prof_small.py > gpurun_out/${TAG}_config2.log 2>&1
timeout 900 ncu --nvtx --nvtx-include "prof/" --set full --clock-control none --import-source on -k regex:"$K" -f -o gpurun_out/${TAG}_all python profiles/prof_small.py > gpurun_out/${TAG}_all.log 2>&1
for n in config2 all; do ncu -i gpurun_out/${TAG}_$n.ncu-rep --page raw --csv > gpurun_out/${TAG}_${n}_raw.csv 2>/dev/null; done
PROF_FEATS=spectral PROF_HOP=512 timeout 600 ncu --nvtx --nvtx-include "prof/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_config2.csv python profiles/prof_small.py > /dev/null 2>&1
timeout 600 ncu --nvtx --nvtx-include "prof/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_all.csv python profiles/prof_small.py > /dev/null 2>&1
PROF_MIXED=1 PROF_FILES=4000 timeout 800 ncu --nvtx --nvtx-include "prof/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_mixed4000.csv python profiles/prof_small.py > /dev/null 2>&1
rm -f gpurun_out/${TAG}_all.ncu-rep gpurun_out/${TAG}_config2.ncu-rep
ls -la gpurun_out | tail -12
