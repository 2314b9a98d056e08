#!/bin/bash
# Round 2, call H (profile set): `ncu --set full` of one compute of profiles/prof_small.py for the full set (hop 1024) and
# config 2's kernels (hop 512, spectral subset), per-launch time lists incl. the mixed corpus at bench scale, and the three
# bench lines (default = full workload with the sink leg, config 2, the reference arm).  Raw CSV exports -> gpurun_out/.
TAG=${1:-r02h}
mkdir -p gpurun_out
export AFX_SINGLE_STREAM=1
K='k_(spectrum|bands|pitch|autocorr|rhythm|peaks|whiten|stats|flux|downmix|trim|eff)'
timeout 900 ncu --nvtx --nvtx-include "prof/" --set full --clock-control none --import-source on -k regex:"$K" -f -o gpurun_out/${TAG}_all python profiles/prof_small.py > gpurun_out/${TAG}_all.log 2>&1
PROF_FEATS=spectral PROF_HOP=512 timeout 600 ncu --nvtx --nvtx-include "prof/" --set full --clock-control none --import-source on -k regex:"$K" -f -o gpurun_out/${TAG}_config2 python profiles/prof_small.py > gpurun_out/${TAG}_config2.log 2>&1
for n in config2 all; do ncu -i gpurun_out/${TAG}_$n.ncu-rep --page raw --csv > gpurun_out/${TAG}_${n}_raw.csv 2>/dev/null; done
timeout 600 ncu --nvtx --nvtx-include "prof/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_all.csv python profiles/prof_small.py > /dev/null 2>&1
PROF_FEATS=spectral PROF_HOP=512 timeout 600 ncu --nvtx --nvtx-include "prof/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_config2.csv python profiles/prof_small.py > /dev/null 2>&1
PROF_MIXED=1 PROF_FILES=4000 timeout 800 ncu --nvtx --nvtx-include "prof/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_mixed4000.csv python profiles/prof_small.py > /dev/null 2>&1
rm -f gpurun_out/${TAG}_all.ncu-rep gpurun_out/${TAG}_config2.ncu-rep
unset AFX_SINGLE_STREAM
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_full_n1.json 2> gpurun_out/${TAG}_bench_full.err; tail -c 300 gpurun_out/${TAG}_bench_full_n1.json; tail -3 gpurun_out/${TAG}_bench_full.err
timeout 600 python bench.py --workload config2 --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_config2_n1.json 2> gpurun_out/${TAG}_bench_config2.err; tail -c 300 gpurun_out/${TAG}_bench_config2_n1.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_reference.err; tail -c 600 gpurun_out/${TAG}_bench_reference_arm.json
ls -la gpurun_out | tail -14
