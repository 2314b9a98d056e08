#!/bin/bash
# Round 2, call X: the closing set on the final code -- the whole `-m gpu` suite, smoke(), `ncu --set full` of one compute of
# profiles/prof_small.py for the full set (hop 1024) and config 2's kernels (hop 512, spectral subset), per-launch time lists
# (400 x 3-s files, 4000 and 12 500 mixed-length files = bench scale), the DRAM-traffic table rebuilt from those captures, and
# then the three bench lines (default = full workload with the sink leg, config 2, the reference arm) reading that table.
TAG=${1:-r02z}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/${TAG}_pytest_gpu.log; cat gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 > gpurun_out/${TAG}_smoke.log; cat gpurun_out/${TAG}_smoke.log
export AFX_SINGLE_STREAM=1
K='k_(spectrum|bands|pitch|autocorr|rhythm|peaks|whiten|stats|flux|downmix|trim|eff)'
timeout 900 ncu --nvtx --nvtx-include "prof/" --set full --clock-control none --import-source on -k regex:"$K" -f -o gpurun_out/${TAG}_all python profiles/prof_small.py > gpurun_out/${TAG}_all.log 2>&1
PROF_FEATS=spectral PROF_HOP=512 timeout 600 ncu --nvtx --nvtx-include "prof/" --set full --clock-control none --import-source on -k regex:"$K" -f -o gpurun_out/${TAG}_config2 python profiles/prof_small.py > gpurun_out/${TAG}_config2.log 2>&1
for n in config2 all; do ncu -i gpurun_out/${TAG}_$n.ncu-rep --page raw --csv > gpurun_out/${TAG}_${n}_raw.csv 2>/dev/null; python profiles/ncu_summary.py gpurun_out/${TAG}_${n}_raw.csv > gpurun_out/${TAG}_ncu_full_summary_${n}.txt; done
for k in k_pitch_hop k_peaks_pipe k_spectrum; do ncu -i gpurun_out/${TAG}_all.ncu-rep --page source --csv -k regex:$k > gpurun_out/${TAG}_source_$k.csv 2>/dev/null; python profiles/ncu_phases.py gpurun_out/${TAG}_source_$k.csv 1.0 > gpurun_out/${TAG}_phases_$k.txt 2>&1; rm -f gpurun_out/${TAG}_source_$k.csv; done
timeout 600 ncu --nvtx --nvtx-include "prof/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_all.csv python profiles/prof_small.py > /dev/null 2>&1
PROF_FEATS=spectral PROF_HOP=512 timeout 600 ncu --nvtx --nvtx-include "prof/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_config2.csv python profiles/prof_small.py > /dev/null 2>&1
PROF_MIXED=1 PROF_FILES=4000 timeout 800 ncu --nvtx --nvtx-include "prof/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_mixed4000.csv python profiles/prof_small.py > /dev/null 2>&1
PROF_MIXED=1 PROF_FILES=12500 timeout 800 ncu --nvtx --nvtx-include "prof/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_mixed12500.csv python profiles/prof_small.py > /dev/null 2>&1
rm -f gpurun_out/${TAG}_all.ncu-rep gpurun_out/${TAG}_config2.ncu-rep
unset AFX_SINGLE_STREAM
cp gpurun_out/${TAG}_all_raw.csv gpurun_out/${TAG}_config2_raw.csv profiles/
python profiles/make_traffic.py spectral:512:102000:profiles/${TAG}_config2_raw.csv all:1024:51200:profiles/${TAG}_all_raw.csv | tail -1
cp profiles/ncu_traffic.json gpurun_out/${TAG}_ncu_traffic.json
timeout 900 python bench.py > gpurun_out/${TAG}_bench_full_n1.json 2> gpurun_out/${TAG}_bench_full.err; tail -c 300 gpurun_out/${TAG}_bench_full_n1.json; tail -3 gpurun_out/${TAG}_bench_full.err
timeout 600 python bench.py --workload config2 --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_config2_n1.json 2> gpurun_out/${TAG}_bench_config2.err; tail -c 300 gpurun_out/${TAG}_bench_config2_n1.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_reference.err; tail -c 600 gpurun_out/${TAG}_bench_reference_arm.json
ls -la gpurun_out | grep ${TAG}
