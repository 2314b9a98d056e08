#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none -k regex:'k_(downmix|trim|eff)$' -c 9 --csv --log-file gpurun_out/r2p_bench_cond.csv python bench.py --steps 2 --warmup 3 --no-sink --no-parity-check > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
python - <<'P'
import csv
rows=list(csv.reader(open('gpurun_out/r2p_bench_cond.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=r; start=i; break
ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit'); idi=h.index('ID')
for r in rows[start+1:]:
    if len(r)>vi: print(r[idi], r[ki].split('(')[0], r[mi], r[vi], r[ui])
P
