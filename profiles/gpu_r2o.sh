#!/bin/bash
# Round 2, call O: per-launch time list of one compute at FULL bench scale (12.5k mixed files, hop 1024, all descriptors).
mkdir -p gpurun_out
export AFX_SINGLE_STREAM=1
PROF_MIXED=1 PROF_FILES=12500 timeout 1200 ncu --nvtx --nvtx-include "prof/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2o_launches_mixed12500.csv python profiles/prof_small.py > gpurun_out/r2o.log 2>&1
tail -3 gpurun_out/r2o.log
python - <<'P'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r2o_launches_mixed12500.csv')))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: h=r; start=i; break
ki=h.index('Kernel Name'); vi=h.index('Metric Value'); ui=h.index('Metric Unit')
acc=collections.OrderedDict(); cnt={}
for r in rows[start+2:]:
    if len(r)<=vi: continue
    k=r[ki].split('(')[0]; v=float(r[vi].replace(',','')); u=r[ui]
    v = v/1e6 if u=='ns' else v/1e3 if u=='us' else v*1e3 if u=='s' else v
    acc[k]=acc.get(k,0)+v; cnt[k]=cnt.get(k,0)+1
tot=sum(acc.values())
for k,v in acc.items(): print(f"{k:28s} {cnt[k]:4d} {v:9.3f} ms {100*v/tot:5.1f}%")
print(tot)
P
