#!/bin/bash
# Round 2, call AA: end-to-end leg of the default bench with fewer / larger chunks and more slots (the 6 % between `e2e` and
# `value` at N = 1 is chunk granularity: every kernel's tail once per chunk).
mkdir -p gpurun_out
for v in ${VARIANTS:-6_3 8_3 8_4 4_3}; do set -- ${v/_/ }
  timeout 400 python bench.py --no-cpu-baseline --no-sink --no-parity-check --steps 6 --warmup 3 --e2e-chunks $1 --e2e-slots $2 > gpurun_out/r2aa_chunks$1_slots$2.json 2> gpurun_out/r2aa_chunks$1_slots$2.err
  python - <<P
import json
d=json.loads(open('gpurun_out/r2aa_chunks$1_slots$2.json').read().strip().splitlines()[-1])
print('chunks $1 slots $2: value %.1f e2e %.1f (%.1f ms vs %.1f ms)' % (d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['ms_per_step']))
P
done
