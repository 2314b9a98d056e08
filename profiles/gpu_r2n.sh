#!/bin/bash
# Round 2, call N: the default bench (one compute stream per context; computes of a device chained) vs unchained contexts.
mkdir -p gpurun_out
timeout 900 python bench.py --steps 8 --warmup 3 --no-sink > gpurun_out/r2n_bench_chain.json 2> gpurun_out/r2n_bench_chain.err

python - <<'P'
import json
for t in ("chain",):
    d = json.loads(open("gpurun_out/r2n_bench_%s.json" % t).read().strip().splitlines()[-1])
    print(t, "value %.2f  ms/step %.1f  e2e %.2f  e2e ms %.1f  parity %s/%s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d.get("parity_checked"), d.get("parity_mismatches")))
P
