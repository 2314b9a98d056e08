#!/bin/bash
# Round 2, call M: quick parity subset + per-group timing (A/B runs of kernel variants).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_longfile.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2m_tests.log; cat gpurun_out/r2m_tests.log
timeout 900 python bench.py --steps 6 --warmup 3 --no-sink > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/r2m_bench.json").read().strip().splitlines()[-1])
print("value %.2f  ms/step %.1f  e2e %.2f  e2e ms %.1f  parity %s/%s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d.get("parity_checked"), d.get("parity_mismatches")))
print({k: round(v["ms"], 1) for k, v in d["roofline"]["groups"].items()})
P
