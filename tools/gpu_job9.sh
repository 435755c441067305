#!/bin/bash
mkdir -p gpurun_out/r2
for rep in 1 2 3 4; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 3 --no-step-launch > gpurun_out/r2/bc_$rep.json 2> gpurun_out/r2/bc_$rep.err
  python - "$rep" <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/r2/bc_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
    print("SHORT", sys.argv[1], "us/step", round(d["ms_per_step"]*1e3,2), "frac", round(d["roofline"]["frac"],3), d["clocks"])
except Exception as ex: print("SHORT ERR", sys.argv[1], ex)
PY
done
