#!/bin/bash
# bench (pipelined) for several persistent-grid sizes of k_step_tile
mkdir -p gpurun_out/r2
for g in "$@"; do
  export PLANTOS_FAST_GRID=$g
  timeout 300 python bench.py --steps 640 --warmup 64 --no-cpu-baseline --e2e-steps 3 > gpurun_out/r2/bg_$g.json 2> gpurun_out/r2/bg_$g.err
  python - "$g" <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/r2/bg_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
    print("GRID", sys.argv[1], "us/step", round(d["ms_per_step"]*1e3,2), "frac", round(d["roofline"]["frac"],3))
except Exception as ex: print("GRID ERR", sys.argv[1], ex)
PY
done
