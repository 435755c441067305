#!/bin/bash
# whole GPU suite + bench (pipelined and plain)
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_gpu.log 2>&1
tail -n 6 gpurun_out/r2/pytest_gpu.log
for mode in "" "--no-pipeline"; do
  timeout 300 python bench.py --steps 640 --warmup 64 --no-cpu-baseline --e2e-steps 3 $mode > gpurun_out/r2/b_full$mode.json 2> gpurun_out/r2/b_full$mode.err
  python - "$mode" <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/r2/b_full%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
    print("BENCH", sys.argv[1], "us/step", round(d["ms_per_step"]*1e3,2), "frac", round(d["roofline"]["frac"],3), "iso", round(d["roofline"]["isolated_launch_us_median"],2))
except Exception as ex: print("BENCH ERR", sys.argv[1], ex)
PY
done
