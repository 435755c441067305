#!/bin/bash
# whole GPU suite + default bench (+ optional extra bench arg sets, one per argument)
mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest_gpu.log 2>&1
tail -n 6 gpurun_out/r2/pytest_gpu.log
i=0
for extra in "" "$@"; do
  timeout 600 python bench.py --steps 640 --warmup 64 --no-cpu-baseline --e2e-steps 3 $extra > gpurun_out/r2/b_full$i.json 2> gpurun_out/r2/b_full$i.err
  python - "$i" "$extra" <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/r2/b_full%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
    print("BENCH [%s]"%sys.argv[2], "us/step", round(d["ms_per_step"]*1e3,2), "frac", round(d["roofline"]["frac"],3), "episodes", d["episode_stats"]["episodes"], {k:(round(v,2) if isinstance(v,float) else v) for k,v in d["step_launch"].items()})
except Exception as ex: print("BENCH ERR", sys.argv[2], ex)
PY
  i=$((i+1))
done
