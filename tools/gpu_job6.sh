#!/bin/bash
mkdir -p gpurun_out/r2
for extra in "" "--no-stagger"; do
timeout 600 python bench.py --steps 640 --warmup 64 --no-cpu-baseline --e2e-steps 3 $extra > gpurun_out/r2/bench_new$extra.json 2> gpurun_out/r2/bench_new$extra.err
python - "$extra" <<'PY'
import json,sys
f="bench_new"+sys.argv[1]
try:
    d=json.loads(open('gpurun_out/r2/%s.json'%f).read().strip().splitlines()[-1])
    print(f, "us/step", round(d["ms_per_step"]*1e3,2), "frac", round(d["roofline"]["frac"],3), "launches", d["gpu_launches"], "episodes", d["episode_stats"]["episodes"], {k:(round(v,2) if isinstance(v,float) else v) for k,v in d["step_launch"].items()})
except Exception as ex: print(f, "ERR", ex)
PY
done
