#!/bin/bash
mkdir -p gpurun_out/r2
for args in "--steps 20 --warmup 5" "--steps 20 --warmup 5" "--steps 100 --warmup 10" "--steps 640 --warmup 64"; do
timeout 600 python bench.py $args --no-cpu-baseline --e2e-steps 3 --no-step-launch > gpurun_out/r2/bench_s.json 2> gpurun_out/r2/bench_s.err
python - "$args" <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/r2/bench_s.json').read().strip().splitlines()[-1])
    print("BENCH", sys.argv[1], "us/step", round(d["ms_per_step"]*1e3,2), "frac", round(d["roofline"]["frac"],3), "launches", d["gpu_launches"], "episodes", d["episode_stats"]["episodes"])
except Exception as ex: print("ERR", sys.argv[1], ex); print(open('gpurun_out/r2/bench_s.err').read()[-1500:])
PY
done
