#!/bin/bash
mkdir -p gpurun_out/r2
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2/pytest_parity.log 2>&1
tail -n 15 gpurun_out/r2/pytest_parity.log
python bench.py --steps 640 --warmup 64 --no-cpu-baseline --e2e-steps 3 > gpurun_out/r2/bench_tile.json 2> gpurun_out/r2/bench_tile.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/bench_tile.json').read().strip().splitlines()[-1])
print("BENCH us/step", round(d["ms_per_step"]*1e3,2), "frac", round(d["roofline"]["frac"],3), "iso", d["roofline"]["isolated_launch_us_median"])
PY
tail -n 5 gpurun_out/r2/bench_tile.err
if [ -f build/libplantos_timing.so ]; then PLANTOS_LIB=build/libplantos_timing.so python tools/exp_timing.py > gpurun_out/r2/timing.log 2>&1; cat gpurun_out/r2/timing.log | head -24; fi
