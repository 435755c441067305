#!/bin/bash
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vecenv.py -m gpu -x -q > gpurun_out/r2/pytest_parity.log 2>&1
tail -n 15 gpurun_out/r2/pytest_parity.log
for mode in "" "--no-pipeline"; do
timeout 300 python bench.py --steps 640 --warmup 64 --no-cpu-baseline --e2e-steps 3 $mode > gpurun_out/r2/bench_tile$mode.json 2> gpurun_out/r2/bench_tile$mode.err
python - "$mode" <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/r2/bench_tile%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
    print("BENCH", sys.argv[1], "us/step", round(d["ms_per_step"]*1e3,2), "frac", round(d["roofline"]["frac"],3), "iso", d["roofline"]["isolated_launch_us_median"])
except Exception as ex: print("BENCH ERR", ex)
PY
tail -n 3 gpurun_out/r2/bench_tile$mode.err
done
