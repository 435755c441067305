#!/bin/bash
# parity subset + bench of the shipped library and of named variants (pipelined and plain)
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vecenv.py -m gpu -x -q > gpurun_out/r2/pytest_parity.log 2>&1
tail -n 4 gpurun_out/r2/pytest_parity.log
for v in base "$@"; do
  if [ "$v" = "base" ]; then unset PLANTOS_LIB; else export PLANTOS_LIB=build/libplantos_$v.so; fi
  for mode in "" "--no-pipeline"; do
  timeout 300 python bench.py --steps 640 --warmup 64 --no-cpu-baseline --e2e-steps 3 $mode > gpurun_out/r2/b_$v$mode.json 2> gpurun_out/r2/b_$v$mode.err
  python - "$v" "$mode" <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/r2/b_%s%s.json'%(sys.argv[1],sys.argv[2])).read().strip().splitlines()[-1])
    print("BENCH", sys.argv[1], sys.argv[2], "us/step", round(d["ms_per_step"]*1e3,2), "frac", round(d["roofline"]["frac"],3), "iso", round(d["roofline"]["isolated_launch_us_median"],2))
except Exception as ex: print("BENCH ERR", sys.argv[1], ex)
PY
  done
done
