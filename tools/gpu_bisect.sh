#!/bin/bash
# usage: gpu_bisect.sh <variant names...>: bench step time of build/libplantos_<name>.so variants (timing-only experiments)
mkdir -p gpurun_out/r2
for v in "$@"; do
  if [ "$v" = "base" ]; then unset PLANTOS_LIB; else export PLANTOS_LIB=build/libplantos_$v.so; fi
  python bench.py --steps 320 --warmup 32 --no-cpu-baseline --e2e-steps 3 > gpurun_out/r2/bisect_$v.json 2> gpurun_out/r2/bisect_$v.err
  python - "$v" <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/r2/bisect_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
    print("VARIANT", sys.argv[1], "us/step", round(d["ms_per_step"]*1e3,2), "iso", round(d["roofline"]["isolated_launch_us_median"],2))
except Exception as ex:
    print("VARIANT", sys.argv[1], "ERR", ex)
PY
done
