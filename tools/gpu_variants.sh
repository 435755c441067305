#!/bin/bash
# usage: tools/gpu_variants.sh base <name>...: 640-step bench of build/libplantos_<name>.so variants (tools/build_variant.sh), one line each
mkdir -p gpurun_out/r2
for v in "$@"; do
  if [ "$v" = "base" ]; then unset PLANTOS_LIB; else export PLANTOS_LIB=build/libplantos_$v.so; fi
  timeout 300 python bench.py --steps 640 --warmup 64 --no-cpu-baseline --e2e-steps 3 --no-step-launch > gpurun_out/r2/bv_$v.json 2> gpurun_out/r2/bv_$v.err
  python - "$v" <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/r2/bv_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
    print("VARIANT", sys.argv[1], "us/step", round(d["ms_per_step"]*1e3,2), "frac", round(d["roofline"]["frac"],3))
except Exception as ex: print("VARIANT ERR", sys.argv[1], ex)
PY
done
