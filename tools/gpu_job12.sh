#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q -k "pipelined or step_many or rollout" 2>&1 | tail -3
for args in "--steps 640 --warmup 64" "--steps 640 --warmup 64 --no-pipeline" "--steps 20 --warmup 5" "--steps 20 --warmup 5 --chunk 10" "--steps 20 --warmup 5 --chunk 5" "--steps 20 --warmup 5 --no-pipeline" "--steps 640 --warmup 64 --chunk 8"; do
  timeout 300 python bench.py $args --no-cpu-baseline --e2e-steps 3 --no-step-launch 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('BENCH [$args] us/step', round(d['ms_per_step']*1e3,2), 'frac', round(d['roofline']['frac'],3), 'launches', d['gpu_launches'])"
done
