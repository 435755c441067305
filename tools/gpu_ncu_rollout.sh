#!/bin/bash
# usage: gpu_ncu_rollout.sh <tag>  -- one ncu --set full capture of the state-resident rollout kernel inside the default bench loop
TAG=$1; shift
mkdir -p gpurun_out/r2
CMD="python bench.py --steps 48 --warmup 16 --no-cpu-baseline --e2e-steps 3 --no-step-launch $@"
$CMD > gpurun_out/r2/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_rollout_tile -s 6 -c 1 -f -o gpurun_out/r2/prof_$TAG $CMD > gpurun_out/r2/ncu_$TAG.log 2>&1
tail -n 3 gpurun_out/r2/ncu_$TAG.log
