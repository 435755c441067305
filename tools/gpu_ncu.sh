#!/bin/bash
# usage: gpu_ncu.sh <tag> <kernel-regex> [extra bench args]  -- one ncu --set full capture of the bench's step kernel
TAG=$1; KRE=$2; shift 2
mkdir -p gpurun_out/r2
CMD="python bench.py --steps 30 --warmup 3 --no-cpu-baseline --e2e-steps 3 --no-graph --no-pipeline $@"
$CMD > gpurun_out/r2/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KRE -s 20 -c 3 -f -o gpurun_out/r2/prof_$TAG $CMD > gpurun_out/r2/ncu_$TAG.log 2>&1
tail -n 3 gpurun_out/r2/ncu_$TAG.log
