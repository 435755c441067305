#!/bin/bash
# round-2 GPU job 1: output-stage micro-benchmark + baseline + persistent-grid experiments
mkdir -p gpurun_out/r2
nvidia-smi > gpurun_out/r2/smi.txt 2>&1
for cfg in "16 1" "8 2" "8 4" "4 8" "16 2" "32 1"; do tools/ubench_decode 131072 $cfg 300; done > gpurun_out/r2/ubench_decode.log 2>&1
python bench.py --steps 640 --warmup 64 --no-cpu-baseline --e2e-steps 3 > gpurun_out/r2/bench_base.json 2> gpurun_out/r2/bench_base.err
for g in 296 444 148; do
PLANTOS_FAST_GRID=$g python bench.py --steps 640 --warmup 64 --no-cpu-baseline --e2e-steps 3 > gpurun_out/r2/bench_grid$g.json 2> gpurun_out/r2/bench_grid$g.err
done
tail -n 3 gpurun_out/r2/*.json gpurun_out/r2/ubench_decode.log
