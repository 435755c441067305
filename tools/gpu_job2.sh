#!/bin/bash
mkdir -p gpurun_out/r2
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2/pytest_parity.log 2>&1
tail -n 15 gpurun_out/r2/pytest_parity.log
python bench.py --steps 640 --warmup 64 --no-cpu-baseline --e2e-steps 3 > gpurun_out/r2/bench_tile.json 2> gpurun_out/r2/bench_tile.err
tail -c 1500 gpurun_out/r2/bench_tile.json; tail -n 5 gpurun_out/r2/bench_tile.err
