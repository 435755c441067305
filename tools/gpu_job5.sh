#!/bin/bash
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_vecenv.py -m gpu -x -q -k "step_many or rollout or pipelined" > gpurun_out/r2/pytest_many.log 2>&1
tail -n 25 gpurun_out/r2/pytest_many.log
timeout 300 python tools/exp_many.py 2>&1 | tail -12
