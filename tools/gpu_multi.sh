#!/bin/bash
# usage: tools/gpu_multi.sh <ngpus> [bench args]  -- the driver's multi-GPU launch of bench.py (torchrun, one rank per GPU)
N=$1; shift
OUT=gpurun_out/ev2
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo_${N}gpu.txt 2>&1
lscpu | grep -i -E "numa|socket|model name|^cpu\(s\)" >> $OUT/topo_${N}gpu.txt 2>&1
NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 20 --warmup 5 "$@" > $OUT/bench_${N}gpu.json 2> $OUT/bench_${N}gpu.err
echo "rc=$?"; tail -c 1500 $OUT/bench_${N}gpu.json
grep -c "NCCL INFO" $OUT/bench_${N}gpu.err; grep -m3 -E "comm .* nranks|Init COMPLETE" $OUT/bench_${N}gpu.err | cut -c1-200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps 640 --warmup 64 --no-cpu-baseline "$@" > $OUT/bench_${N}gpu_640steps.json 2>> $OUT/bench_${N}gpu.err
tail -c 600 $OUT/bench_${N}gpu_640steps.json
