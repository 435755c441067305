#!/bin/bash
# Round-2 evidence run (one GPU): test log, bench lines, ncu launch list, ncu full captures, steady-state DRAM.
# Everything lands in gpurun_out/ev2/; tools/summarize_profiles.py r2 turns it into profiles/r2_*.
OUT=gpurun_out/ev2
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; tail -n 3 $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/bench_1gpu.json 2> $OUT/bench_1gpu.err; tail -c 400 $OUT/bench_1gpu.json
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/bench_1gpu_20steps.json 2>> $OUT/bench_1gpu.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference_cpu.json 2> $OUT/bench_reference_cpu.err
for loop in graph eager; do
  timeout 300 python bench.py --loop $loop --steps 640 --warmup 64 --no-cpu-baseline --no-step-launch --e2e-steps 3 > $OUT/bench_loop_$loop.json 2>> $OUT/bench_1gpu.err
  timeout 300 python bench.py --loop $loop --no-pipeline --steps 640 --warmup 64 --no-cpu-baseline --no-step-launch --e2e-steps 3 > $OUT/bench_loop_${loop}_plain.json 2>> $OUT/bench_1gpu.err
done
timeout 300 python bench.py --no-stagger --steps 640 --warmup 64 --no-cpu-baseline --e2e-steps 3 > $OUT/bench_nostagger.json 2>> $OUT/bench_1gpu.err
timeout 600 python tools/bench_presets.py 320 > $OUT/presets.log 2>&1
# ncu: launch list of the default bench (short), no cache flush between launches
CMD="python bench.py --steps 64 --warmup 16 --no-cpu-baseline --e2e-steps 3"
$CMD > $OUT/plain_launches.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
# ncu: full captures of the rollout kernel and of the single-step kernel
CMD="python bench.py --steps 48 --warmup 16 --no-cpu-baseline --e2e-steps 3 --no-step-launch"
$CMD > $OUT/plain_full_rollout.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_tile -s 4 -c 1 -f -o $OUT/prof_rollout $CMD > $OUT/ncu_full_rollout.log 2>&1
CMD="python bench.py --loop eager --no-pipeline --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 3 --no-step-launch"
$CMD > $OUT/plain_full_step.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_tile -s 30 -c 1 -f -o $OUT/prof_step $CMD > $OUT/ncu_full_step.log 2>&1
# steady-state DRAM traffic (application replay, no cache flush)
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,gpu__time_duration.sum
CMD="python bench.py --steps 96 --warmup 16 --no-cpu-baseline --e2e-steps 3 --no-step-launch"
ncu --replay-mode application --cache-control none --clock-control none --metrics $M -k regex:k_tile -s 6 -c 3 --csv --log-file $OUT/steady_dram_rollout.csv $CMD > $OUT/ncu_dram_rollout.log 2>&1
CMD="python bench.py --loop eager --no-pipeline --steps 60 --warmup 5 --no-cpu-baseline --e2e-steps 3 --no-step-launch"
ncu --replay-mode application --cache-control none --clock-control none --metrics $M -k regex:k_tile -s 60 -c 4 --csv --log-file $OUT/steady_dram_step.csv $CMD > $OUT/ncu_dram_step.log 2>&1
ls -la $OUT | head -40
