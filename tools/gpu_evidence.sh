#!/bin/bash
# Round-2 evidence runs (one GPU).  usage: tools/gpu_evidence.sh <stage>; one profiler invocation per stage, and
# only after the same command has exited 0 without the profiler.  Everything lands in gpurun_out/ev2/;
# tools/summarize_profiles.py r2 turns it into profiles/r2_*.
#   tests     pytest -m gpu
#   bench     bench.py lines: default, the driver's --steps 20 --warmup 5, reference arm, graph / eager loops, no-stagger
#   presets   tools/bench_presets.py (the other BASELINE.json configurations)
#   launches  ncu launch list of the default bench loop (no cache flush between launches)
#   rollout   ncu --set full of one k_rollout_tile launch        step     ncu --set full of one k_step_tile launch
#   others    ncu --set full of k_step_generic / k_reset_all / k_wrc_build / k_reset_done (one launch each)
#   dram_rollout / dram_step   steady-state DRAM traffic (application replay, no cache flush)
OUT=gpurun_out/ev2
mkdir -p $OUT
# hash of the kernel sources the captures below are taken from (bench.py reports the DRAM figure only for these sources)
python -c "from rl_env_b200.build import kernel_source_hash; print(kernel_source_hash())" > $OUT/kernel_source_sha_$1.txt
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,gpu__time_duration.sum
case "$1" in
tests)
  nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu.txt 2>&1
  timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; tail -n 3 $OUT/pytest_gpu.log ;;
bench)
  timeout 600 python bench.py > $OUT/bench_1gpu.json 2> $OUT/bench_1gpu.err; tail -c 300 $OUT/bench_1gpu.json
  timeout 300 python bench.py --steps 20 --warmup 5 > $OUT/bench_1gpu_20steps.json 2>> $OUT/bench_1gpu.err
  timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference_cpu.json 2> $OUT/bench_reference_cpu.err
  for loop in graph eager; do
    timeout 300 python bench.py --loop $loop --steps 640 --warmup 64 --no-cpu-baseline --no-step-launch --e2e-steps 3 > $OUT/bench_loop_$loop.json 2>> $OUT/bench_1gpu.err
    timeout 300 python bench.py --loop $loop --no-pipeline --steps 640 --warmup 64 --no-cpu-baseline --no-step-launch --e2e-steps 3 > $OUT/bench_loop_${loop}_plain.json 2>> $OUT/bench_1gpu.err
  done
  timeout 300 python bench.py --no-stagger --steps 640 --warmup 64 --no-cpu-baseline --e2e-steps 3 > $OUT/bench_nostagger.json 2>> $OUT/bench_1gpu.err
  timeout 300 python bench.py --loop graph --no-stagger --steps 640 --warmup 64 --no-cpu-baseline --no-step-launch --e2e-steps 3 > $OUT/bench_loop_graph_nostagger.json 2>> $OUT/bench_1gpu.err
  timeout 300 python bench.py --preset default --steps 640 --warmup 64 --no-cpu-baseline --e2e-steps 3 > $OUT/bench_preset_default.json 2>> $OUT/bench_1gpu.err
  timeout 300 python bench.py --preset xl --steps 96 --warmup 16 --no-cpu-baseline --e2e-steps 3 > $OUT/bench_preset_xl.json 2>> $OUT/bench_1gpu.err
  timeout 300 python bench.py --envs-per-gpu 4096 --steps 640 --warmup 64 --no-cpu-baseline --e2e-steps 20 > $OUT/bench_4096.json 2>> $OUT/bench_1gpu.err
  ls -la $OUT ;;
presets)
  timeout 900 python tools/bench_presets.py 320 > $OUT/presets.log 2>&1; cat $OUT/presets.log ;;
launches)
  CMD="python bench.py --chunk 16 --steps 64 --warmup 16 --no-cpu-baseline --e2e-steps 3 --no-step-launch"
  $CMD > $OUT/plain_launches.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file $OUT/launches.csv $CMD > $OUT/ncu_launches.log 2>&1
  tail -n 2 $OUT/ncu_launches.log ;;
rollout)
  CMD="python bench.py --chunk 16 --steps 48 --warmup 16 --no-cpu-baseline --e2e-steps 3 --no-step-launch"
  $CMD > $OUT/plain_full_rollout.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k_rollout_tile -s 6 -c 1 -f -o $OUT/prof_rollout $CMD > $OUT/ncu_full_rollout.log 2>&1
  tail -n 2 $OUT/ncu_full_rollout.log ;;
step)
  CMD="python bench.py --loop eager --no-pipeline --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 3 --no-step-launch"
  $CMD > $OUT/plain_full_step.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:k_step_tile -s 700 -c 1 -f -o $OUT/prof_step $CMD > $OUT/ncu_full_step.log 2>&1
  tail -n 2 $OUT/ncu_full_step.log ;;
others)
  CMD="python tools/run_other_kernels.py"
  $CMD > $OUT/plain_others.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:'k_step_generic|k_reset_all|k_wrc_build|k_reset_done|k_step_fast' -s 10 -c 12 -f -o $OUT/prof_others $CMD > $OUT/ncu_full_others.log 2>&1
  tail -n 2 $OUT/ncu_full_others.log ;;
dram_rollout)
  CMD="python bench.py --chunk 16 --steps 96 --warmup 16 --no-cpu-baseline --e2e-steps 3 --no-step-launch"
  $CMD > $OUT/plain_dram_rollout.log 2>&1 &&
  ncu --replay-mode application --cache-control none --clock-control none --metrics $M -k regex:k_rollout_tile -s 44 -c 3 --csv --log-file $OUT/steady_dram_rollout.csv $CMD > $OUT/ncu_dram_rollout.log 2>&1
  tail -n 4 $OUT/steady_dram_rollout.csv | cut -c1-300 ;;
dram_step)
  CMD="python bench.py --loop eager --no-pipeline --steps 60 --warmup 5 --no-cpu-baseline --e2e-steps 3 --no-step-launch"
  $CMD > $OUT/plain_dram_step.log 2>&1 &&
  ncu --replay-mode application --cache-control none --clock-control none --metrics $M -k regex:k_step_tile -s 700 -c 4 --csv --log-file $OUT/steady_dram_step.csv $CMD > $OUT/ncu_dram_step.log 2>&1
  tail -n 4 $OUT/steady_dram_step.csv | cut -c1-300 ;;
*) echo "unknown stage $1"; exit 1 ;;
esac
