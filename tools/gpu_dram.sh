#!/bin/bash
# steady-state DRAM / L2 traffic of the step kernel (no cache flush between launches, application replay)
TAG=$1; KRE=$2; shift 2
mkdir -p gpurun_out/r2
CMD="python bench.py --steps 60 --warmup 5 --no-cpu-baseline --e2e-steps 3 --no-graph $@"
$CMD > gpurun_out/r2/plain_dram_$TAG.log 2>&1 &&
ncu --replay-mode application --cache-control none --clock-control none \
  --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_read_lookup_hit.sum,lts__t_sectors_op_read_lookup_miss.sum,gpu__time_duration.sum \
  -k regex:$KRE -s 40 -c 4 --csv --log-file gpurun_out/r2/steady_dram_$TAG.csv $CMD > gpurun_out/r2/ncu_dram_$TAG.log 2>&1
tail -n 40 gpurun_out/r2/steady_dram_$TAG.csv | cut -d, -f5,13-
