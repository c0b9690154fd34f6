#!/bin/bash
# launch list of one full step (both forwards + pose), cold-cache serialised times
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 120 --csv --log-file gpurun_out/r2w_launches.csv python bench.py --no-cpu-baseline --steps 1 --warmup 1 --chunks 1 > /dev/null 2>&1; wc -l gpurun_out/r2w_launches.csv
