#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum --clock-control none -s 0 -c 420 --csv --log-file gpurun_out/step_metrics.csv python bench.py --steps 1 --warmup 1 --no-graph --cpu-seconds 0.2 --no-extras > gpurun_out/ncu_sm.log 2>&1
wc -l gpurun_out/step_metrics.csv
