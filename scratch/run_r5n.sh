#!/bin/bash
mkdir -p gpurun_out
python scratch/mb_conv.py up4 3 > gpurun_out/plain_up4_r5.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/plain_up4_r5.log
ncu --set full --clock-control none --import-source on -k regex:tapgemm_halo -s 3 -c 1 -f -o gpurun_out/r5_up4_pair python scratch/mb_conv.py up4 3 > gpurun_out/ncu_up4_r5.log 2>&1; echo "ncu rc=$?"
python bench.py --steps 2 --warmup 1 --no-graph --no-extras --cpu-seconds 0.2 > gpurun_out/plain_step_r5.log 2>&1; echo "plain step rc=$?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -s 0 -c 600 --csv --log-file gpurun_out/step_metrics_r5.csv python bench.py --steps 2 --warmup 1 --no-graph --cpu-seconds 0.2 --no-extras > gpurun_out/ncu_sm_r5.log 2>&1; echo "ncu step rc=$?"
wc -l gpurun_out/step_metrics_r5.csv
