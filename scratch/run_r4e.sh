#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-graph --cpu-seconds 0.2 --no-extras > gpurun_out/plain_step.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 700 --csv --log-file gpurun_out/launches_r4e.csv python bench.py --steps 2 --warmup 1 --no-graph --cpu-seconds 0.2 --no-extras > gpurun_out/ncu_step.log 2>&1
wc -l gpurun_out/launches_r4e.csv
