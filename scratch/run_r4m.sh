#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:smallcin_wgrad_cols -c 1 -f -o gpurun_out/r4_wcols python bench.py --steps 1 --warmup 0 --no-graph --cpu-seconds 0.2 --no-extras > gpurun_out/ncu_wcols.log 2>&1
ls -la gpurun_out/r4_wcols.ncu-rep
