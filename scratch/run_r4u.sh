#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:prepare_kernel -s 1 -c 1 -f -o gpurun_out/r4_prepare python bench.py --steps 1 --warmup 0 --no-graph --cpu-seconds 0.2 --no-extras > gpurun_out/ncu_prep.log 2>&1
ls -la gpurun_out/r4_prepare.ncu-rep
