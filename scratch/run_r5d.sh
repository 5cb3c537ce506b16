#!/bin/bash
mkdir -p gpurun_out
A="python bench.py --steps 1 --warmup 0 --no-graph --cpu-seconds 0.2 --no-extras"
ncu --set full --import-source on --clock-control none -k regex:tapgemm_halo -s 21 -c 1 -f -o gpurun_out/r5_head_dgrad $A > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:tapgemm_halo -s 41 -c 1 -f -o gpurun_out/r5_stem_dgrad $A > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
