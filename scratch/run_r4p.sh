#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:tapgemm_xpack -s 3 -c 1 -f -o gpurun_out/r4_xpack_stem python scratch/mb_conv.py stem 3 > gpurun_out/ncu_xp.log 2>&1
ls -la gpurun_out/r4_xpack_stem.ncu-rep
