#!/bin/bash
mkdir -p gpurun_out
python scratch/mb_conv.py wup2 3 > gpurun_out/plain_wup2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tapwgrad_pair -s 3 -c 1 -f -o gpurun_out/r2_wup2_pair python scratch/mb_conv.py wup2 3 > gpurun_out/ncu_wup2.log 2>&1
TBI_WGRAD_NO_PAIR=1 python scratch/mb_conv.py wup2 3 > gpurun_out/plain_wup2b.log 2>&1 && \
TBI_WGRAD_NO_PAIR=1 ncu --set full --clock-control none --import-source on -k regex:tapwgrad_tc -s 3 -c 1 -f -o gpurun_out/r2_wup2_single python scratch/mb_conv.py wup2 3 > gpurun_out/ncu_wup2b.log 2>&1
ls -la gpurun_out/r2_wup2*.ncu-rep
