#!/bin/bash
mkdir -p gpurun_out
python scratch/mb_conv.py up4 3 > gpurun_out/plain_up4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tapgemm_halo -s 3 -c 1 -f -o gpurun_out/r2_up4_dual python scratch/mb_conv.py up4 3 > gpurun_out/ncu_up4_dual.log 2>&1
TBI_TC_NO_DUAL=1 python scratch/mb_conv.py up4 3 > gpurun_out/plain_up4_nd.log 2>&1 && \
TBI_TC_NO_DUAL=1 ncu --set full --clock-control none --import-source on -k regex:tapgemm_halo -s 3 -c 1 -f -o gpurun_out/r2_up4_single python scratch/mb_conv.py up4 3 > gpurun_out/ncu_up4_single.log 2>&1
python scratch/mb_conv.py wup4 3 > gpurun_out/plain_wup4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tapwgrad_tc -s 3 -c 1 -f -o gpurun_out/r2_wup4 python scratch/mb_conv.py wup4 3 > gpurun_out/ncu_wup4.log 2>&1
tail -2 gpurun_out/ncu_up4_dual.log gpurun_out/ncu_up4_single.log gpurun_out/ncu_wup4.log
ls -la gpurun_out/*.ncu-rep
