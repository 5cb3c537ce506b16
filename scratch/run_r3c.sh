#!/bin/bash
mkdir -p gpurun_out
python scratch/mb_conv.py wup3 3 > gpurun_out/plain_wup3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tapwgrad_pair -s 3 -c 1 -f -o gpurun_out/r2_wup3_pair python scratch/mb_conv.py wup3 3 > gpurun_out/ncu_wup3.log 2>&1
python scratch/mb_conv.py wup2 3 > gpurun_out/plain_wup2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tapwgrad_pair -s 3 -c 1 -f -o gpurun_out/r2_wup2_pair2 python scratch/mb_conv.py wup2 3 > gpurun_out/ncu_wup2.log 2>&1
for c in up4 up3 up2 wup4 wup3 wup2 wcc3 stem dstem wstem cc2; do timeout 120 python scratch/mb_conv.py $c 10 2>&1 | tail -1; done
