#!/bin/bash
mkdir -p gpurun_out
echo "== up4 pair"; timeout 120 python scratch/trace_stream.py up4 2>&1 | grep -E "^M|steady|^E" | head -30
python scratch/mb_conv.py up4 3 > gpurun_out/plain_up4p.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tapgemm_halo -s 3 -c 1 -f -o gpurun_out/r2_up4_pair python scratch/mb_conv.py up4 3 > gpurun_out/ncu_up4_pair.log 2>&1
ls -la gpurun_out/r2_up4_pair.ncu-rep
