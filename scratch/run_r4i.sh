#!/bin/bash
mkdir -p gpurun_out
python scratch/trace_halo.py s32 > gpurun_out/trace_s32.log 2>&1
cat gpurun_out/trace_s32.log | tail -45
python scratch/mb_conv.py stem 3 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tapgemm_halo -s 3 -c 1 -f -o gpurun_out/r4_stem_fwd python scratch/mb_conv.py stem 3 > gpurun_out/ncu_stem.log 2>&1
python scratch/mb_conv.py dstem 3 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:tapgemm_halo -s 3 -c 1 -f -o gpurun_out/r4_stem_dgrad python scratch/mb_conv.py dstem 3 > gpurun_out/ncu_dstem.log 2>&1
for c in stem dstem s16 p32 copy; do timeout 120 python scratch/mb_conv.py $c 10 2>&1 | tail -1; done
