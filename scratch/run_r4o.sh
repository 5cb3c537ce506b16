#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gpu.py -x -q > gpurun_out/t_r4o.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_r4o.log
tail -30 gpurun_out/t_r4o.log
for c in stem dstem s16; do timeout 120 python scratch/mb_conv.py $c 10 2>&1 | tail -1; done
