#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_model_gpu.py -x -q 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tapgemm_halo" -c 100 --csv --log-file gpurun_out/halo_times.csv python bench.py --steps 1 --warmup 0 --no-graph --cpu-seconds 0.2 --no-extras > /dev/null 2>&1
TBI_TC_RESIDENT_RAGGED=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tapgemm_halo" -c 100 --csv --log-file gpurun_out/halo_times_old.csv python bench.py --steps 1 --warmup 0 --no-graph --cpu-seconds 0.2 --no-extras > /dev/null 2>&1
