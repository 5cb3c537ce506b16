#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_model_gpu.py -x -q 2>&1 | tail -3
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], d["ms_per_step"], d["value"], d["e2e"]["value"])'
B="python bench.py --steps 30 --warmup 5 --no-extras --cpu-seconds 0.2"
$B 2>/dev/null | python -c "$P" lag
TBI_HALO_PF_LAG=0 $B 2>/dev/null | python -c "$P" lag0
$B 2>/dev/null | python -c "$P" lag
N="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:tapgemm_halo|tapgemm_xpack -c 100 --csv"
A="python bench.py --steps 1 --warmup 0 --no-graph --cpu-seconds 0.2 --no-extras"
$N --log-file gpurun_out/halo_lag.csv $A > /dev/null 2>&1
