#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_tc_gpu.py tests/test_model_gpu.py tests/test_parity_fullres_gpu.py -x -q 2>&1 | tail -3
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], d["ms_per_step"], d["value"], d["e2e"]["value"])'
B="python bench.py --steps 50 --warmup 5 --no-extras --cpu-seconds 0.2"
$B 2>/dev/null | python -c "$P" deep
TBI_TC_NO_DEEP_RING=1 $B 2>/dev/null | python -c "$P" base
$B 2>/dev/null | python -c "$P" deep
TBI_TC_NO_DEEP_RING=1 $B 2>/dev/null | python -c "$P" base
N="ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tapgemm_tc_kernel -c 40 --csv"
A="python bench.py --steps 1 --warmup 0 --no-graph --cpu-seconds 0.2 --no-extras"
$N --log-file gpurun_out/tc_deep.csv $A > /dev/null 2>&1
TBI_TC_NO_DEEP_RING=1 $N --log-file gpurun_out/tc_base.csv $A > /dev/null 2>&1
