#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_model_gpu.py -x -q 2>&1 | tail -3
B="python bench.py --steps 20 --warmup 5 --no-extras --cpu-seconds 0.2"
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], d["ms_per_step"], d["value"])'
$B 2>/dev/null | python -c "$P" default
TBI_HALO_NO_SIDE_PF=1 $B 2>/dev/null | python -c "$P" no_pf
TBI_HALO_RAGGED_BN32=1 $B 2>/dev/null | python -c "$P" ragged32
TBI_HALO_SLAB_BLOCKED=1 $B 2>/dev/null | python -c "$P" blocked
N="ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tapgemm_halo -c 100 --csv"
A="python bench.py --steps 1 --warmup 0 --no-graph --cpu-seconds 0.2 --no-extras"
$N --log-file gpurun_out/halo_pf.csv $A > /dev/null 2>&1
TBI_HALO_NO_SIDE_PF=1 $N --log-file gpurun_out/halo_nopf.csv $A > /dev/null 2>&1
TBI_HALO_RAGGED_BN32=1 $N --log-file gpurun_out/halo_r32.csv $A > /dev/null 2>&1
TBI_HALO_SLAB_BLOCKED=1 $N --log-file gpurun_out/halo_blk.csv $A > /dev/null 2>&1
