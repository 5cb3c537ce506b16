#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py tests/test_parity_fullres_gpu.py tests/test_reference_fixture_gpu.py -x -q 2>&1 | tail -3
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], d["ms_per_step"], d["value"], d["e2e"]["value"])'
B="python bench.py --steps 50 --warmup 5 --no-extras --cpu-seconds 0.2"
$B 2>/dev/null | python -c "$P" fused
TBI_HEAD_FUSED_LOSS=0 $B 2>/dev/null | python -c "$P" unfused
$B 2>/dev/null | python -c "$P" fused
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"softmax_loss|convt_scatter4" -c 4 --csv --log-file gpurun_out/loss_times.csv python bench.py --steps 1 --warmup 1 --no-graph --cpu-seconds 0.2 --no-extras > /dev/null 2>&1
grep -v "^==" gpurun_out/loss_times.csv | awk -F'","' '{print $5, $NF}' | tail -4
