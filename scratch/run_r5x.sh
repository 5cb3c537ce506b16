#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_model_gpu.py tests/test_parity_fullres_gpu.py tests/test_reference_fixture_gpu.py tests/test_ops_gpu.py -x -q 2>&1 | tail -3
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], d["ms_per_step"], d["value"])'
B="python bench.py --steps 40 --warmup 5 --no-extras --cpu-seconds 0.2"
$B 2>/dev/null | python -c "$P" split
TBI_SPLITATT_BWD_ONE_CALL=1 $B 2>/dev/null | python -c "$P" one_call
$B 2>/dev/null | python -c "$P" split
TBI_SPLITATT_BWD_ONE_CALL=1 $B 2>/dev/null | python -c "$P" one_call
