#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_model_gpu.py tests/test_parity_fullres_gpu.py tests/test_reference_fixture_gpu.py tests/test_prepare_gpu.py -x -q 2>&1 | tail -3
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], d["ms_per_step"], d["value"], d["e2e"]["value"])'
B="python bench.py --steps 50 --warmup 5 --no-extras --cpu-seconds 0.2"
$B 2>/dev/null | python -c "$P" overlap
TBI_NO_ADAM_OVERLAP=1 $B 2>/dev/null | python -c "$P" no_overlap
$B 2>/dev/null | python -c "$P" overlap
TBI_NO_ADAM_OVERLAP=1 $B 2>/dev/null | python -c "$P" no_overlap
