#!/bin/bash
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], d["ms_per_step"], d["value"])'
B="python bench.py --steps 40 --warmup 5 --no-extras --cpu-seconds 0.2"
$B 2>/dev/null | python -c "$P" base
TBI_EXPERIMENT_SKIP_COLSUM=1 $B 2>/dev/null | python -c "$P" skip_colsum
$B 2>/dev/null | python -c "$P" base
TBI_EXPERIMENT_SKIP_COLSUM=1 $B 2>/dev/null | python -c "$P" skip_colsum
