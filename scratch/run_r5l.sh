#!/bin/bash
mkdir -p gpurun_out
P='import json,sys
for l in sys.stdin:
    l=l.strip()
    if l.startswith("{"):
        d=json.loads(l); print(sys.argv[1], d.get("shape_U_r"), "fwd", d.get("fwd_us"), d.get("fwd_frac_of_hbm_peak"), "bwd", d.get("bwd_us"), d.get("bwd_frac_of_hbm_peak"))'
python bench_splitatt.py 2>/dev/null | python -c "$P" base
TBI_LIB=$PWD/scratch/libs/lib_sa3.so python bench_splitatt.py 2>/dev/null | python -c "$P" sa3
TBI_LIB=$PWD/scratch/libs/lib_sa4.so python bench_splitatt.py 2>/dev/null | python -c "$P" sa4
Q='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], d["ms_per_step"], d["value"])'
B="python bench.py --steps 30 --warmup 5 --no-extras --cpu-seconds 0.2"
$B 2>/dev/null | python -c "$Q" base
TBI_LIB=$PWD/scratch/libs/lib_sa3.so $B 2>/dev/null | python -c "$Q" sa3
TBI_LIB=$PWD/scratch/libs/lib_sa4.so $B 2>/dev/null | python -c "$Q" sa4
