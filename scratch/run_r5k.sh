#!/bin/bash
mkdir -p gpurun_out
S=$(date +%s)
python bench.py > gpurun_out/b_r5k.json 2> gpurun_out/b_r5k.err; echo "bench rc=$? secs=$(( $(date +%s) - S ))"
S=$(date +%s)
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/b_r5k_ref.json 2> gpurun_out/b_r5k_ref.err; echo "ref rc=$? secs=$(( $(date +%s) - S ))"
for l in 0 1 2 3 4; do TBI_HALO_PF_LAG=$l python bench.py --steps 30 --warmup 5 --no-extras --cpu-seconds 0.2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lag $l', d['ms_per_step'])"; done
