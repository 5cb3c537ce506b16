#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"softmax_loss" -c 2 --csv --log-file gpurun_out/loss2_times.csv python bench.py --steps 1 --warmup 1 --no-graph --cpu-seconds 0.2 --no-extras > /dev/null 2>&1
grep -v "^==" gpurun_out/loss2_times.csv | awk -F'","' '{print $NF}' | tail -2
python bench.py --steps 40 --warmup 5 --no-extras --cpu-seconds 0.2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'])"
