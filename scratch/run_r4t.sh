#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_prepare_gpu.py tests/test_model_gpu.py -x -q 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"prepare_kernel|softmax_loss|convt_gather4|colsum_kernel|convt_scatter4" -c 12 --csv --log-file gpurun_out/prep_times.csv python bench.py --steps 1 --warmup 1 --no-graph --cpu-seconds 0.2 --no-extras > /dev/null 2>&1
python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/b_r4t.json 2> gpurun_out/b_r4t.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/b_r4t.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_call'])"
