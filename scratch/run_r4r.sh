#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_model_gpu.py tests/test_parity_fullres_gpu.py -x -q > gpurun_out/t_r4r.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_r4r.log
tail -4 gpurun_out/t_r4r.log
python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/b_r4r.json 2> gpurun_out/b_r4r.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/b_r4r.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
for c in wstem; do timeout 120 python scratch/mb_conv.py $c 10 2>&1 | tail -1; done
