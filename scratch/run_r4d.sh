#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_tc_gpu.py tests/test_reference_fixture_gpu.py tests/test_vit_gpu.py -x -q > gpurun_out/t_r4d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_r4d.log
tail -25 gpurun_out/t_r4d.log
for c in wstem; do timeout 120 python scratch/mb_conv.py $c 10 2>&1 | tail -1; done
python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/b_r4d.json 2> gpurun_out/b_r4d.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/b_r4d.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
