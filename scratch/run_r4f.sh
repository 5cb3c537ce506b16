#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_tc_gpu.py tests/test_reference_fixture_gpu.py -x -q > gpurun_out/t_r4f.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_r4f.log
tail -5 gpurun_out/t_r4f.log
for c in wstem wup4 wup3 wup2 wcc3; do timeout 120 python scratch/mb_conv.py $c 10 2>&1 | tail -1; done
python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/b_r4f.json 2> gpurun_out/b_r4f.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/b_r4f.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
