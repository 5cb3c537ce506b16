#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_model_gpu.py tests/test_parity_fullres_gpu.py -x -q > gpurun_out/t_r4q.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_r4q.log
tail -4 gpurun_out/t_r4q.log
TBI_TC_XPACK=all timeout 600 python -m pytest tests/test_tc_gpu.py tests/test_model_gpu.py -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/b_r4q.json 2> gpurun_out/b_r4q.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/b_r4q.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
