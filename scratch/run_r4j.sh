#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_tc_gpu.py tests/test_model_gpu.py tests/test_ops_gpu.py tests/test_parity_fullres_gpu.py -x -q > gpurun_out/t_r4j.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_r4j.log
tail -5 gpurun_out/t_r4j.log
python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/b_r4j.json 2> gpurun_out/b_r4j.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/b_r4j.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
bash scratch/run_r4e.sh
