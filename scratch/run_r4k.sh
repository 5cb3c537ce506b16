#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_tc_gpu.py tests/test_model_gpu.py tests/test_ops_gpu.py tests/test_parity_fullres_gpu.py -x -q > gpurun_out/t_r4k.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_r4k.log
tail -5 gpurun_out/t_r4k.log
python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/b_r4k.json 2> gpurun_out/b_r4k.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/b_r4k.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:smallcin -c 6 --csv --log-file gpurun_out/smallcin_times.csv python bench.py --steps 1 --warmup 1 --no-graph --cpu-seconds 0.2 --no-extras > /dev/null 2>&1
grep -v "^==" gpurun_out/smallcin_times.csv | cut -d, -f5,15 | tail -6
