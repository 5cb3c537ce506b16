#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_model_gpu.py tests/test_ops_gpu.py -x -q 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:smallcin -c 4 --csv --log-file gpurun_out/smallcin_times.csv python bench.py --steps 1 --warmup 1 --no-graph --cpu-seconds 0.2 --no-extras > /dev/null 2>&1
grep -v "^==" gpurun_out/smallcin_times.csv | tail -4
