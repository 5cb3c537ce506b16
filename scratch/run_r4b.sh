#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_reference_fixture_gpu.py tests/test_vit_gpu.py tests/test_data_gpu.py -x -q > gpurun_out/t_r4b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_r4b.log
tail -30 gpurun_out/t_r4b.log
python bench.py --steps 10 --warmup 3 > gpurun_out/b_r4b.json 2> gpurun_out/b_r4b.err; echo "bench rc=$?"
tail -5 gpurun_out/b_r4b.err
