#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_reference_fixture_gpu.py tests/test_vit_gpu.py -x -q > gpurun_out/t_r4c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_r4c.log
tail -30 gpurun_out/t_r4c.log
python scratch/vb_graph.py 2>&1 | tail -20
