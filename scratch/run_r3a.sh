#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_prepare_gpu.py -m gpu -q --maxfail=30 -rf > gpurun_out/t_r3a.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  |timeout" gpurun_out/t_r3a.log | cut -c1-300 | head -20
