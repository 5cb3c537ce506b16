#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_data_gpu.py -m gpu -q --maxfail=20 -rf > gpurun_out/t_r2s.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/t_r2s.log | cut -c1-500 | head -30
