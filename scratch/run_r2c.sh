#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_fullres.jsonl
timeout 1500 python -m pytest tests/test_parity_fullres_gpu.py tests/test_dp_gpu.py -m gpu -q --maxfail=30 -rf 2>&1 > gpurun_out/t_r2c.log
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/t_r2c.log | head -60
