#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_vit_gpu.py -m gpu -q --maxfail=20 -rf -s > gpurun_out/t_r2r.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  |largest relative|probs rel|bf16 whole" gpurun_out/t_r2r.log | cut -c1-500 | head -50
