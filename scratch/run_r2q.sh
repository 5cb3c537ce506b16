#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_variant_b_gpu.py -m gpu -q --maxfail=20 -rf -s > gpurun_out/t_r2q.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  |largest relative|bf16 storage alone|encoder rel|decoder rel" gpurun_out/t_r2q.log | cut -c1-400 | head -40
