#!/bin/bash
mkdir -p gpurun_out
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 800 compute-sanitizer --tool initcheck --print-limit 40 python scratch/vit_init.py > gpurun_out/initcheck.log 2>&1
grep -c "Uninitialized" gpurun_out/initcheck.log
grep -A12 "Uninitialized" gpurun_out/initcheck.log | grep -E "Uninitialized|at .*\(|by thread" | head -40
tail -5 gpurun_out/initcheck.log
