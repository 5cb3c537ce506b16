#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_fullres.jsonl
timeout 1500 python -m pytest tests/test_model_gpu.py tests/test_parity_fullres_gpu.py tests/test_tc_gpu.py -m gpu -q --maxfail=30 -rf > gpurun_out/t_r2v.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/t_r2v.log | cut -c1-300 | head -40
for m in 1 0; do TBI_HEAD_FWD_GEMM=$m timeout 600 python bench.py --steps 30 --warmup 5 --cpu-seconds 1 --no-extras > gpurun_out/b_r2v_$m.json 2> gpurun_out/b_r2v_$m.err; python -c "
import json; d=json.load(open('gpurun_out/b_r2v_$m.json')); print('head gemm $m:', d['ms_per_step'], d['value'], d['e2e']['value'])"; done
