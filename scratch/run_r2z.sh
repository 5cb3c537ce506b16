#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_fullres.jsonl
timeout 1500 python -m pytest tests/test_prepare_gpu.py tests/test_model_gpu.py tests/test_parity_fullres_gpu.py -m gpu -q --maxfail=30 -rf -k "not layer and not pairs" > gpurun_out/t_r2z.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  |timeout" gpurun_out/t_r2z.log | cut -c1-300 | head -40
timeout 600 python bench.py --steps 30 --warmup 5 --cpu-seconds 1 --no-extras > gpurun_out/b_r2z.json 2> gpurun_out/b_r2z.err; python -c "
import json; d=json.load(open('gpurun_out/b_r2z.json')); print('now:', d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'])"
tail -2 gpurun_out/b_r2z.err
