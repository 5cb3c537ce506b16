#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_fullres.jsonl
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 -rf > gpurun_out/t_r2k.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/t_r2k.log | head -40
TBI_TC_NO_DUAL=1 timeout 300 python bench.py --steps 30 --warmup 5 --cpu-seconds 1 > gpurun_out/b_r2k.json 2> gpurun_out/b_r2k.err; python -c "
import json; d=json.load(open('gpurun_out/b_r2k.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'])"
TBI_TC_NO_DUAL=1 TBI_WGRAD_NO_PAIR=1 timeout 300 python bench.py --steps 30 --warmup 5 --cpu-seconds 1 > gpurun_out/b_r2k_np.json 2> gpurun_out/b_r2k_np.err; python -c "
import json; d=json.load(open('gpurun_out/b_r2k_np.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'])"
