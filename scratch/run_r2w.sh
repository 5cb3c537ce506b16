#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_fullres.jsonl
timeout 1500 python -m pytest tests/test_parity_fullres_gpu.py tests/test_tc_gpu.py tests/test_model_gpu.py -m gpu -q --maxfail=30 -rf > gpurun_out/t_r2w.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  |timeout" gpurun_out/t_r2w.log | cut -c1-300 | head -40
echo PAIR_RES; for c in stem s16 p32 dstem; do timeout 120 python scratch/mb_conv.py $c 10 2>&1 | tail -1; done
echo NO_PAIR_RES; for c in stem s16 p32 dstem; do TBI_TC_NO_PAIR_RESIDENT=1 timeout 120 python scratch/mb_conv.py $c 10 2>&1 | tail -1; done
timeout 600 python bench.py --steps 30 --warmup 5 --cpu-seconds 1 --no-extras > gpurun_out/b_r2w.json 2> gpurun_out/b_r2w.err; python -c "
import json; d=json.load(open('gpurun_out/b_r2w.json')); print('pair-res on:', d['ms_per_step'], d['value'], d['e2e']['value'])"
TBI_TC_NO_PAIR_RESIDENT=1 timeout 600 python bench.py --steps 30 --warmup 5 --cpu-seconds 1 --no-extras > gpurun_out/b_r2w0.json 2> gpurun_out/b_r2w0.err; python -c "
import json; d=json.load(open('gpurun_out/b_r2w0.json')); print('pair-res off:', d['ms_per_step'], d['value'], d['e2e']['value'])"
