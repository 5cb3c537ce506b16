#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_fullres_gpu.py tests/test_tc_gpu.py -m gpu -q --maxfail=10 -rf -k "not whole_graph or 2-1" > gpurun_out/t_r2e.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/t_r2e.log | head -40
for c in up4 up3 up2 cc2; do timeout 120 python scratch/mb_conv.py $c 10 2>&1 | tail -1; done
for c in up4 up3 up2 cc2; do TBI_TC_NO_DUAL=1 timeout 120 python scratch/mb_conv.py $c 10 2>&1 | tail -1; done
timeout 300 python bench.py --steps 20 --warmup 5 --cpu-seconds 1 > gpurun_out/b_r2e.json 2> gpurun_out/b_r2e.err; python -c "
import json; d=json.load(open('gpurun_out/b_r2e.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'])"
TBI_TC_NO_DUAL=1 timeout 300 python bench.py --steps 20 --warmup 5 --cpu-seconds 1 > gpurun_out/b_r2e_nd.json 2> gpurun_out/b_r2e_nd.err; python -c "
import json; d=json.load(open('gpurun_out/b_r2e_nd.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'])"
tail -3 gpurun_out/b_r2e.err
