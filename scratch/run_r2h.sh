#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_fullres_gpu.py -m gpu -q --maxfail=20 -rf -k "wgrad_cta_pairs or conv3x3_layer or convt_layer" > gpurun_out/t_r2h.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  |timeout|trap" gpurun_out/t_r2h.log | head -40
for c in wup4 wup3 wup2 wcc3; do timeout 120 python scratch/mb_conv.py $c 10 2>&1 | tail -1; done
for c in wup4 wup3 wup2 wcc3; do TBI_WGRAD_NO_PAIR=1 timeout 120 python scratch/mb_conv.py $c 10 2>&1 | tail -1; done
