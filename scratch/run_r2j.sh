#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_fullres_gpu.py -m gpu -q --maxfail=20 -rf -k "wgrad_cta_pairs or convt_layer" > gpurun_out/t_r2j.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  |timeout|trap" gpurun_out/t_r2j.log | head -40
echo "T=2 (2 pairs per SM pair)"; for c in wup4 wup3 wup2 wcc3; do timeout 120 python scratch/mb_conv.py $c 10 2>&1 | tail -1; done
echo "T=4"; for c in wup4 wup3 wup2; do TBI_WGRAD_PAIR_T=4 timeout 120 python scratch/mb_conv.py $c 10 2>&1 | tail -1; done
