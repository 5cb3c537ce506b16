#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_fullres_gpu.py -m gpu -q --maxfail=20 -rf -k "cta_pair or conv3x3_layer or convt_layer" > gpurun_out/t_r2l.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  |timeout|trap" gpurun_out/t_r2l.log | head -40
echo PAIR; for c in up4 up3 up2 cc2; do timeout 120 python scratch/mb_conv.py $c 10 2>&1 | tail -1; done
echo NO_PAIR; for c in up4 up3 up2 cc2; do TBI_TC_NO_PAIR=1 timeout 120 python scratch/mb_conv.py $c 10 2>&1 | tail -1; done
