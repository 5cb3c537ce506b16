#!/bin/bash
echo "PAIR 2 CTAs/SM"; for c in up4 up3 up2; do timeout 120 python scratch/mb_conv.py $c 10 2>&1 | tail -1; done
echo "PAIR 1 CTA/SM deep rings (200KB)"; for c in up4 up3 up2; do TBI_HALO_BUDGET_KB=196 timeout 120 python scratch/mb_conv.py $c 10 2>&1 | tail -1; done
echo "dgrad-ish: full bench"; timeout 300 python bench.py --steps 30 --warmup 5 --cpu-seconds 1 > gpurun_out/b_r2n.json 2> gpurun_out/b_r2n.err; python -c "
import json; d=json.load(open('gpurun_out/b_r2n.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'])"
tail -2 gpurun_out/b_r2n.err
TBI_TC_NO_PAIR=1 TBI_WGRAD_NO_PAIR=1 timeout 300 python bench.py --steps 30 --warmup 5 --cpu-seconds 1 > gpurun_out/b_r2n_np.json 2> gpurun_out/b_r2n_np.err; python -c "
import json; d=json.load(open('gpurun_out/b_r2n_np.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'])"
