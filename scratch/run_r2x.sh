#!/bin/bash
for v in "" "TBI_EXP_SKIP_PREPARE=1" "TBI_EXP_NO_COLSUM=1" "TBI_EXP_SKIP_PREPARE=1 TBI_EXP_NO_COLSUM=1"; do
  env $v timeout 600 python bench.py --steps 30 --warmup 5 --cpu-seconds 0.5 --no-extras > gpurun_out/b_x.json 2> gpurun_out/b_x.err; python -c "
import json; d=json.load(open('gpurun_out/b_x.json')); print('[$v]', d['ms_per_step'], d['value'])"; tail -1 gpurun_out/b_x.err
done
