#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/b_r4h.json 2> gpurun_out/b_r4h.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/b_r4h.json').read()); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
python -m pytest tests/test_tc_gpu.py tests/test_model_gpu.py -x -q 2>&1 | tail -3
ncu --set full --clock-control none --import-source on -k regex:"smallcin_fwd|smallcin_wgrad" -c 2 -f -o gpurun_out/r4_smallcin python bench.py --steps 1 --warmup 0 --no-graph --cpu-seconds 0.2 --no-extras > gpurun_out/ncu_smallcin.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"tapgemm_halo_kernel<32" -c 3 -f -o gpurun_out/r4_halo32 python bench.py --steps 1 --warmup 0 --no-graph --cpu-seconds 0.2 --no-extras > gpurun_out/ncu_halo32.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
