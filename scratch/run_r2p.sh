#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_dp_gpu.py -m gpu -q -rf -s > gpurun_out/t_dp2.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  |DP2" gpurun_out/t_dp2.log | cut -c1-900 | head -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --cpu-seconds 1 > gpurun_out/b_dp2.json 2> gpurun_out/b_dp2.err
python -c "
import json; d=json.loads(open('gpurun_out/b_dp2.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'])"
tail -3 gpurun_out/b_dp2.err
