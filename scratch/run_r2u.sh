#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dp_gpu.py -m gpu -q -rf > gpurun_out/t_dp2b.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/t_dp2b.log | cut -c1-400 | head
for mode in cuts mb54; do
  if [ $mode = mb54 ]; then export TBI_BUCKET_MB=54; else unset TBI_BUCKET_MB; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 5 --cpu-seconds 1 > gpurun_out/b_dp2_$mode.json 2> gpurun_out/b_dp2_$mode.err
  python -c "
import json; d=json.loads(open('gpurun_out/b_dp2_$mode.json').read().strip().splitlines()[-1]); print('$mode', d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'])"
done
