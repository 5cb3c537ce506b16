#!/bin/bash
mkdir -p gpurun_out
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], d["ms_per_step"], d["value"], d["e2e"]["value"], d["clocks"])'
python bench.py --steps 50 --warmup 5 --no-extras --cpu-seconds 0.2 2>/dev/null | python -c "$P" n1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 --no-extras --cpu-seconds 0.2 2>gpurun_out/n2.err | python -c "$P" n2
timeout 600 python -m pytest tests/test_dp_gpu.py -x -q 2>&1 | tail -2
