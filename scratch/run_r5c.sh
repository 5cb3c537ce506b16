#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_model_gpu.py -x -q 2>&1 | tail -3
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], d["ms_per_step"], d["value"], d["e2e"]["value"])'
python bench.py --steps 30 --warmup 5 --no-extras --cpu-seconds 0.2 2>/dev/null | python -c "$P" n1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --no-extras --cpu-seconds 0.2 2>gpurun_out/n2.err | python -c "$P" n2
TBI_BUCKET_MB=54 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 --no-extras --cpu-seconds 0.2 2>/dev/null | python -c "$P" n2_54mb
NCCL_MAX_NCHANNELS=8 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 30 --warmup 5 --no-extras --cpu-seconds 0.2 2>/dev/null | python -c "$P" n2_8ch
