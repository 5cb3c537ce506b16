#!/bin/bash
mkdir -p gpurun_out
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], d["ms_per_step"], d["value"], d["e2e"]["value"])'
A="--steps 40 --warmup 5 --no-extras --cpu-seconds 0.2"
python bench.py $A 2>/dev/null | python -c "$P" n1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$T --master-port 29511 bench.py --gpus 2 $A 2>gpurun_out/n2_bf16.err | python -c "$P" n2_bf16comm
TBI_DP_FP32_COMM=1 $T --master-port 29512 bench.py --gpus 2 $A 2>/dev/null | python -c "$P" n2_fp32comm
$T --master-port 29513 bench.py --gpus 2 $A 2>/dev/null | python -c "$P" n2_bf16comm
timeout 600 python -m pytest tests/test_dp_gpu.py -x -q 2>&1 | tail -2
