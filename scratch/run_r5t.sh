#!/bin/bash
mkdir -p gpurun_out
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], d["ms_per_step"], d["value"], d["e2e"]["value"])'
A="--steps 40 --warmup 5 --no-extras --cpu-seconds 0.2"
python bench.py $A 2>/dev/null | python -c "$P" n1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$T --master-port 29511 bench.py --gpus 2 $A 2>/dev/null | python -c "$P" n2_default
TBI_BUCKET_CUTS=conv3_2/c1/w $T --master-port 29512 bench.py --gpus 2 $A 2>/dev/null | python -c "$P" n2_cut_conv3_2
TBI_BUCKET_CUTS=conv4_1/c1/w $T --master-port 29513 bench.py --gpus 2 $A 2>/dev/null | python -c "$P" n2_cut_conv4_1
TBI_BUCKET_CUTS=conv3_1/c1/w $T --master-port 29514 bench.py --gpus 2 $A 2>/dev/null | python -c "$P" n2_cut_conv3_1
TBI_BUCKET_MB=400 $T --master-port 29515 bench.py --gpus 2 $A 2>/dev/null | python -c "$P" n2_one_bucket
