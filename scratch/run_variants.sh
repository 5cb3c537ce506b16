#!/bin/bash
# microbench + full step for each experimental library build in scratch/libs
cd /root/repo
for v in "$@"; do
  export TBI_LIB=/root/repo/scratch/libs/lib_$v.so
  echo "=== $v"
  for c in stem dstem dc2 cc2 up3; do timeout 120 python scratch/mb_conv.py $c 10 2>&1 | tail -1; done
  timeout 300 python bench.py --steps 10 --warmup 3 --cpu-seconds 0.5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('bench ms_per_step', d['ms_per_step'], 'value', d['value'])"
done
