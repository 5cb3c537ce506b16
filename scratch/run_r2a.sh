#!/bin/bash
# round 2, GPU call A: full GPU test suite (incl. the new full-resolution parity tests) + benches of r2k1 / r4k4 / r3k4
mkdir -p gpurun_out; rm -f gpurun_out/parity_fullres.jsonl
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 -rf 2>&1 | tail -150 > gpurun_out/t_r2a.log
tail -5 gpurun_out/t_r2a.log
timeout 300 python bench.py --steps 20 --warmup 5 --cpu-seconds 2 > gpurun_out/b_r2a_r2k1.json 2> gpurun_out/b_r2a_r2k1.err; tail -c 600 gpurun_out/b_r2a_r2k1.json
timeout 300 python bench.py --steps 20 --warmup 5 --cpu-seconds 2 --radix 4 --kpaths 4 > gpurun_out/b_r2a_r4k4.json 2> gpurun_out/b_r2a_r4k4.err; head -c 300 gpurun_out/b_r2a_r4k4.json
timeout 300 python bench.py --steps 20 --warmup 5 --cpu-seconds 2 --radix 3 --kpaths 4 > gpurun_out/b_r2a_r3k4.json 2> gpurun_out/b_r2a_r3k4.err; head -c 300 gpurun_out/b_r2a_r3k4.json
tail -3 gpurun_out/b_r2a_r3k4.err
