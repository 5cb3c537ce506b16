#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scratch/dbg_r3k4.py 64 > gpurun_out/dbg_r3k4_64.log 2>&1; tail -60 gpurun_out/dbg_r3k4_64.log
timeout 600 python scratch/dbg_r3k4.py 256 > gpurun_out/dbg_r3k4_256.log 2>&1; tail -60 gpurun_out/dbg_r3k4_256.log
