#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t_r4a.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_r4a.log
tail -3 gpurun_out/t_r4a.log
python bench.py > gpurun_out/b_r4a.json 2> gpurun_out/b_r4a.err; echo "bench rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r4a.log 2>&1; echo "smoke rc=$?"
