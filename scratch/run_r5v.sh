#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/b_final.json 2> gpurun_out/b_final.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/b_final.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['ms_per_call'], d['roofline']['frac'], d['roofline']['step_frac_of_peak_sustained'], d['gpu_launches'], d['clocks'])"
