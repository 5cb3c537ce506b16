#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/parity_fullres.jsonl
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 -rf > gpurun_out/t_r2t.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/t_r2t.log | cut -c1-300 | head -40
timeout 600 python bench.py --steps 30 --warmup 5 --cpu-seconds 2 > gpurun_out/b_r2t.json 2> gpurun_out/b_r2t.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/b_r2t.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'])
print(json.dumps(d.get('other_configs'), indent=1)[:3500])
PY
tail -3 gpurun_out/b_r2t.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
