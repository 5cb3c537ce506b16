"""summarise one training step out of an ncu launch list: python scratch/step_list.py <csv> [filter]"""
import csv, re, sys
from collections import defaultdict
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith('==')]
for row in csv.DictReader(lines):
    rows.append(row)
names = [row['Kernel Name'] for row in rows]
idx = [i for i, n in enumerate(names) if 'adam_dev' in n]
a, b = idx[-2], idx[-1]
flt = sys.argv[2] if len(sys.argv) > 2 else None
tot = 0; agg = defaultdict(lambda: [0.0, 0])
for i in range(a + 1, b + 1):
    row = rows[i]
    d = float(row['Metric Value'].replace(',', '')) / 1000
    tot += d
    nm = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('<unnamed>::', '')[:48]
    agg[nm][0] += d; agg[nm][1] += 1
    if flt and re.search(flt, nm):
        print(f"{i-a:4d} {d:8.1f} {nm:48s} grid={row['Grid Size']}")
print('total', round(tot, 1), 'launches', b - a)
for nm, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{t:8.1f} {100*t/tot:5.1f}% n={n:3d} {nm}")
