"""aggregate one training step out of an ncu gpu__time_duration launch list (raw ncu csv): the rows between the last two adam launches"""
import csv, collections, re, sys
path = sys.argv[1]
rows = [r for r in csv.DictReader(l for l in open(path) if not l.startswith('=='))]
names = [r['Kernel Name'] for r in rows]
ends = [i + 1 for i, n in enumerate(names) if 'adam_dev_kernel' in n or 'adam_kernel' in n]      # Adam, then its counter advance
lo, hi = (ends[-2] + 1, ends[-1] + 1) if len(ends) >= 2 else (0, len(rows))
step = rows[lo:hi]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0]); tot = 0
for r in step:
    v = float(r['Metric Value'].replace(',', '')); u = r['Metric Unit']
    v = v / 1e3 if u in ('ns', 'nsecond') else v * 1e3 if u in ('ms', 'msecond') else v
    name = re.sub(r'^void ', '', re.sub(r'\(.*', '', r['Kernel Name'])).replace('<unnamed>::', '')
    a = agg[name]; a[0] += 1; a[1] += v; a[2] = max(a[2], v); tot += v
print("one step: total us %.0f launches %d" % (tot, len(step)))
for k, (c, t, m) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print("%8.1f us %5.1f%% n=%4d avg %7.1f max %7.1f  %s" % (t, 100 * t / tot, c, t / c, m, k[:100]))
