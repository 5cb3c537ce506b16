import csv, collections, re, sys
path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/launches_r1.csv'
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
r = csv.DictReader(lines)
agg = collections.defaultdict(lambda: [0, 0.0, 0.0]); tot = 0
for row in r:
    name = row['Kernel Name']; v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    if unit == 'ns': v /= 1e3
    elif unit == 'ms': v *= 1e3
    name = re.sub(r'^void ', '', re.sub(r'\(.*', '', name)).replace('<unnamed>::', '')
    agg[name][0] += 1; agg[name][1] += v; agg[name][2] = max(agg[name][2], v); tot += v
print("total us %.0f launches %d" % (tot, sum(a[0] for a in agg.values())))
for k, (c, t, m) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 22]:
    print("%8.1f us %5.1f%% n=%4d avg %7.1f max %7.1f  %s" % (t, 100*t/tot, c, t/c, m, k[:90]))
