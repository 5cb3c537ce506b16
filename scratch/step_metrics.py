"""per-kernel table (time, DRAM bytes, instructions, issue %, tensor %) of the last full step in an ncu --metrics csv"""
import csv, re, sys
from collections import OrderedDict
lines=[l for l in open(sys.argv[1]) if not l.startswith('==')]
rows=list(csv.DictReader(lines))
byid=OrderedDict()
for r in rows:
    k=int(r['ID'])
    d=byid.setdefault(k, {'name': re.sub(r'\(.*','',r['Kernel Name']).replace('void ','').replace('<unnamed>::','')[:44], 'grid': r['Grid Size']})
    d[r['Metric Name']]=float(r['Metric Value'].replace(',',''))
    d['unit_'+r['Metric Name']]=r['Metric Unit']
ids=list(byid)
ad=[i for i in ids if 'adam_dev' in byid[i]['name']]
a,b=ad[-2],ad[-1]
thr=float(sys.argv[2]) if len(sys.argv)>2 else 40
def mb(v,u): return v/1e6 if u=='byte' else v*1e3 if u=='Gbyte' else v if u=='Mbyte' else v/1e3 if u=='Kbyte' else v
tot=0; exc=0
print(f"{'#':>4} {'us':>7} {'rdMB':>7} {'wrMB':>7} {'TB/s':>5} {'floor':>6} {'Minst':>6} {'iss%':>5} {'tc%':>5}  name")
for i in ids:
    if not (a < i <= b): continue
    d=byid[i]
    t=d['gpu__time_duration.sum']
    t_us = t/1000 if d['unit_gpu__time_duration.sum'].startswith('n') else t
    rdm=mb(d['dram__bytes_read.sum'],d['unit_dram__bytes_read.sum']); wrm=mb(d['dram__bytes_write.sum'],d['unit_dram__bytes_write.sum'])
    tot+=t_us
    floor=(rdm+wrm)/6.5488
    if t_us>=thr:
        print(f"{i-a:4d} {t_us:7.1f} {rdm:7.1f} {wrm:7.1f} {(rdm+wrm)/t_us/1e3:5.2f} {floor:6.1f} {d['sm__inst_executed.sum']/1e6:6.1f} {d['smsp__issue_active.avg.pct_of_peak_sustained_active']:5.1f} {d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']:5.1f}  {d['name']} {d['grid']}")
print('total', round(tot,1), 'launches', b-a)
