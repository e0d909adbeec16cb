"""Aggregates an `ncu --csv` launch list (metrics: gpu__time_duration.sum, sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,
dram__bytes_read.sum, dram__bytes_write.sum) by kernel family:  launches, time, share of the total, time-weighted tensor-pipe
active %, DRAM GB/s.      python tests/ncu_by_kernel.py launches.csv [title] > profiles/<name>.txt"""
import collections
import csv
import io
import re
import sys

lines = open(sys.argv[1]).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
rows = list(csv.DictReader(io.StringIO('\n'.join(lines[start:]))))
per = collections.OrderedDict()
for r in rows:
    d = per.setdefault(r['ID'], {'name': re.sub(r'\(.*', '', r['Kernel Name']).replace('void ', '')})
    v = float(r['Metric Value'].replace(',', '')) if r['Metric Value'] not in ('', 'n/a') else 0.0
    u, m = r['Metric Unit'], r['Metric Name']
    if m == 'gpu__time_duration.sum':
        d['us'] = v / 1e3 if u in ('ns', 'nsecond') else (v if u in ('us', 'usecond') else v * 1e3)
    elif m.startswith('sm__pipe_tensor_cycles_active'):
        d['tensor'] = v
    elif m.startswith('dram__bytes'):
        mult = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
        d['bytes'] = d.get('bytes', 0.0) + v * mult
fam = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
total = 0.0
for d in per.values():
    name = re.sub(r'<.*', '', d['name']) if d['name'].startswith('at::') else d['name']
    f = fam[name]
    us = d.get('us', 0.0)
    f[0] += 1; f[1] += us; f[2] += d.get('tensor', 0.0) * us; f[3] += d.get('bytes', 0.0)
    total += us
print('%s: %d launches, %.1f us total (ncu gpu__time_duration.sum, --clock-control none; serialised, cold caches: compare shares)'
      % (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1], len(per), total))
print('%7s %10s %5s %9s %9s %9s  %s' % ('share', 'time', 'n', 'us/launch', 'tensor%', 'DRAM GB/s', 'kernel'))
for name, (n, us, tw, by) in sorted(fam.items(), key=lambda x: -x[1][1]):
    print('%6.2f%% %8.1f us %5d %9.1f %9.1f %9.0f  %s' % (100 * us / total, us, n, us / n, tw / us if us else 0.0, by / us / 1e3 if us else 0.0, name))
