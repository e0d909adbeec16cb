"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name: launches, total us, share."""
import csv, re, sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(r['Metric Value'].replace(',', ''))
    unit = r['Metric Unit']
    us = v / 1e3 if unit in ('ns', 'nsecond') else (v if unit in ('us', 'usecond') else v * 1e3)
    name = re.sub(r'\(.*', '', r['Kernel Name'])
    rows.append((name, us))
tot = sum(u for _, u in rows)
agg = defaultdict(lambda: [0, 0.0])
for n, u in rows:
    agg[n][0] += 1; agg[n][1] += u
print('%d launches, %.1f us total' % (len(rows), tot))
for n, (c, u) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print('%6.2f%%  %9.1f us  %5d x %7.1f us  %s' % (100 * u / tot, u, c, u / c, n[:110]))
