#!/bin/bash
python scratch/bench_radius.py 2>&1 | tail -4
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/radius_launches.csv python scratch/bench_radius.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/radius_launches.csv')))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr = rows[hi]; ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r) <= vi: continue
    v = float(r[vi].replace(',', '')); v = v / 1e3 if r[ui] == 'ns' else v
    n = r[ki].split('(')[0][-40:]
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:16]:
    print('%-42s %4d %10.1f us  %8.1f us/call' % (k, c, t, t / c))
PY
