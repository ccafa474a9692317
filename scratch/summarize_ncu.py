"""Summarise gpurun_out/launches.csv (+ optional ncu --page raw csv) into profiles/<tag>_*.  Usage:
python scratch/summarize_ncu.py <tag> [launches.csv] [raw.csv]"""
import collections
import csv
import re
import sys

tag = sys.argv[1]
launches = sys.argv[2] if len(sys.argv) > 2 else 'gpurun_out/launches.csv'
raw = sys.argv[3] if len(sys.argv) > 3 else None


def short(n):
    n = re.sub(r'^void ', '', n)
    n = re.sub(r'\(.*', '', n)
    return n.replace('se3et::', '')


rows = list(csv.reader(open(launches)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui, gi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit'), hdr.index('Grid Size')
agg = collections.OrderedDict()
tot = 0.0
lines = []
for r in data:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(',', ''))
    v = v / 1e3 if r[ui] == 'ns' else v * 1e3 if r[ui] == 'ms' else v
    name = short(r[ki])
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
    lines.append('%-48s %-18s %10.1f' % (name[:48], r[gi], v))
with open('profiles/%s_launches_by_kernel.txt' % tag, 'w') as f:
    f.write('# ncu --metrics gpu__time_duration.sum --clock-control none, one launch sequence (cold-cache, serialised)\n')
    f.write('# total %.1f us over %d launches\n' % (tot, sum(a[0] for a in agg.values())))
    f.write('%-48s %6s %12s %7s\n' % ('kernel', 'calls', 'us', 'share'))
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write('%-48s %6d %12.1f %6.1f%%\n' % (k[:48], c, t, 100 * t / tot))
with open('profiles/%s_launch_list.txt' % tag, 'w') as f:
    f.write('%-48s %-18s %10s\n' % ('kernel', 'grid', 'us'))
    f.write('\n'.join(lines) + '\n')
if raw:
    rows = list(csv.reader(open(raw)))
    hdr, data = rows[0], rows[2:]
    want = ['Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
            'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
            'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic']
    want = [w for w in want if w in hdr]
    with open('profiles/%s_ncu_full_summary.csv' % tag, 'w') as f:
        w = csv.writer(f)
        w.writerow(['kernel'] + want)
        w.writerow(['unit'] + [rows[1][hdr.index(x)] for x in want])
        for r in data:
            w.writerow([short(r[hdr.index('Kernel Name')])] + [r[hdr.index(x)] for x in want])
print('ok')
