"""Executed warp-instructions by SASS opcode for one kernel of an `ncu --page source --csv` export."""
import csv, sys
path = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = list(csv.reader(open(path)))
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == 'Address']
i0 = hdr_i[which]; i1 = hdr_i[which + 1] - 1 if which + 1 < len(hdr_i) else len(rows)
print(rows[i0 - 1][:2])
hdr = rows[i0]; data = rows[i0 + 1:i1]
si = hdr.index('Source'); ai = hdr.index('Warp Stall Sampling (All Samples)'); ei = hdr.index('Instructions Executed')
ops = {}; smp = {}
for r in data:
    t = r[si].split()
    op = t[1] if t[0].startswith('@') else t[0]
    op = op.rstrip(';')
    base = '.'.join(op.split('.')[:2]) if op.split('.')[0] in ('LDSM', 'STSM', 'LDGSTS', 'STS', 'LDS', 'SYNCS', 'BAR') else op.split('.')[0]
    ops[base] = ops.get(base, 0) + int(r[ei]); smp[base] = smp.get(base, 0) + int(r[ai])
tot = sum(ops.values()); ts = sum(smp.values())
print('total executed warp-instructions', tot, 'samples', ts)
for k, v in sorted(ops.items(), key=lambda x: -x[1])[:40]:
    print('%-14s %12d %5.1f%%   samples %5.1f%%' % (k, v, 100 * v / tot, 100 * smp[k] / ts))
