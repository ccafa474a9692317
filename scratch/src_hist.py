import csv, sys
path = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0; B = int(sys.argv[3]) if len(sys.argv) > 3 else 120
rows = list(csv.reader(open(path)))
hdr_i = [i for i,r in enumerate(rows) if r and r[0]=='Address']
i0 = hdr_i[which]; i1 = hdr_i[which+1]-1 if which+1 < len(hdr_i) else len(rows)
print(rows[i0-1][:2])
hdr = rows[i0]; data = rows[i0+1:i1]
si = hdr.index('Source'); ai = hdr.index('Warp Stall Sampling (All Samples)'); ei = hdr.index('Instructions Executed')
cols = {h:j for j,h in enumerate(hdr)}
tot = sum(int(r[ai]) for r in data)
print('instrs', len(data), 'samples', tot)
agg = {}
for h,j in cols.items():
    if h.startswith('stall_') and 'Not Issued' not in h:
        agg[h] = sum(int(r[j]) for r in data if r[j].isdigit())
print(sorted(agg.items(), key=lambda x:-x[1])[:8])
for b in range(0, len(data), B):
    chunk = data[b:b+B]
    s = sum(int(r[ai]) for r in chunk)
    if s < tot * 0.002: continue
    ex = max(int(r[ei]) for r in chunk)
    ops = {}
    for r in chunk:
        t = r[si].split()
        op = t[1] if t[0].startswith('@') else t[0]
        op = op.split('.')[0]
        ops[op] = ops.get(op,0)+1
    key = [k for k in ('HMMA','LDGSTS','ST','STS','MUFU','LDSM','SYNCS','LDTM','UTCHMMA','UTMALDG','LDG','STL','LDL','LDS','SHFL','ATOM','RED','BAR') if k in ops]
    print('%5d %6d %5.1f%% maxexec %10d %s' % (b, s, 100*s/tot, ex, ' '.join('%s:%d'%(k,ops[k]) for k in key)))
top = sorted(data, key=lambda r: -int(r[ai]))[:14]
for r in top:
    print(r[ai].rjust(7), r[ei].rjust(10), r[si][:90])
