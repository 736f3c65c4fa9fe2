"""Top stall sites of an `ncu --page source --csv` export (SASS view), per kernel launch."""
import csv, sys, collections, re
path = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
kernels = []; cur = None
for row in csv.reader(open(path)):
    if row and row[0] == 'Kernel Name':
        cur = {'name': row[1], 'rows': [], 'hdr': None}; kernels.append(cur); continue
    if cur is None: continue
    if cur['hdr'] is None: cur['hdr'] = row; continue
    cur['rows'].append(row)
print([k['name'][25:60] for k in kernels])
k = kernels[which]; h = {n: i for i, n in enumerate(k['hdr'])}
stalls = [n for n in k['hdr'] if n.startswith('stall_') and 'Not Issued' not in n]
tot = sum(int(r[h['# Samples']]) for r in k['rows']); texec = sum(int(r[h['Instructions Executed']]) for r in k['rows'])
print(k['name'][:80], 'samples', tot, 'warp-instr', texec)
agg = {s: sum(int(r[h[s]]) for r in k['rows']) for s in stalls}
print(sorted(agg.items(), key=lambda kv: -kv[1])[:8])
for i, r in sorted(enumerate(k['rows']), key=lambda ir: -int(ir[1][h['# Samples']]))[:top]:
    st = sorted(((s, int(r[h[s]])) for s in stalls), key=lambda kv: -kv[1])[:2]
    print(f"{i:5d} {r[h['# Samples']]:>6s} {r[h['Instructions Executed']]:>10s}  {r[h['Source']][:80]:80s} {st}")
