#!/usr/bin/env python
"""Top stall sites of one kernel from an `ncu --page source --csv` export (SASS view).
usage: ncu_src_top.py file.csv [kernel-substring] [occurrence] [topN]"""
import csv, sys
path = sys.argv[1]; sub = sys.argv[2] if len(sys.argv) > 2 else ''; occ = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
secs = []; cur = None
for row in csv.reader(open(path)):
    if row and row[0] == 'Kernel Name':
        cur = {'name': row[1], 'rows': [], 'hdr': None}; secs.append(cur); continue
    if cur is None: continue
    if cur['hdr'] is None: cur['hdr'] = row; continue
    cur['rows'].append(row)
secs = [s for s in secs if sub in s['name']]
s = secs[occ]
h = {n: i for i, n in enumerate(s['hdr'])}
rows = s['rows']
tot = sum(int(r[h['# Samples']]) for r in rows)
inst = sum(int(r[h['Instructions Executed']]) for r in rows)
print(s['name'], 'samples', tot, 'warp-inst', inst, 'SASS lines', len(rows))
stalls = [n for n in s['hdr'] if n.startswith('stall_') and 'Not Issued' not in n]
agg = {n: sum(int(r[h[n]]) for r in rows) for n in stalls}
print('stall totals:', ', '.join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
order = sorted(range(len(rows)), key=lambda i: -int(rows[i][h['# Samples']]))[:top]
for i in sorted(order):
    r = rows[i]
    st = {n[6:]: int(r[h[n]]) for n in stalls if int(r[h[n]])}
    st = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"{i:5d} {int(r[h['# Samples']]):6d} {100*int(r[h['# Samples']])/tot:5.1f}% ex={r[h['Instructions Executed']]:>8s} {r[h['Source']].strip()[:70]:70s} {st}")
