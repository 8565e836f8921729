#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: stall mix, hottest SASS instructions, instruction count.
usage: tools/ncu_src.py file.csv [top_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
data = [dict(zip(h, r)) for r in rows[hi + 1:] if len(r) == len(h) and r[0] != "Address"]
def I(x):
    try: return int(x)
    except ValueError: return 0
tot = sum(I(d['# Samples']) for d in data)
ex = sum(I(d['Instructions Executed']) for d in data)
print('sass instrs', len(data), 'samples', tot, 'warp-instr executed', ex)
stalls = [k for k in h if k.startswith('stall_') and 'Not Issued' not in k]
agg = {k: sum(I(d[k]) for d in data) for k in stalls}
for k, v in sorted(agg.items(), key=lambda x: -x[1])[:10]:
    print(f'  {k:28s} {v:8d} {100 * v / max(tot, 1):5.1f}%')
for d in sorted(data, key=lambda d: -I(d['# Samples']))[:topn]:
    s = I(d['# Samples'])
    print(f"{s:6d} {100 * s / max(tot, 1):5.1f}% ex={d['Instructions Executed']:>8s} {d['Source'].strip()[:80]:80s}", {k[6:]: d[k] for k in stalls if I(d[k]) > s * 0.3})
