#!/usr/bin/env python
"""tools/launch_summary.py file.csv: per-kernel count / average / maximum duration of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hi]; data = [dict(zip(h, r)) for r in rows[hi + 1:] if len(r) == len(h)]
agg = collections.OrderedDict()
for d in data:
    k = d["Kernel Name"][:70]; v = float(d["Metric Value"].replace(",", ""))
    a = agg.setdefault(k, [0, 0.0, 0.0]); a[0] += 1; a[1] += v; a[2] = max(a[2], v)
for k, (n, v, m) in agg.items():
    print(f"{k:72s} n={n:3d} avg_us={v / n / 1e3:10.1f} max_us={m / 1e3:10.1f}")
