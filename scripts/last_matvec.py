#!/usr/bin/env python
"""Per-launch metrics of the last H*v in an ncu --csv launch list (star kernels)."""
import csv
import sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, gi, mi, ii = (hdr.index(c) for c in ('Kernel Name', 'Metric Value', 'Grid Size', 'Metric Name', 'ID'))
cur = {}
for r in rows[1:]:
    if 'star_up' in r[ki] or 'star_dw' in r[ki]:
        cur.setdefault(r[ii], {'k': r[ki][7:16], 'g': r[gi]})[r[mi][:12]] = float(r[vi].replace(',', ''))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 18
tot = 0
for i in sorted(cur, key=int)[-n:]:
    d = cur[i]
    tot += d.get('gpu__time_du', 0)
    print(i, d['k'], d['g'], {k: round(v, 1) for k, v in d.items() if k not in ('k', 'g')})
print("total ns", tot)
