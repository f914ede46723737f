#!/usr/bin/env python
"""Summarise an ncu --csv launch list (gpu__time_duration.sum etc.) per kernel and grid size."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, gi, mi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size'), hdr.index('Metric Name')
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    agg.setdefault((r[ki][:34], r[gi], r[mi]), []).append(v)
pat = sys.argv[2] if len(sys.argv) > 2 else ''
for (k, g, m), v in agg.items():
    if pat in k:
        print(f"{k:34s} {g:16s} {m[:34]:34s} n={len(v):3d} avg={sum(v) / len(v):16.1f} total={sum(v):16.1f}")
