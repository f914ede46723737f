#!/usr/bin/env python
"""Table of the launches of ONE H*v from an ncu --csv launch list taken with
   --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
Usage: hxv_launch_table.py launches.csv [n_launches_per_hxv] [traffic.json key]
Prints the last complete H*v (star/fringe kernels) and writes the DRAM traffic per H*v into profiles/traffic.json."""
import csv
import json
import os
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, gi, mi, ii, ui = (hdr.index(c) for c in ('Kernel Name', 'Metric Value', 'Grid Size', 'Metric Name', 'ID', 'Metric Unit'))
cur = {}
for r in rows[1:]:
    if 'k_star_' in r[ki] or 'k_fringe_' in r[ki]:
        d = cur.setdefault(int(r[ii]), {'k': r[ki].split('(')[0].replace('void ', ''), 'g': r[gi]})
        v = float(r[vi].replace(',', ''))
        u = r[ui]
        if r[mi].startswith('dram__bytes'):
            v *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
        elif r[mi].startswith('gpu__time'):
            v *= {'ns': 1e-3, 'us': 1, 'usecond': 1, 'ms': 1e3, 'msecond': 1e3, 'nsecond': 1e-3}.get(u, 1e-3)
        d[r[mi]] = v
ids = sorted(cur)
# one H*v = from a k_fringe_dw / first k_star_dw launch up to the launch before the next one
starts = [i for n, i in enumerate(ids) if cur[i]['k'].startswith(('k_fringe_dw', 'k_star_dw')) and
          (n == 0 or not cur[ids[n - 1]]['k'].startswith(('k_fringe_dw', 'k_star_dw')))]
a, b = starts[-2], starts[-1]
sel = [i for i in ids if a <= i < b]
tot_t = sum(cur[i]['gpu__time_duration.sum'] for i in sel)
tot_r = sum(cur[i].get('dram__bytes_read.sum', 0) for i in sel)
tot_w = sum(cur[i].get('dram__bytes_write.sum', 0) for i in sel)
print(f"{'kernel':28s} {'grid':14s} {'time_us':>9s} {'share':>7s} {'dram_rd_MB':>11s} {'dram_wr_MB':>11s}")
for i in sel:
    d = cur[i]
    t = d['gpu__time_duration.sum']
    print(f"{d['k'][:28]:28s} {d['g']:14s} {t:9.1f} {100 * t / tot_t:6.1f}% {d.get('dram__bytes_read.sum', 0) / 1e6:11.1f} {d.get('dram__bytes_write.sum', 0) / 1e6:11.1f}")
print(f"{'TOTAL per H*v':28s} {len(sel):<14d} {tot_t:9.1f} {100.0:6.1f}% {tot_r / 1e6:11.1f} {tot_w / 1e6:11.1f}")
up = sum(cur[i]['gpu__time_duration.sum'] for i in sel if '_up' in cur[i]['k'])
print(f"# share of the up pass = {100 * up / tot_t:.1f} %, of the down pass = {100 * (1 - up / tot_t):.1f} %")
print(f"# DRAM traffic per H*v = {(tot_r + tot_w) / 1e9:.3f} GB")
if len(sys.argv) > 3:
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'profiles', 'traffic.json')
    t = json.load(open(p)) if os.path.exists(p) else {}
    t[sys.argv[3]] = tot_r + tot_w
    json.dump(t, open(p, 'w'), indent=1)
