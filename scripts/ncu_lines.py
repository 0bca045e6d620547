#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line: share of executed warp instructions
and stall samples per line. usage: ncu_lines.py export.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if len(r) > 8 and r[0] == "Line No" and "Instructions Executed" in r:
        hdr = r
        il, isrc, ie, ism, ith = 0, 1, r.index("Instructions Executed"), r.index("# Samples"), r.index("Thread Instructions Executed")
        continue
    if hdr is None or len(r) <= ie:
        continue
    try:
        v, sm, th = float(r[ie] or 0), float(r[ism] or 0), float(r[ith] or 0)
    except ValueError:
        continue
    a = agg.setdefault(r[il], [0.0, 0.0, 0.0, r[isrc]])
    a[0] += v; a[1] += sm; a[2] += th
tot = sum(a[0] for a in agg.values()) or 1; tots = sum(a[1] for a in agg.values()) or 1
print("total warp instructions %.3e, samples %d" % (tot, tots))
for ln, a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print("%5.1f%% inst %5.1f%% stall  thr/inst %4.1f  L%-5s %s" % (100 * a[0] / tot, 100 * a[1] / tots, a[2] / a[0] if a[0] else 0, ln, a[3].strip()[:150]))
