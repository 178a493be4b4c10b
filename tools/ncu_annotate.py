#!/usr/bin/env python
"""Annotated source listing from an ncu source page (csv, cuda,sass): instructions per warp-step and stall-sample share
per source line.  usage: ncu_annotate.py src.csv warp_steps file.cuh [first_line last_line]"""
import csv, sys, os
from collections import defaultdict
path, per, fname = sys.argv[1], float(sys.argv[2]), sys.argv[3]
lo = int(sys.argv[4]) if len(sys.argv) > 4 else 1
hi = int(sys.argv[5]) if len(sys.argv) > 5 else 10**9
rows = list(csv.reader(open(path)))
cur_file = cur_line = hdr = None
agg = defaultdict(lambda: [0.0, 0, 0])
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r; ii = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples"); ith = hdr.index("Thread Instructions Executed"); continue
    if hdr is None or len(r) < len(hdr) - 5: continue
    if r[0] != "": cur_line = int(r[0]); continue
    try:
        a = agg[(cur_file, cur_line)]; a[0] += int(r[ii]) / per; a[1] += int(r[isamp]); a[2] += int(r[ith])
    except Exception: pass
tot_s = sum(a[1] for a in agg.values()); tot_i = sum(a[0] for a in agg.values())
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(root, "libzombsole_b200", "csrc", fname)).read().split("\n")
si = ss = 0
for n, line in enumerate(src, 1):
    if n < lo or n > hi: continue
    a = agg.get((fname, n))
    if a:
        si += a[0]; ss += a[1]
        print("%5d %7.1f %5.2f%% %5.1f | %s" % (n, a[0], 100.0 * a[1] / tot_s, a[2] / max(1.0, a[0] * per), line[:120]))
    else:
        print("%5d %7s %6s %5s | %s" % (n, "", "", "", line[:120]))
print("range: %.1f inst/warp-step, %.1f%% of samples (kernel total %.1f inst/warp-step)" % (si, 100.0 * ss / tot_s, tot_i))
