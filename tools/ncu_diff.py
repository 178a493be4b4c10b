#!/usr/bin/env python
"""Compare two ncu source pages (csv, cuda,sass) line by line: instructions per warp-step.
usage: ncu_diff.py a.csv warp_steps_a b.csv warp_steps_b [top]"""
import csv, sys
from collections import defaultdict

def load(path, per):
    rows = list(csv.reader(open(path)))
    cur_file = cur_line = hdr = None
    cur_src = ""
    agg = defaultdict(lambda: [0.0, 0.0, ""])
    for r in rows:
        if not r: continue
        if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
        if r[0] == "Line No":
            hdr = r; ii = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples"); continue
        if hdr is None or len(r) < len(hdr) - 5: continue
        if r[0] != "": cur_line = int(r[0]); cur_src = r[1].strip(); continue
        try:
            a = agg[(cur_file, cur_line)]; a[0] += int(r[ii]) / per; a[1] += int(r[isamp]); a[2] = cur_src
        except Exception: pass
    return agg

A = load(sys.argv[1], float(sys.argv[2])); B = load(sys.argv[3], float(sys.argv[4]))
top = int(sys.argv[5]) if len(sys.argv) > 5 else 60
ta = sum(v[0] for v in A.values()); tb = sum(v[0] for v in B.values())
sa = sum(v[1] for v in A.values()); sb = sum(v[1] for v in B.values())
print("total inst/warp-step: A %.1f  B %.1f" % (ta, tb))
keys = set(A) | set(B)
rows = []
for k in keys:
    a = A.get(k, [0, 0, ""]); b = B.get(k, [0, 0, ""])
    rows.append((a[0] - b[0], k, a, b))
rows.sort(key=lambda r: -abs(r[0]))
print("%-14s %5s %8s %8s %7s %7s  %s" % ("file", "line", "A", "B", "A samp%", "B samp%", "source"))
for d, k, a, b in rows[:top]:
    print("%-14s %5d %8.1f %8.1f %6.2f%% %6.2f%%  %s" % (k[0][:14], k[1], a[0], b[0], 100 * a[1] / sa, 100 * b[1] / sb, (a[2] or b[2])[:100]))
