#!/usr/bin/env python
"""Phase-level instruction breakdown of world_step from /tmp/_src.csv (ncu source page, cuda,sass)."""
import csv, sys, os
from collections import defaultdict
per = float(sys.argv[1]) if len(sys.argv) > 1 else 204800.0
rows = list(csv.reader(open('/tmp/_src.csv')))
cur_file = cur_line = hdr = None
agg = defaultdict(lambda: [0, 0, 0])
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r; ii = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples"); ith = hdr.index("Thread Instructions Executed"); continue
    if hdr is None or len(r) < len(hdr) - 5: continue
    if r[0] != "": cur_line = int(r[0]); continue
    try:
        a = agg[(cur_file, cur_line)]; a[0] += int(r[ii]); a[1] += int(r[isamp]); a[2] += int(r[ith])
    except Exception: pass
tots = sum(a[1] for a in agg.values())
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(root, 'libzombsole_b200/csrc/zs_world.cuh')).read().split('\n')
marks = [("tables", "per-step tables"), ("ranks", "dict-order rank of every thing"), ("distances", "closest(self, others) (utils.py:23-31) for everybody"),
         ("decide", "get_actions (core.py:80-101)"), ("wander", "int nd = 0;"), ("action list", "actions list in actor (dict) order"),
         ("draws+FY partners", "draws of this step, generated"), ("shuffle+execute", "random.shuffle + execute_actions"),
         ("broadcast+clean", "k = __shfl_sync(ZS_FULL, k, 0);"), ("end", "World.spawn_in_random (core.py:40-66)")]
idx = []
for name, m in marks:
    for n, l in enumerate(src, 1):
        if m in l: idx.append((n, name)); break
tot_all = 0
for (s, name), (e, _) in zip(idx, idx[1:]):
    sel = [a for (f, l), a in agg.items() if f == "zs_world.cuh" and s <= l < e]
    i = sum(a[0] for a in sel); sm = sum(a[1] for a in sel); t = sum(a[2] for a in sel)
    print("%-20s %7.1f inst/step  samples %5.1f%%  thr %.1f" % (name, i / per, 100.0 * sm / tots, t / max(1, i)))
byf = defaultdict(lambda: [0, 0])
for (f, l), a in agg.items(): byf[f][0] += a[0]; byf[f][1] += a[1]
for f, (i, sm) in sorted(byf.items(), key=lambda kv: -kv[1][0]):
    print("%-28s %7.1f inst/step  samples %5.1f%%" % (f, i / per, 100.0 * sm / tots))
print("total %.1f inst/step" % (sum(v[0] for v in byf.values()) / per))
