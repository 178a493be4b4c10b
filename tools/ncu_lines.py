#!/usr/bin/env python
"""Aggregate an ncu source page (--print-source cuda,sass --csv) by CUDA source line:
instructions executed, stall samples and average active threads per line."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
rows = list(csv.reader(open(path)))
cur_file, cur_line, cur_src = None, None, ""
agg = defaultdict(lambda: [0, 0, 0, ""])  # inst, samples, thread_inst, src
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        i_inst = hdr.index("Instructions Executed")
        i_samp = hdr.index("# Samples")
        i_thr = hdr.index("Thread Instructions Executed")
        continue
    if hdr is None or len(r) < len(hdr) - 5:
        continue
    if r[0] != "":
        cur_line, cur_src = r[0], r[1].strip()
        continue
    try:
        inst, samp, thr = int(r[i_inst]), int(r[i_samp]), int(r[i_thr])
    except (ValueError, IndexError):
        continue
    a = agg[(cur_file, int(cur_line))]
    a[0] += inst; a[1] += samp; a[2] += thr; a[3] = cur_src
tot_i = sum(a[0] for a in agg.values()); tot_s = sum(a[1] for a in agg.values())
print("total inst %d samples %d" % (tot_i, tot_s))
byfile = defaultdict(lambda: [0, 0])
for (f, l), a in agg.items():
    byfile[f][0] += a[0]; byfile[f][1] += a[1]
for f, (i, s) in byfile.items():
    print("  %-16s inst %5.1f%%  samples %5.1f%%" % (f, 100.0 * i / tot_i, 100.0 * s / tot_s))
print("%-16s %5s %7s %7s %6s  %s" % ("file", "line", "inst%", "samp%", "thr", "source"))
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-16s %5d %6.2f%% %6.2f%% %6.1f  %s" % (f, l, 100.0 * a[0] / tot_i, 100.0 * a[1] / tot_s,
                                                   a[2] / max(1, a[0]), a[3][:110]))

# phase breakdown for zs_world.cuh by function (line ranges read from the file itself)
import os, re
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for fname in ("zs_world.cuh", "zs_obs.cuh", "zs_b200.cu", "zs_device.cuh"):
    fpath = os.path.join(root, "libzombsole_b200", "csrc", fname)
    if not os.path.exists(fpath):
        continue
    starts = []
    for n, line in enumerate(open(fpath), 1):
        m = re.match(r"^(?:template.*)?(?:__device__|__global__|static|extern).*?([A-Za-z_0-9]+)\s*\(", line)
        if m and not line.startswith(" "):
            starts.append((n, m.group(1)))
    if not starts:
        continue
    starts.append((10 ** 9, "end"))
    tot = defaultdict(lambda: [0, 0, 0])
    for (f, l), a in agg.items():
        if f != fname:
            continue
        name = "?"
        for (s, nm), (s2, _) in zip(starts, starts[1:]):
            if s <= l < s2:
                name = nm
                break
        tot[name][0] += a[0]; tot[name][1] += a[1]; tot[name][2] += a[2]
    print("--", fname)
    for nm, (i, s, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print("   %-28s inst %5.1f%%  samples %5.1f%%  avg threads %4.1f" % (nm, 100.0 * i / tot_i, 100.0 * s / tot_s, t / max(1, i)))
