"""Aggregate an ncu source page (ncu -i X.ncu-rep --page source --csv --print-source cuda,sass) by source FUNCTION:
executed warp instructions, stall samples, average active threads.    python tools/ncu_by_function.py page.csv [env_steps]"""
import bisect
import csv
import os
import re
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "libzombsole_b200", "csrc")


def function_starts(path):
    out = []
    with open(path) as f:
        for i, line in enumerate(f, 1):
            m = re.search(r"(?:__device__|__global__)[^;(]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", line)
            if m and not line.strip().startswith("//"):
                out.append((i, m.group(1)))
    return out


starts = {fn: function_starts(os.path.join(CSRC, fn)) for fn in os.listdir(CSRC) if fn.endswith((".cu", ".cuh"))}
rows = list(csv.reader(open(sys.argv[1])))
steps = float(sys.argv[2]) if len(sys.argv) > 2 else None
cur_file = cur_line = None
hdr = None
agg = defaultdict(lambda: [0, 0, 0])
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        i_inst, i_samp, i_thr = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
        continue
    if hdr is None or len(r) < len(hdr) - 5:
        continue
    if r[0] != "":
        cur_line = int(r[0])
        continue
    try:
        inst, samp, thr = int(r[i_inst]), int(r[i_samp]), int(r[i_thr])
    except (ValueError, IndexError):
        continue
    st = starts.get(cur_file)
    name = cur_file
    if st:
        k = bisect.bisect_right([s[0] for s in st], cur_line) - 1
        name = st[k][1] if k >= 0 else cur_file
    a = agg[(cur_file, name)]
    a[0] += inst; a[1] += samp; a[2] += thr
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print("total warp instructions %d, samples %d%s" % (ti, ts, "" if not steps else ", %.0f warp instructions per env-step" % (ti / steps)))
for (f, n), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print("  %-16s %-26s inst %5.1f%%%s  samples %5.1f%%  threads %4.1f" % (
        f, n, 100.0 * a[0] / ti, "" if not steps else " (%5.0f/step)" % (a[0] / steps), 100.0 * a[1] / ts, a[2] / max(1, a[0])))
