#!/usr/bin/env python
"""SASS listing of the hot loop with per-instruction samples: ncu --page source --print-source sass --csv.
usage: ncu_sass.py sass.csv warp_steps [min_exec_frac]  -> instructions executed at least min_exec_frac times per warp-step"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
per = float(sys.argv[2]); thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
hdr = rows[1]
ia, isrc, isamp, iinst = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot = sum(int(r[isamp]) for r in rows[2:] if len(r) > iinst)
base = int(rows[2][ia], 16)
acc = 0
for r in rows[2:]:
    if len(r) <= iinst: continue
    ex = int(r[iinst]) / per; sm = int(r[isamp])
    if ex < thr: continue
    acc += sm
    print("%05x %6.2f %5.2f%% %6.1f%%  %s" % (int(r[ia], 16) - base, ex, 100.0 * sm / tot, 100.0 * acc / tot, r[isrc].strip()[:90]))
