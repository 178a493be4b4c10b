#!/bin/bash
# usage: tools/ncu_summary.sh <report.ncu-rep> [top lines]   -> key metrics, stall mix, per-line/per-function breakdown
rep=$1; top=${2:-40}
ncu -i $rep --page raw --csv 2>/dev/null > /tmp/_raw.csv
python - <<PY
import csv
rows=list(csv.reader(open('/tmp/_raw.csv')))
hdr,units,vals=rows[0],rows[1],rows[2]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__inst_executed.avg.per_cycle_elapsed','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__waves_per_multiprocessor','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__warps_eligible.avg.per_cycle_active','sm__cycles_elapsed.avg','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','lts__t_bytes.sum','launch__grid_size']
for i,h in enumerate(hdr):
    if h in want: print("%-70s %-12s %s"%(h,units[i],vals[i]))
PY
ncu -i $rep --page source --print-source cuda,sass --csv 2>/dev/null > /tmp/_src.csv
python - <<PY
import csv
from collections import defaultdict
rows=list(csv.reader(open('/tmp/_src.csv')))
hdr=None; tot=defaultdict(int)
for r in rows:
    if r and r[0]=="Line No": hdr=r; continue
    if hdr is None or not r or r[0]!="" or len(r)<len(hdr)-5: continue
    for i,h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h:
            try: tot[h]+=int(r[i])
            except: pass
s=sum(tot.values())
print("stalls: "+", ".join("%s %.1f%%"%(k[6:],100.0*v/s) for k,v in sorted(tot.items(), key=lambda kv:-kv[1])[:9]))
PY
python $(dirname $0)/ncu_lines.py /tmp/_src.csv $top
