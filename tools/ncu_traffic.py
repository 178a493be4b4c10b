#!/usr/bin/env python
"""profiles/r01_traffic.json from an ncu --set full capture of the timed zs_rollout launch of
`bench.py --steps K --warmup W` (N envs): DRAM bytes and instructions per env-step.
usage: ncu_traffic.py report.ncu-rep N K summary_path"""
import csv, json, subprocess, sys
rep, N, K, src = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
def get(name):
    i = hdr.index(name)
    v = float(vals[i].replace(",", ""))
    u = units[i]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0}.get(u, 1.0)
steps = N * K
out = {"kernel": vals[hdr.index("Kernel Name")], "capture": src, "env_steps_in_launch": steps,
       "dram_bytes_read": get("dram__bytes_read.sum"), "dram_bytes_write": get("dram__bytes_write.sum"),
       "dram_bytes_per_env_step": (get("dram__bytes_read.sum") + get("dram__bytes_write.sum")) / steps,
       "inst_per_env_step": get("smsp__inst_executed.sum") / steps, "duration_us": get("gpu__time_duration.sum")}
print(json.dumps(out, indent=1))
