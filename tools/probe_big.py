import sys; sys.path.insert(0, "."); sys.path.insert(0, "tests"); sys.path.insert(0, "tools")
from perf_probe import probe
for N, K in ((65536, 100), (1 << 20, 20)):
    probe("c1_bridge_ext", N, K)
