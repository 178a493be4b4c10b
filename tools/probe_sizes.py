import sys; sys.path.insert(0, "."); sys.path.insert(0, "tests"); sys.path.insert(0, "tools")
from perf_probe import probe
for N in (512, 1024, 2048, 4096, 8192, 16384):
    probe("c1_bridge_ext", N, 400)
