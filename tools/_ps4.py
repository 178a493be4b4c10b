import sys
sys.path.insert(0, "tools"); sys.path.insert(0, "."); sys.path.insert(0, "tests")
from perf_probe import probe
probe("c1_bridge_ext", 4096, 200)
