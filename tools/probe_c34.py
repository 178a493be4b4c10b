import sys
sys.path.insert(0, "/root/repo/tools"); sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import probe_large as pl
pl.run("c4_maze_safehouse", 131072, 20)
pl.run("c3_city_evac", 65536, 100)
pl.run("c5_bridge_channels", 1 << 20, 50)
