import sys
sys.path.insert(0, "tools"); sys.path.insert(0, "."); sys.path.insert(0, "tests")
import probe_large as pl
pl.run("c4_maze_safehouse", 131072, 20)
