"""One fused rollout of a named parity config: python tools/probe_one.py <config> <N> <K>  (profiling target)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import parity_util as pu
from libzombsole_b200 import abi
from libzombsole_b200.engine import ZsEngine

name, N, K = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
cfg, m = pu.build(pu.CONFIGS[name], N, 0, auto_reset=True, max_episode_steps=1000)
eng = ZsEngine(cfg, m)
obs = eng.new_obs(4)
eng.rollout(K, 0, None, abi.ACTIONS_DISCRETE, obs, None, None, None)   # warm-up launch
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
torch.cuda.synchronize()
ev[0].record()
eng.rollout(K, K, None, abi.ACTIONS_DISCRETE, obs, None, None, None)   # the profiled launch
ev[1].record()
torch.cuda.synchronize()
print("%s N=%d K=%d: %.3e env-steps/s" % (name, N, K, N * K / ev[0].elapsed_time(ev[1]) * 1e3))
eng.close()
