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
rew, term, trunc = eng.new_outputs(K)
acts = torch.zeros((2 * K, N, eng.A), dtype=torch.int32, device=eng.device)
for s in range(2 * K):
    eng.fill_synthetic_actions(s, acts[s])
eng.rollout(K, 0, acts[:K], abi.ACTIONS_DISCRETE, obs, rew, term, trunc)   # warm-up launch
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
torch.cuda.synchronize()
ev[0].record()
eng.rollout(K, K, acts[K:], abi.ACTIONS_DISCRETE, obs, rew, term, trunc)   # the profiled launch
ev[1].record()
torch.cuda.synchronize()
print("%s N=%d K=%d: %.3e env-steps/s" % (name, N, K, N * K / ev[0].elapsed_time(ev[1]) * 1e3))
eng.close()
