import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import parity_util as pu
from libzombsole_b200 import abi
from libzombsole_b200.engine import ZsEngine
N=4096
cfg, m = pu.build(pu.CONFIGS["c1_bridge_ext"], N, 0, auto_reset=True, max_episode_steps=1000)
eng = ZsEngine(cfg, m)
obs = eng.new_obs(13)
K=512
rew, term, trunc = eng.new_outputs(K)
acts = torch.zeros((K, N, 1), dtype=torch.int32, device=eng.device)
eng.fill_synthetic_tape(0, acts)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for name, o in (("with obs (FAST)", obs), ("obs=None (non-FAST kernel, no observation)", None)):
    for _ in range(2): eng.rollout(K, 0, acts, abi.ACTIONS_DISCRETE, o, rew, term, trunc)
    torch.cuda.synchronize(); ev[0].record()
    for _ in range(10): eng.rollout(K, 0, acts, abi.ACTIONS_DISCRETE, o, rew, term, trunc)
    ev[1].record(); torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / 10
    print("%-45s K=%d: %.2f us/step  %.3e env-steps/s" % (name, K, ms * 1e3 / K, N * K / ms * 1e3))
