"""Scratch probe: fused rollout throughput at several batch sizes, lanes-per-env forced by ZS_LANES_PER_ENV."""
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import parity_util as pu
from libzombsole_b200 import abi
from libzombsole_b200.engine import ZsEngine

def run(name, N, K, slots):
    cfg, m = pu.build(pu.CONFIGS[name], N, 0, auto_reset=True, max_episode_steps=1000)
    eng = ZsEngine(cfg, m)
    obs = eng.new_obs(slots)
    rew, term, trunc = eng.new_outputs(K)
    acts = torch.zeros((2 * K, N, eng.A), dtype=torch.int32, device=eng.device)
    for s in range(2 * K):
        eng.fill_synthetic_actions(s, acts[s])
    eng.rollout(K, 0, acts[:K], abi.ACTIONS_DISCRETE, obs, rew, term, trunc)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    eng.rollout(K, K, acts[K:], abi.ACTIONS_DISCRETE, obs, rew, term, trunc)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1])
    print("%s lanes=%s N=%d K=%d: %.2f us/step %.3e env-steps/s" % (name, os.environ.get("ZS_LANES_PER_ENV", "auto"), N, K, ms / K * 1e3, N * K / ms * 1e3), flush=True)
    eng.close()

if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "c1_bridge_ext"
    for N in [int(a) for a in sys.argv[2:]] or [4096, 8192, 16384, 65536]:
        run(name, N, 400 if N <= 16384 else 100, max(2, min(16, (300 << 20) // (N * 5328))))
