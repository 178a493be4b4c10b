"""Scratch performance probe (not the bench): per-step launches vs fused rollout at several N."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import parity_util as pu
from libzombsole_b200 import abi
from libzombsole_b200.engine import ZsEngine

def probe(name, N, K=200, slots=8):
    cfg, m = pu.build(pu.CONFIGS[name], N, 0, auto_reset=True, max_episode_steps=1000)
    eng = ZsEngine(cfg, m)
    obs = eng.new_obs(slots)
    rew, term, trunc = eng.new_outputs(K)
    acts = torch.zeros((K, N, eng.A), dtype=torch.int32, device=eng.device)
    for s in range(K):
        eng.fill_synthetic_actions(s, acts[s])
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    # per-step launches
    for s in range(20):
        eng.step(acts[s], abi.ACTIONS_DISCRETE, obs[s % slots], rew[s], term[s], trunc[s])
    torch.cuda.synchronize()
    ev[0].record()
    for s in range(K):
        eng.step(acts[s], abi.ACTIONS_DISCRETE, obs[s % slots], rew[s], term[s], trunc[s])
    ev[1].record()
    torch.cuda.synchronize()
    t_step = ev[0].elapsed_time(ev[1]) / K
    # fused rollout, actions from tensor
    eng.rollout(K, 0, acts, abi.ACTIONS_DISCRETE, obs, rew, term, trunc)
    torch.cuda.synchronize()
    ev[2].record()
    eng.rollout(K, 0, acts, abi.ACTIONS_DISCRETE, obs, rew, term, trunc)
    ev[3].record()
    torch.cuda.synchronize()
    t_roll = ev[2].elapsed_time(ev[3]) / K
    ev[2].record()
    eng.rollout(K, 0, None, abi.ACTIONS_DISCRETE, obs, rew, term, trunc)
    ev[3].record()
    torch.cuda.synchronize()
    t_syn = ev[2].elapsed_time(ev[3]) / K
    st = eng.episode_stats().cpu().tolist()
    print("%-20s N=%8d  per-step %8.1f us (%.3e/s)  rollout %8.1f us (%.3e/s)  synthetic %8.1f us (%.3e/s) stats %s" % (
        name, N, t_step * 1e3, N / t_step * 1e3, t_roll * 1e3, N / t_roll * 1e3, t_syn * 1e3, N / t_syn * 1e3, st), flush=True)
    eng.close()

if __name__ == "__main__":
    for name, N, K in [("c1_bridge_ext", 4096, 400), ("c1_bridge_ext", 65536, 100), ("c1_bridge_ext", 1 << 20, 20),
                       ("c5_bridge_channels", 65536, 50), ("c3_city_evac", 65536, 30), ("c4_maze_safehouse", 16384, 20)]:
        probe(name, N, K)
