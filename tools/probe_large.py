"""Large-batch runs of BASELINE configs 3, 4 and 5 on one GPU (scratch tool, not the bench):
config 5 = 8,388,608 bridge envs with the full-map observation (simple and channels),
config 4's per-GPU share = 131,072 maze/safehouse envs with 100 zombies,
config 3 = 65,536 four-agent evacuation envs."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import parity_util as pu
from libzombsole_b200 import abi
from libzombsole_b200.engine import ZsEngine

def run(name, N, K, slots=1):
    cfg, m = pu.build(pu.CONFIGS[name], N, 0, auto_reset=True, max_episode_steps=1000)
    t0 = time.time()
    eng = ZsEngine(cfg, m)
    obs = eng.new_obs(slots) if slots > 1 else eng.new_obs()
    torch.cuda.synchronize()
    t_init = time.time() - t0
    eng.rollout(2, 0, None, abi.ACTIONS_DISCRETE, obs, None, None, None)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    eng.rollout(K, 2, None, abi.ACTIONS_DISCRETE, obs, None, None, None)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1])
    gb = eng.state.numel() / 1e9 + obs.numel() * 4 / 1e9
    print("%-20s N=%9d K=%3d  %8.2f ms/step  %.3e env-steps/s  (state+obs %.1f GB, init %.1fs, lanes/env %s) stats %s" % (
        name, N, K, ms / K, N * K / ms * 1e3, gb, t_init, "auto", eng.episode_stats().cpu().tolist()), flush=True)
    eng.close()
    del obs, eng
    torch.cuda.empty_cache()

if __name__ == "__main__":
    run("c1_bridge_ext", 8388608, 20)
    run("c5_bridge_channels", 8388608 // 2, 20)
    run("c4_maze_safehouse", 131072, 20)
    run("c3_city_evac", 65536, 40)
