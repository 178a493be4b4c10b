"""Smoke-sized runs of every kernel family for compute-sanitizer (tools/sanitize.sh): configs[1] in both lane layouts,
config 3 (surroundings observation), config 4 (the general step kernel) and a randoman game — constructor reset,
single steps (from the parked image), a masked reset, a fused rollout with auto-resets, an encode."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import os
import numpy as np
import torch
import parity_util as pu
from libzombsole_b200 import abi
from libzombsole_b200.engine import ZsEngine

CASES = [("c1_bridge_ext", 64, "32"), ("c1_bridge_ext", 64, "16"), ("c3_city_evac", 16, "32"), ("c4_maze_safehouse", 6, "32"),
         ("bots_randoman", 32, "32"), ("c5_bridge_channels", 32, "16")]
only = sys.argv[1:]
for name, N, lanes in CASES:
    if only and name not in only:
        continue
    os.environ["ZS_LANES_PER_ENV"] = lanes
    cfg, m = pu.build(pu.CONFIGS[name], N, 3, auto_reset=True, max_episode_steps=9)
    eng = ZsEngine(cfg, m)
    obs = eng.new_obs(2)
    rew, term, trunc = eng.new_outputs(6)
    acts = torch.zeros((6, N, eng.A), dtype=torch.int32, device=eng.device)
    for s in range(6):
        eng.fill_synthetic_actions(s, acts[s])
    for s in range(3):
        eng.step(acts[s], abi.ACTIONS_DISCRETE, obs[0], rew[s], term[s], trunc[s])
    eng.reset(torch.from_numpy((np.arange(N) % 3 == 0).astype(np.uint8)), obs[1])
    eng.rollout(6, 0, acts, abi.ACTIONS_DISCRETE, obs, rew, term, trunc)
    eng.rollout(5, 10, None, abi.ACTIONS_DISCRETE, obs, None, None, None)
    eng.encode_obs(obs[0])
    if eng.compact_words():
        rec = torch.zeros((N, eng.compact_words()), dtype=torch.int32, device=eng.device)
        eng.step_compact(acts[0], abi.ACTIONS_DISCRETE, rec, obs[0])
    torch.cuda.synchronize()
    print("ran", name, N, "lanes", eng.lanes_per_env(), flush=True)
    eng.close()
