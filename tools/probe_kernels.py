"""Scratch probe: the reset and encode kernels against the HBM roofline (CUDA events, all envs, large batch).
Algorithmic bytes per env (SURVEY §8d's canonical layout): state read R = 12*M + 2*S + cells/4 + 24, state write
Wr = 12*M + 24, observation O.  reset = R + Wr + O (slots keep their last position until re-placed, so the state is
read), encode = R + O."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import parity_util as pu
from libzombsole_b200 import abi
from libzombsole_b200.engine import ZsEngine

PEAK = 6538.3e9

def run(name, N, reps=10):
    cfg, m = pu.build(pu.CONFIGS[name], N, 0, auto_reset=True, max_episode_steps=1000)
    eng = ZsEngine(cfg, m)
    slots = max(2, min(16, (400 << 20) // (N * eng.obs_elems * 4)))
    obs = eng.new_obs(slots)
    eng.rollout(30, 0, None, abi.ACTIONS_DISCRETE, obs[0], None, None, None)  # worlds in mid-episode, some damage
    M, S, cells, O = eng.M, eng.S, eng.cells, eng.obs_elems * 4
    R, Wr = 12 * M + 2 * S + cells // 4 + 24, 12 * M + 24
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    out = {}
    for what, fn, b in (("encode", lambda i: eng.encode_obs(obs[i % slots]), R + O),
                        ("reset", lambda i: eng.reset(None, obs[i % slots]), R + Wr + O)):
        for i in range(3): fn(i)
        torch.cuda.synchronize()
        ev[0].record()
        for i in range(reps): fn(i)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / reps
        out[what] = (ms, N / ms * 1e3, N * b / ms * 1e3 / PEAK)
        print("%-20s N=%8d %-7s %8.3f ms  %.3e envs/s  B_alg %6d  %.2f of the HBM roofline" % (name, N, what, ms, N / ms * 1e3, b, out[what][2]), flush=True)
    eng.close()

if __name__ == "__main__":
    if len(sys.argv) > 2:
        run(sys.argv[1], int(sys.argv[2]))
        sys.exit(0)
    for name, N in (("c1_bridge_ext", 65536), ("c1_bridge_ext", 1 << 20), ("c5_bridge_channels", 1 << 19), ("c3_city_evac", 65536), ("c4_maze_safehouse", 131072)):
        run(name, N)
