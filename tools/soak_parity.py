"""Parity soak (not part of the test suite: longer and wider than it): for every parity configuration and a few seeds, a
fused rollout with same-step auto-resets, the same steps as single launches, and the oracle — observations of the last
step, every reward's float64 bits, every flag, the episode statistics must agree exactly.

    python tools/soak_parity.py [steps] [seeds]
"""
import sys
import time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import torch
import parity_util as pu
from libzombsole_b200 import abi
from libzombsole_b200.engine import ZsEngine
from oracle import oracle as orc


def one(name, N, K, seed):
    cfg, m = pu.build(pu.CONFIGS[name], N, seed, auto_reset=True, max_episode_steps=60)
    ref = orc.OracleEnv(cfg, m)
    o, r, te, tr = ref.rollout_synthetic(K, 0)
    # fused
    eng = ZsEngine(cfg, m)
    obs = eng.new_obs()
    rew, term, trunc = eng.new_outputs(K)
    eng.rollout(K, 0, None, abi.ACTIONS_DISCRETE, obs, rew, term, trunc)
    ok = (np.array_equal(obs.cpu().numpy().reshape(N, -1), o)
          and np.array_equal(rew.cpu().numpy().view(np.uint64).reshape(K, N, -1), r.view(np.uint64).reshape(K, N, -1))
          and np.array_equal(term.cpu().numpy(), te) and np.array_equal(trunc.cpu().numpy(), tr)
          and np.array_equal(eng.episode_stats().cpu().numpy(), ref.stats()))
    eng.close()
    # the same steps in launches of 1, 3 and 8 steps (the short-launch shape, the parked images in between)
    eng = ZsEngine(cfg, m)
    obs = eng.new_obs()
    rew, term, trunc = eng.new_outputs(K)
    s = 0
    for k in [1, 3, 8] * K:
        k = min(k, K - s)
        if k <= 0:
            break
        eng.rollout(k, s, None, abi.ACTIONS_DISCRETE, obs, rew[s:s + k], term[s:s + k], trunc[s:s + k])
        s += k
    ok2 = (np.array_equal(obs.cpu().numpy().reshape(N, -1), o)
           and np.array_equal(rew.cpu().numpy().view(np.uint64).reshape(K, N, -1), r.view(np.uint64).reshape(K, N, -1))
           and np.array_equal(term.cpu().numpy(), te) and np.array_equal(trunc.cpu().numpy(), tr)
           and np.array_equal(eng.episode_stats().cpu().numpy(), ref.stats()))
    eng.close()
    return ok, ok2


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    seeds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    bad = 0
    t0 = time.time()
    for name in pu.CONFIGS:
        big = pu.CONFIGS[name].get("map_name") in ("maze_for_safehouse", "city_for_evacuation", "fort")
        N = 96 if big else 384
        for seed in range(1, seeds + 1):
            ok, ok2 = one(name, N, K if not big else max(20, K // 3), 100 * seed + 7)
            if not (ok and ok2):
                bad += 1
            print("%-24s N=%4d seed=%4d fused %s, in short launches %s   (%.0f s)" % (
                name, N, 100 * seed + 7, "ok" if ok else "DIFFERS", "ok" if ok2 else "DIFFERS", time.time() - t0), flush=True)
    print("soak: %d mismatching runs" % bad)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
