import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import parity_util as pu
from libzombsole_b200 import abi
from test_cuda_properties import engine
from oracle import oracle as orc
name, N = "c5_bridge_channels", 2500
K = 6
bad_total = 0
for trial in range(40):
    eng, cfg, m = engine(name, N, seed=17 + trial, base=3)
    ref = orc.OracleEnv(cfg, m)
    obs = eng.new_obs()
    for rep in range(4):
        eng.rollout(K, 100 * rep, None, abi.ACTIONS_DISCRETE, obs, None, None, None)
        o, _, _, _ = ref.rollout_synthetic(K, 100 * rep)
        got = obs.cpu().numpy().reshape(N, -1)
        if not np.array_equal(got, o):
            bad = np.argwhere(got != o)
            bad_total += len(bad)
            cells = 1332
            for (e, idx) in bad[:5]:
                c = idx % cells
                print("trial %d rep %d env %d plane %d cell %d (x=%d,y=%d): got %s want %s | planes got %s want %s" % (
                    trial, rep, e, idx // cells, c, c % 111, c // 111, got[e, idx], o[e, idx],
                    [got[e, p * cells + c] for p in range(3)], [o[e, p * cells + c] for p in range(3)]), flush=True)
    eng.close(); ref.close()
print("total differing elements:", bad_total)
