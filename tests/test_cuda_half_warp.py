"""Two envs per warp (half-warp lane groups, ZS_LANES_PER_ENV=16): the mode large batches of small worlds
run in.  Both envs of a warp share every warp primitive of the step loop, while the rare paths (world
init, minimum-zombie respawn, sequential execute) may be taken by one of them alone — the cases below are
chosen so that neighbouring envs diverge in every one of those ways.  Same bar: bit-exact vs the oracle."""
import numpy as np
import pytest

import parity_util as pu
from libzombsole_b200 import abi
from test_cuda_parity import lockstep
from test_cuda_properties import engine, state_snapshot, assert_same_state

pytestmark = pytest.mark.gpu

HALF_WARP_CASES = [
    ("c1_bridge_ext", 64, 120, 0), ("c1_bridge_ext", 30, 60, 13), ("c5_bridge_channels", 32, 60, 0),
    ("gym_v0_alone", 64, 80, 0), ("gym_surroundings", 32, 60, 0), ("surroundings_channels", 32, 80, 0),
    ("safehouse_small", 32, 120, 0), ("multi_boxed_2p", 64, 80, 0), ("survival_minz", 32, 120, 25),
    ("bots_hamsters", 32, 120, 0), ("minz_allcells", 16, 80, 0), ("no_zombies", 16, 30, 0),
    ("bots_randoman", 64, 150, 0),
]


@pytest.fixture(autouse=True)
def half_warp(monkeypatch):
    monkeypatch.setenv("ZS_LANES_PER_ENV", "16")


def lanes_of(eng):
    return eng.lanes_per_env()


@pytest.mark.parametrize("name,N,T,mes", HALF_WARP_CASES)
def test_half_warp_matches_oracle_lockstep(name, N, T, mes):
    errs = lockstep(pu.CONFIGS[name], N, T, seed=77 + N + T, base=3, mes=mes)
    assert not errs, "\n".join(errs[:3])


def test_half_warp_mode_is_what_ran():
    eng, cfg, m = engine("c1_bridge_ext", 64)
    assert lanes_of(eng) == 16
    eng.close()
    eng, cfg, m = engine("c1_bridge_ext", 63)  # odd batch: one env per warp
    assert lanes_of(eng) == 32
    eng.close()
    eng, cfg, m = engine("c3_city_evac", 64)   # 24 slots do not fit 16 lanes
    assert lanes_of(eng) == 32
    eng.close()


@pytest.mark.parametrize("name,N,K", [("c1_bridge_ext", 4096, 64), ("survival_minz", 2048, 48), ("bots_hamsters", 1024, 48)])
def test_half_warp_fused_rollout_matches_oracle(name, N, K):
    from oracle import oracle as orc
    eng, cfg, m = engine(name, N, seed=11)
    assert lanes_of(eng) == 16
    obs = eng.new_obs()
    rew, term, trunc = eng.new_outputs(K)
    eng.rollout(K, 0, None, abi.ACTIONS_DISCRETE, obs, rew, term, trunc)
    ref = orc.OracleEnv(cfg, m)
    o, r, te, tr = ref.rollout_synthetic(K, 0)
    assert np.array_equal(obs.cpu().numpy().reshape(N, -1), o)
    assert np.array_equal(rew.cpu().numpy().view(np.uint64).reshape(K, N, -1), r.view(np.uint64).reshape(K, N, -1))
    assert np.array_equal(term.cpu().numpy(), te) and np.array_equal(trunc.cpu().numpy(), tr)
    assert np.array_equal(eng.episode_stats().cpu().numpy(), ref.stats())
    eng.close()


def test_half_warp_equals_full_warp_state(monkeypatch):
    """Same trajectories whichever way the lanes are split: final state of a fused rollout."""
    N, K = 2048, 96
    snaps = []
    for lanes in ("16", "32"):
        monkeypatch.setenv("ZS_LANES_PER_ENV", lanes)
        eng, cfg, m = engine("c1_bridge_ext", N, seed=5)
        assert lanes_of(eng) == int(lanes)
        eng.rollout(K, 0, None, abi.ACTIONS_DISCRETE, None, None, None, None)
        snaps.append(state_snapshot(eng))
        M = eng.M
        eng.close()
    assert_same_state(snaps[0], snaps[1], M)
