"""The parked on-chip image (csrc/zs_device.cuh: EnvS, "the IMAGE"): a launch leaves ranks, occupancy grid, patch
and dead-body lists of every env in device memory and the next launch starts from them with one bulk copy instead of
re-deriving them from the state buffer.  Nothing observable may depend on which way a launch started:

* a mixed sequence of launches (masked resets, single steps, short and long fused rollouts, encodes) with the images
  and without them (ZS_IMAGE_MB=0) gives identical outputs and an identical canonical state after every launch;
* a write into the state buffer followed by state_written() is honoured (the reference's tests edit lives between
  steps, tests/test_game.py:41-66), checked against the oracle;
* every step of a fused rollout writes its observation to its ring slot (one TMA bulk copy of the pristine planes, then
  the patches): each slot equals the oracle's observation after that step, with one slot and with K slots, in both
  lane layouts, paired TMA copy on and off.
"""
import numpy as np
import pytest
import torch

import parity_util as pu
from libzombsole_b200 import abi
from test_cuda_properties import assert_same_state, engine, state_snapshot

pytestmark = pytest.mark.gpu


def _run_sequence(name, N, seed):
    """A fixed mixed sequence of launches; returns what every launch produced plus the state after it."""
    eng, cfg, m = engine(name, N, seed=seed, base=3)
    A = cfg.n_agents
    K = 40
    tape = torch.zeros((K, N, A), dtype=torch.int32, device=eng.device)
    for s in range(K):
        eng.fill_synthetic_actions(50 + s, tape[s])
    obs, ring = eng.new_obs(), eng.new_obs(3)
    rew, term, trunc = eng.new_outputs(K)
    log = []

    def snap(tag, *tensors):
        torch.cuda.synchronize()
        log.append((tag, [t.clone() for t in tensors], state_snapshot(eng)))

    rs = np.random.RandomState(seed)
    t = 0
    for rnd in range(3):
        for _ in range(3):  # single steps
            eng.step(tape[t], abi.ACTIONS_DISCRETE, obs, rew[t], term[t], trunc[t])
            snap("step%d" % t, obs, rew[t], term[t], trunc[t])
            t += 1
        mask = torch.from_numpy((rs.rand(N) < 0.3).astype(np.uint8))
        eng.reset(mask, obs)  # masked reset: the other envs' images stay as they are
        snap("reset%d" % rnd, obs[mask.bool().to(obs.device)])
        eng.rollout(3, 0, tape[t:t + 3], abi.ACTIONS_DISCRETE, ring, rew[t:t + 3], term[t:t + 3], trunc[t:t + 3])
        snap("roll3_%d" % rnd, ring, rew[t:t + 3], term[t:t + 3])
        t += 3
        eng.encode_obs(obs)
        snap("encode%d" % rnd, obs)
        eng.rollout(6, 0, tape[t:t + 6], abi.ACTIONS_DISCRETE, obs, rew[t:t + 6], term[t:t + 6], trunc[t:t + 6])
        snap("roll6_%d" % rnd, obs, rew[t:t + 6], term[t:t + 6], trunc[t:t + 6])
        t += 6
    eng.reset(None, obs)
    snap("reset_all", obs)
    eng.step(tape[t], abi.ACTIONS_DISCRETE, obs, rew[t], term[t], trunc[t])
    snap("last", obs, rew[t])
    M = eng.M
    eng.close()
    return log, M


@pytest.mark.parametrize("name,N", [("c1_bridge_ext", 96), ("c1_bridge_ext", 4096), ("c5_bridge_channels", 2500),
                                    ("c3_city_evac", 64), ("c4_maze_safehouse", 24), ("bots_randoman", 128),
                                    ("survival_minz", 96), ("box_arena", 32)])
def test_launches_from_image_equal_launches_from_state(monkeypatch, name, N):
    with_img, M = _run_sequence(name, N, 17)
    monkeypatch.setenv("ZS_IMAGE_MB", "0")
    without, _ = _run_sequence(name, N, 17)
    assert len(with_img) == len(without)
    for (tag, ta, sa), (_, tb, sb) in zip(with_img, without):
        for i, (x, y) in enumerate(zip(ta, tb)):
            xv, yv = (x.view(torch.int64), y.view(torch.int64)) if x.dtype == torch.float64 else (x, y)
            if not torch.equal(xv, yv):
                bad = (xv != yv).nonzero()
                raise AssertionError("%s: output %d differs at %d places, first %s: %s vs %s" % (
                    tag, i, len(bad), bad[0].tolist(), xv[tuple(bad[0].tolist())].item(), yv[tuple(bad[0].tolist())].item()))
        assert_same_state(sa, sb, M)


@pytest.mark.parametrize("name,N", [("c1_bridge_ext", 64), ("c4_maze_safehouse", 8)])
def test_state_written_is_honoured(name, N):
    """Lives edited in the state buffer between steps (as the reference's tests do) reach the next step."""
    from cuda_engine import CudaEngine
    from oracle import oracle as orc
    cfgd = pu.CONFIGS[name]
    cfg, m = pu.build(cfgd, N, 5)
    cu, ref = CudaEngine(cfg, m), orc.OracleEnv(cfg, m)
    nf = pu.n_fixed_slots(cfgd)
    rs = np.random.RandomState(1)
    A = cfg.n_agents
    for t in range(30):
        if t % 4 == 1:  # edit an agent's life and a wall's life in a few envs
            for e in range(0, N, 3):
                life, slife, si = int(rs.randint(1, 60)), int(rs.randint(1, 150)), int(rs.randint(0, cu.S))
                cu.eng.fields["life"][e, nf - 1] = life
                cu.eng.fields["static_life"][e, si] = slife
                ref.set_life(e, nf - 1, life)
                ref.set_static_life(e, si, slife)
            cu.eng.state_written()
            cu._cache = None
        actions = rs.randint(0, 6, size=(N, A)).astype(np.int32)
        got, want = cu.step(actions, abi.ACTIONS_DISCRETE), ref.step(actions, abi.ACTIONS_DISCRETE)
        assert np.array_equal(got[0], want[0]), "obs differs at tick %d" % t
        assert np.array_equal(got[1].view(np.uint64), want[1].view(np.uint64)), "reward bits differ at tick %d" % t
        assert np.array_equal(got[2], want[2]) and np.array_equal(got[3], want[3])
        errs = []
        for e in range(0, N, max(1, N // 8)):
            errs += pu.compare_record("env%d tick%d" % (e, t), ref.export(e), cu.export(e), nf, check_obs=False)
        assert not errs, "\n".join(errs[:3])
        done = (want[2] | want[3]).astype(np.uint8)
        if done.any():
            assert np.array_equal(cu.reset(done)[done.astype(bool)], ref.reset(done)[done.astype(bool)])
    cu.close()
    ref.close()


@pytest.mark.parametrize("name,N,lanes,pair", [
    ("c1_bridge_ext", 1024, "32", ""), ("c1_bridge_ext", 1024, "16", ""), ("c1_bridge_ext", 1024, "16", "1"),
    ("c5_bridge_channels", 512, "16", ""), ("c5_bridge_channels", 512, "32", ""), ("gym_v0_alone", 300, "16", ""),
    ("c3_city_evac", 128, "32", ""), ("c4_maze_safehouse", 32, "32", "")])
@pytest.mark.parametrize("slots", [1, 0])  # 0 = one slot per step
def test_every_step_of_a_fused_rollout_writes_its_observation(monkeypatch, name, N, lanes, pair, slots):
    from oracle import oracle as orc
    monkeypatch.setenv("ZS_LANES_PER_ENV", lanes)
    if pair:
        monkeypatch.setenv("ZS_NO_TMA_PAIR", pair)
    K = 24
    eng, cfg, m = engine(name, N, seed=29, base=77)
    assert eng.lanes_per_env() == int(lanes)
    ref = orc.OracleEnv(cfg, m)
    if slots == 0:
        ring = eng.new_obs(K)
        ring.fill_(-7)
        eng.rollout(K, 5, None, abi.ACTIONS_DISCRETE, ring, None, None, None)
        got = ring.cpu().numpy().reshape(K, N, -1)
        for s in range(K):
            o, _, _, _ = ref.rollout_synthetic(1, 5 + s)
            assert np.array_equal(got[s], o), "slot %d differs from the oracle's observation after step %d" % (s, s)
    else:
        # one slot: step s+1's bulk copy of the pristine planes overwrites the row step s patched; a rollout that stops
        # after k steps must show exactly step k's observation, for every k
        for k in (1, 2, 3, 7):
            obs = eng.new_obs()
            obs.fill_(-7)
            eng.rollout(k, 100 * k, None, abi.ACTIONS_DISCRETE, obs, None, None, None)
            o, _, _, _ = ref.rollout_synthetic(k, 100 * k)
            assert np.array_equal(obs.cpu().numpy().reshape(N, -1), o), "one-slot rollout of %d steps" % k
    eng.close()
    ref.close()
