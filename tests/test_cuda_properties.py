"""GPU tests at BASELINE.json's full sizes: direct parity against the multi-threaded oracle where it
finishes in seconds, and size-independent properties (fused rollout == K single steps, shard
independence, encode idempotence, Philox known answers on the device)."""
import numpy as np
import pytest
import torch

import parity_util as pu
from libzombsole_b200 import abi, philox

pytestmark = pytest.mark.gpu


def engine(name, N, seed=0, base=0, mes=1000, auto_reset=True):
    from libzombsole_b200.engine import ZsEngine
    cfg, m = pu.build(pu.CONFIGS[name], N, seed, env_index_base=base, max_episode_steps=mes, auto_reset=auto_reset)
    return ZsEngine(cfg, m), cfg, m


def state_snapshot(eng):
    torch.cuda.synchronize()
    return {k: v.clone() for k, v in eng.fields.items()}


def assert_same_state(a, b, M):
    for k in a:
        x, y = a[k], b[k]
        if k == "stamp":  # only the order matters; compare the order of in-world things
            inw = (a["meta"][:, :M] & 0x80) != 0
            big = torch.iinfo(torch.int32).max
            ox = torch.where(inw, x[:, :M], torch.full_like(x[:, :M], big)).argsort(dim=1, stable=True)
            oy = torch.where(inw, y[:, :M], torch.full_like(y[:, :M], big)).argsort(dim=1, stable=True)
            assert torch.equal(ox, oy), k
        elif k == "scalars":
            keep = [abi.S_T, abi.S_EPISODE, abi.S_DEATHS, abi.S_ZOMBIE_DEATHS, abi.S_FLAGS, abi.S_PREV_ZOMBIE_DEATHS, abi.S_EPISODE_STEPS]
            assert torch.equal(x[:, keep], y[:, keep]), k
        else:
            assert torch.equal(x, y), k


def test_config2_full_size_matches_oracle():
    """BASELINE configs[1]: 4,096 bridge/extermination envs, synthetic actions, 64 steps, auto-reset —
    every reward bit, flag and the final observation equal the oracle's."""
    from oracle import oracle as orc
    N, K = 4096, 64
    eng, cfg, m = engine("c1_bridge_ext", N, seed=3)
    obs = eng.new_obs()
    rew, term, trunc = eng.new_outputs(K)
    eng.rollout(K, 0, None, abi.ACTIONS_DISCRETE, obs, rew, term, trunc)
    ref = orc.OracleEnv(cfg, m)
    o, r, te, tr = ref.rollout_synthetic(K, 0)
    assert np.array_equal(obs.cpu().numpy().reshape(N, -1), o)
    assert np.array_equal(rew.cpu().numpy().view(np.uint64).reshape(K, N, 1), r.view(np.uint64))
    assert np.array_equal(term.cpu().numpy(), te) and np.array_equal(trunc.cpu().numpy(), tr)
    assert np.array_equal(eng.episode_stats().cpu().numpy(), ref.stats())
    assert te.sum() > 1000  # thousands of episodes ended and were re-initialised on the device
    eng.close()


@pytest.mark.parametrize("name,N,K", [("c3_city_evac", 2048, 40), ("c4_maze_safehouse", 512, 24), ("c5_bridge_channels", 2048, 48)])
def test_other_configs_large_batch_match_oracle(name, N, K):
    from oracle import oracle as orc
    eng, cfg, m = engine(name, N, seed=5, base=1 << 20)
    obs = eng.new_obs()
    rew, term, trunc = eng.new_outputs(K)
    eng.rollout(K, 7, None, abi.ACTIONS_DISCRETE, obs, rew, term, trunc)
    ref = orc.OracleEnv(cfg, m)
    o, r, te, tr = ref.rollout_synthetic(K, 7)
    assert np.array_equal(obs.cpu().numpy().reshape(N, -1), o)
    assert np.array_equal(rew.cpu().numpy().view(np.uint64).reshape(r.shape), r.view(np.uint64))
    assert np.array_equal(term.cpu().numpy(), te) and np.array_equal(trunc.cpu().numpy(), tr)
    eng.close()


@pytest.mark.parametrize("name,N", [("c1_bridge_ext", 4096), ("c3_city_evac", 1024), ("survival_minz", 512)])
def test_fused_rollout_equals_single_steps(name, N):
    """One K-step launch == K one-step launches with auto-reset (outputs, final state)."""
    K = 50
    a, cfg, _ = engine(name, N, seed=11)
    b, _, _ = engine(name, N, seed=11)
    A = cfg.n_agents
    tape = torch.zeros((K, N, A), dtype=torch.int32, device=a.device)
    for s in range(K):
        a.fill_synthetic_actions(100 + s, tape[s])
    obs_a, obs_b = a.new_obs(), b.new_obs()
    ra, ta, ua = a.new_outputs(K)
    rb, tb, ub = b.new_outputs(K)
    a.rollout(K, 0, tape, abi.ACTIONS_DISCRETE, obs_a, ra, ta, ua)
    for s in range(K):
        b.step(tape[s], abi.ACTIONS_DISCRETE, obs_b, rb[s], tb[s], ub[s])
    assert torch.equal(obs_a, obs_b) and torch.equal(ra.view(torch.int64), rb.view(torch.int64))
    assert torch.equal(ta, tb) and torch.equal(ua, ub)
    assert_same_state(state_snapshot(a), state_snapshot(b), a.M)
    # the in-kernel synthetic stream is the same tape
    c, _, _ = engine(name, N, seed=11)
    obs_c = c.new_obs()
    rc, tc, uc = c.new_outputs(K)
    c.rollout(K, 100, None, abi.ACTIONS_DISCRETE, obs_c, rc, tc, uc)
    assert torch.equal(obs_a, obs_c) and torch.equal(ra.view(torch.int64), rc.view(torch.int64))
    for e in (a, b, c):
        e.close()


@pytest.mark.parametrize("name,N,mes", [("c1_bridge_ext", 1024, 7), ("gym_v0_alone", 777, 3), ("multi_boxed_2p", 300, 5)])
def test_time_limit_inside_fused_rollout_matches_oracle(name, N, mes):
    """gymnasium's TimeLimit (max_episode_steps) truncates and re-initialises inside one fused launch."""
    from oracle import oracle as orc
    K = 40
    eng, cfg, m = engine(name, N, seed=(1 << 40) + 17, base=9, mes=mes)
    obs = eng.new_obs()
    rew, term, trunc = eng.new_outputs(K)
    eng.rollout(K, 3, None, abi.ACTIONS_DISCRETE, obs, rew, term, trunc)
    ref = orc.OracleEnv(cfg, m)
    o, r, te, tr = ref.rollout_synthetic(K, 3)
    assert np.array_equal(obs.cpu().numpy().reshape(N, -1), o)
    assert np.array_equal(rew.cpu().numpy().view(np.uint64).reshape(r.shape), r.view(np.uint64))
    assert np.array_equal(term.cpu().numpy(), te) and np.array_equal(trunc.cpu().numpy(), tr)
    assert tr.sum() >= N * (K // mes) // 2  # truncations did happen
    eng.close()


def test_shard_independence():
    """Env g evolves identically whether it is env g of one batch or env 0 of a shard based at g."""
    K = 40
    whole, _, _ = engine("c1_bridge_ext", 96, seed=21, base=1000)
    obs_w = whole.new_obs()
    rw, tw, uw = whole.new_outputs(K)
    whole.rollout(K, 0, None, abi.ACTIONS_DISCRETE, obs_w, rw, tw, uw)
    for base, n in ((1000, 32), (1032, 64)):
        part, _, _ = engine("c1_bridge_ext", n, seed=21, base=base)
        obs_p = part.new_obs()
        rp, tp, up = part.new_outputs(K)
        part.rollout(K, 0, None, abi.ACTIONS_DISCRETE, obs_p, rp, tp, up)
        lo = base - 1000
        assert torch.equal(obs_p, obs_w[lo:lo + n]) and torch.equal(rp.view(torch.int64), rw[:, lo:lo + n].view(torch.int64))
        assert torch.equal(tp, tw[:, lo:lo + n])
        part.close()
    whole.close()


@pytest.mark.parametrize("name", ["c1_bridge_ext", "c5_bridge_channels", "gym_surroundings", "c3_city_evac", "multi_fort_32p"])
def test_encode_obs_is_idempotent_and_matches_step_obs(name):
    eng, cfg, _ = engine(name, 256, seed=2, auto_reset=False)
    obs, obs2 = eng.new_obs(), eng.new_obs()
    rew, term, trunc = eng.new_outputs()
    acts = torch.zeros((256, cfg.n_agents), dtype=torch.int32, device=eng.device)
    for s in range(30):
        eng.fill_synthetic_actions(s, acts)
        eng.step(acts, abi.ACTIONS_DISCRETE, obs, rew, term, trunc)
    eng.encode_obs(obs2)
    assert torch.equal(obs, obs2)
    before = state_snapshot(eng)
    eng.encode_obs(obs2)
    assert torch.equal(obs, obs2)
    after = state_snapshot(eng)
    for k in before:
        assert torch.equal(before[k], after[k])
    eng.close()


def test_observation_value_ranges():
    """simple encoding: 256*thing + 16*weapon + scaled life (observation.py:47-53), fresh walls 1039, boxes 257."""
    eng, cfg, m = engine("c1_bridge_ext", 512, seed=9)
    obs = eng.new_obs(4)
    eng.rollout(200, 0, None, abi.ACTIONS_DISCRETE, obs, None, None, None)
    o = obs[0].cpu().numpy()
    assert o.min() >= 0 and o.max() < 2048
    codes = o >> 8
    assert set(np.unique(codes)).issubset({0, 1, 2, 3, 4, 5, 6, 7})
    assert ((codes == 7).sum(axis=(1, 2, 3)) <= 1).all()           # at most one agent per env
    assert ((codes == 6).sum(axis=(1, 2, 3)) <= 2).all()           # two terminators
    assert ((codes == 5).sum(axis=(1, 2, 3)) <= 10).all()          # ten zombies
    weapons = (o >> 4) & 15
    assert set(np.unique(weapons[codes == 5])) <= {1} and set(np.unique(weapons[codes == 7])) <= {13}
    eng.close()


def test_device_synthetic_actions_match_host_philox():
    eng, cfg, _ = engine("c3_city_evac", 300, seed=77, base=12345)
    acts = torch.zeros((300, 4), dtype=torch.int32, device=eng.device)
    for step in (0, 5, 1 << 20):
        eng.fill_synthetic_actions(step, acts)
        want = philox.synthetic_actions(77, 12345, 300, 4, step, 7)
        assert np.array_equal(acts.cpu().numpy(), want)
    eng.close()


def test_vector_env_public_api():
    from libzombsole_b200.gym_env import ZombsoleVectorEnv
    from libzombsole_b200.gym.multiagent_env import MultiagentZombsoleVectorEnv
    env = ZombsoleVectorEnv("extermination", ["terminator", "terminator"], "bridge", 0, initial_zombies=10, num_envs=128, seed=1)
    obs, info = env.reset()
    assert obs.shape == (128, 1, 12, 111) and obs.dtype == torch.int32 and obs.is_cuda
    total = 0
    for s in range(80):
        obs, reward, term, trunc, info = env.step(torch.randint(0, 6, (128,), device=env.device))
        assert reward.dtype == torch.float64 and term.dtype == torch.bool
        total += int((term | trunc).sum())
    assert total > 0  # episodes end and are re-initialised in the same call
    # dict actions and (type, dx, dy) rows are accepted too
    env.step([{"action_type": "attack_closest"}] * 128)
    env.step(np.tile(np.array([[abi.ACT_MOVE, 1, 0]], np.int32), (128, 1)))
    g = env.game(5)
    assert len(g.agents) == 1 and len(g.players) == 2 and g.world.size == (111, 12)
    env.close()
    menv = MultiagentZombsoleVectorEnv("evacuation", [], "city_for_evacuation", ["0", "1", "2", "3"], initial_zombies=20, num_envs=64)
    obs, reward, term, trunc, info = menv.step(torch.randint(-1, 7, (64, 4), device=menv.device))
    assert obs.shape == (64, 4, 3, 21, 21) and reward.shape == (64, 4) and info["agent_mask"].shape == (64, 4)
    menv.close()


@pytest.mark.parametrize("name,N,K,lanes", [("c1_bridge_ext", 4096, 64, "16"), ("c1_bridge_ext", 4096, 64, "32"),
                                            ("gym_v0_alone", 1024, 80, "16"), ("safehouse_small", 1000, 64, "32")])
def test_standard_rollout_shape_matches_oracle(name, N, K, lanes, monkeypatch):
    """The rollout shape with its own compiled-in kernel (one agent, world observation, discrete action tape in,
    observation / reward / flags out, nothing else): every reward bit, flag and the final observation vs the oracle."""
    from oracle import oracle as orc
    monkeypatch.setenv("ZS_LANES_PER_ENV", lanes)
    eng, cfg, m = engine(name, N, seed=21)
    assert eng.lanes_per_env() == int(lanes)
    obs = eng.new_obs(3)
    rew, term, trunc = eng.new_outputs(K)
    acts = torch.zeros((K, N, eng.A), dtype=torch.int32, device=eng.device)
    for s in range(K):
        eng.fill_synthetic_actions(s, acts[s])
    eng.rollout(K, 0, acts, abi.ACTIONS_DISCRETE, obs, rew, term, trunc)
    ref = orc.OracleEnv(cfg, m)
    o, r, te, tr = ref.rollout_synthetic(K, 0)
    assert np.array_equal(obs[(K - 1) % 3].cpu().numpy().reshape(N, -1), o)
    assert np.array_equal(rew.cpu().numpy().view(np.uint64).reshape(K, N, 1), r.view(np.uint64))
    assert np.array_equal(term.cpu().numpy(), te) and np.array_equal(trunc.cpu().numpy(), tr)
    assert np.array_equal(eng.episode_stats().cpu().numpy(), ref.stats())
    eng.close()


def test_host_output_mode_matches_device_outputs():
    """ZombsoleVectorEnv(host_outputs=True): the kernel writes observation / reward / flags straight into pinned host
    memory and reads the pinned host action tensor in place; same values as the device-output env."""
    from libzombsole_b200.gym_env import ZombsoleVectorEnv
    kw = dict(rules_name="extermination", player_names=["terminator", "terminator"], map_name="bridge", agent_id=0,
              initial_zombies=10, minimum_zombies=0, num_envs=512, seed=9, max_episode_steps=50)
    dev_env, host_env = ZombsoleVectorEnv(**kw), ZombsoleVectorEnv(host_outputs=True, **kw)
    o0, _ = dev_env.reset()
    h0, _ = host_env.reset()
    assert h0.device.type == "cpu" and h0.is_pinned() and torch.equal(o0.cpu(), h0)
    g = torch.Generator().manual_seed(1)
    for t in range(120):
        a = torch.randint(0, 6, (512,), generator=g, dtype=torch.int32).pin_memory()
        o, r, te, tr, _ = dev_env.step(a)
        ho, hr, hte, htr, _ = host_env.step(a)
        assert torch.equal(o.cpu(), ho) and torch.equal(r.cpu().view(torch.int64), hr.view(torch.int64))
        assert torch.equal(te.cpu(), hte) and torch.equal(tr.cpu(), htr)
    dev_env.close()
    host_env.close()


def test_multiagent_host_output_mode_matches_device_outputs():
    """MultiagentZombsoleVectorEnv(host_outputs=True): observations, per-agent rewards, flags and the agent mask land in pinned
    host memory, the agents' lives after the step in agent_life_host; same values as the device-output env."""
    from libzombsole_b200.gym.multiagent_env import MultiagentZombsoleVectorEnv
    kw = dict(rules_name="evacuation", player_names=["terminator"], map_name="village_for_evacuation", agent_ids=["0", "1", "2"],
              initial_zombies=15, minimum_zombies=0, num_envs=96, seed=5, max_episode_steps=40)
    dev_env, host_env = MultiagentZombsoleVectorEnv(**kw), MultiagentZombsoleVectorEnv(host_outputs=True, **kw)
    o0, _ = dev_env.reset()
    h0, _ = host_env.reset()
    assert h0.device.type == "cpu" and h0.is_pinned() and torch.equal(o0.cpu(), h0)
    g = torch.Generator().manual_seed(2)
    P = dev_env.cfg.n_bots
    for t in range(90):
        a = torch.randint(-1, 7, (96, 3), generator=g, dtype=torch.int32)
        o, r, te, tr, info = dev_env.step(a.cuda())
        ho, hr, hte, htr, hinfo = host_env.step(a.pin_memory() if t % 2 else a)
        assert torch.equal(o.cpu(), ho) and torch.equal(r.cpu().view(torch.int64), hr.view(torch.int64))
        assert torch.equal(te.cpu(), hte) and torch.equal(tr.cpu(), htr)
        assert torch.equal(info["agent_mask"].cpu(), hinfo["agent_mask"])
        assert torch.equal(dev_env.engine.fields["life"][:, P:P + 3].cpu(), host_env.agent_life_host)
    dev_env.close()
    host_env.close()


def test_tape_fill_equals_per_step_fill():
    """zs_fill_synthetic_tape (one launch) writes the same ids as zs_fill_synthetic_actions step by step."""
    eng, cfg, _ = engine("c3_city_evac", 300, seed=9, base=17)
    K = 11
    tape = torch.zeros((K, eng.N, eng.A), dtype=torch.int32, device=eng.device)
    eng.fill_synthetic_tape(40, tape)
    one = torch.zeros((eng.N, eng.A), dtype=torch.int32, device=eng.device)
    for s in range(K):
        eng.fill_synthetic_actions(40 + s, one)
        assert torch.equal(tape[s], one), s
    eng.close()
