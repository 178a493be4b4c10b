"""world_size-2 gloo test of the N>1 path's host logic: contiguous sharding by global env index,
shard-independent trajectories (the draw counter is the GLOBAL index), and the stats all-reduce.
The per-rank engine here is the CPU oracle; on GPUs the same shard arithmetic feeds ZsEngine."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import parity_util as pu
from libzombsole_b200.distributed import shard_envs, all_reduce_stats

TOTAL, STEPS, SEED = 24, 60, 99


def test_shard_envs_partition():
    for total in (1, 7, 24, 4096, 1 << 20):
        for world in (1, 2, 3, 8):
            nxt = 0
            for r in range(world):
                base, n = shard_envs(total, r, world)
                assert base == nxt and n >= total // world
                nxt = base + n
            assert nxt == total


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    base, n = shard_envs(TOTAL, rank, world)
    cfg, m = pu.build(pu.CONFIGS["c1_bridge_ext"], n, SEED, env_index_base=base, max_episode_steps=1000, auto_reset=True)
    env = orc.OracleEnv(cfg, m)
    obs, reward, term, trunc = env.rollout_synthetic(STEPS, 0)
    stats = torch.from_numpy(env.stats().copy())
    local_episodes = int(stats[0])
    all_reduce_stats(stats)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), obs=obs, reward=reward, term=term, trunc=trunc,
             stats=stats.numpy(), local_episodes=local_episodes, base=base, n=n)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_equal_one_batch(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from oracle import oracle as orc
    cfg, m = pu.build(pu.CONFIGS["c1_bridge_ext"], TOTAL, SEED, env_index_base=0, max_episode_steps=1000, auto_reset=True)
    env = orc.OracleEnv(cfg, m)
    obs, reward, term, trunc = env.rollout_synthetic(STEPS, 0)
    whole_stats = env.stats()
    parts = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    assert [int(p["base"]) for p in parts] == [0, 12] and [int(p["n"]) for p in parts] == [12, 12]
    assert np.array_equal(np.concatenate([p["obs"] for p in parts]), obs)
    assert np.array_equal(np.concatenate([p["reward"] for p in parts], axis=1).view(np.uint64), reward.view(np.uint64))
    assert np.array_equal(np.concatenate([p["term"] for p in parts], axis=1), term)
    assert np.array_equal(np.concatenate([p["trunc"] for p in parts], axis=1), trunc)
    # the all-reduced statistics are the whole job's on every rank
    for p in parts:
        assert np.array_equal(p["stats"], whole_stats)
    assert sum(int(p["local_episodes"]) for p in parts) == int(whole_stats[0]) > 0
