"""Where an end-to-end step of host_outputs="compact" spends its time (host clock, phase by phase)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import torch
from libzombsole_b200.gym_env import ZombsoleVectorEnv
from libzombsole_b200 import abi

KW = dict(rules_name="extermination", player_names=["terminator", "terminator"], map_name="bridge", agent_id=0,
          initial_zombies=10, minimum_zombies=0, observation_scope="world", agent_weapon="rifle")
N = 4096
threads = int(sys.argv[1]) if len(sys.argv) > 1 else 0
env = ZombsoleVectorEnv(num_envs=N, seed=0, max_episode_steps=1000, host_outputs="compact", host_threads=threads, **KW)
acts = torch.from_numpy(np.random.RandomState(0).randint(0, 6, size=(400, N)).astype(np.int32)).pin_memory()
for s in range(50):
    env.step(acts[s])
eng = env.engine
stream = torch.cuda.current_stream(env.device)
T = dict(stage=0.0, launch=0.0, d2h=0.0, sync=0.0, expand=0.0, total=0.0)
n = 300
for s in range(n):
    t0 = time.perf_counter()
    a, fmt = env._stage_actions(acts[50 + s])
    t1 = time.perf_counter()
    eng.step_compact(a, fmt, env._records, env._dev_obs)
    t2 = time.perf_counter()
    env._records_host.copy_(env._records, non_blocking=True)
    t3 = time.perf_counter()
    stream.synchronize()
    t4 = time.perf_counter()
    over = eng.expand_compact(env._records_host, env._records_prev, env.obs, env.reward, env._term, env._trunc, env._overflow,
                              False, env.host_threads)
    t5 = time.perf_counter()
    T["stage"] += t1 - t0; T["launch"] += t2 - t1; T["d2h"] += t3 - t2; T["sync"] += t4 - t3; T["expand"] += t5 - t4; T["total"] += t5 - t0
print("threads", threads, {k: round(v / n * 1e6, 1) for k, v in T.items()}, "us per step ->", "%.3e env-steps/s" % (N * n / T["total"]))
ent = (env._records_host[:, 0] & 0xffff).float()
print("entries per record: mean %.1f max %d" % (ent.mean().item(), int(ent.max().item())))
env.close()
