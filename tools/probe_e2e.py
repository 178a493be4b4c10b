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
# "direct": the kernel reads the pinned action buffer and writes the pinned record buffer itself (no copies, one sync)
direct = len(sys.argv) > 2 and sys.argv[2] == "direct"
streamed = len(sys.argv) > 2 and sys.argv[2] == "streamed"  # the product path of host_outputs="compact" (zs_step_host)
env = ZombsoleVectorEnv(num_envs=N, seed=0, max_episode_steps=1000, host_outputs="compact" if streamed else "compact-copy",
                        host_threads=threads, **KW)
acts = torch.from_numpy(np.random.RandomState(0).randint(0, 6, size=(400, N)).astype(np.int32)).pin_memory()
for s in range(50):
    env.step(acts[s])
if streamed:
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    env.engine.step_host_stats()
    for s in range(300):
        env.step(acts[50 + s])
    dt = time.perf_counter() - t0
    print("   inside zs_step_host: calls %d, mean us from entry to launches issued %.1f / previous cells restored %.1f / flag seen %.1f / return %.1f"
          % env.engine.step_host_stats())
    print("streamed threads", threads, "%.1f us per step -> %.3e env-steps/s" % (dt / 300 * 1e6, N * 300 / dt))
    env.close()
    sys.exit(0)
eng = env.engine
stream = torch.cuda.current_stream(env.device)
T = dict(stage=0.0, launch=0.0, d2h=0.0, sync=0.0, expand=0.0, total=0.0)
n = 300
for s in range(n):
    t0 = time.perf_counter()
    if direct:
        t1 = t0
        eng.step_compact(acts[50 + s].view(N, 1), abi.ACTIONS_DISCRETE, env._records_host, env._dev_obs)
        t3 = t2 = time.perf_counter()
    else:
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
print("direct" if direct else "copies", "threads", threads, {k: round(v / n * 1e6, 1) for k, v in T.items()}, "us per step ->", "%.3e env-steps/s" % (N * n / T["total"]))
ent = (env._records_host[:, 0] & 0xffff).float()
print("entries per record: mean %.1f max %d" % (ent.mean().item(), int(ent.max().item())))
env.close()
