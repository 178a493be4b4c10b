"""Kernel-side timeline of ONE isolated zs_step_host launch (development build, -DZS_TRACE; run through
tools/trace_launch.sh with TRACE_PY=tools/trace_host_step.py): when the warps enter, finish their step and exit,
relative to the first warp's entry — next to the host-side figures of zs_step_host_stats."""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import numpy as np
import torch
from trace_launch import grab, q
from libzombsole_b200 import _native
from libzombsole_b200.gym_env import ZombsoleVectorEnv

KW = dict(rules_name="extermination", player_names=["terminator", "terminator"], map_name="bridge", agent_id=0,
          initial_zombies=10, minimum_zombies=0, observation_scope="world", agent_weapon="rifle")
N = 4096
env = ZombsoleVectorEnv(num_envs=N, seed=0, max_episode_steps=1000, host_outputs="compact", **KW)
acts = torch.from_numpy(np.random.RandomState(0).randint(0, 6, size=(400, N)).astype(np.int32)).pin_memory()
for s in range(100):
    env.step(acts[s])
env.engine.step_host_stats()
L = _native.lib()
for s in range(100, 160):
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 60e-6:
        pass
    env.step(acts[s])
torch.cuda.synchronize()
print("host side: calls %d, mean us from entry to launches issued %.1f / previous cells restored %.1f / flag seen %.1f / return %.1f"
      % env.engine.step_host_stats())
t = grab(L)
t = t[t[:, 0] > 0].astype(np.int64)
t0 = t[:, 0].min()
print("kernel side (last launch): %d warps" % len(t))
print("  entry after first entry   ", q(t[:, 0] - t0))
print("  staging                   ", q(t[:, 1] - t[:, 0]))
print("  load_state                ", q(t[:, 2] - t[:, 1]))
print("  template wait             ", q(t[:, 29] - t[:, 3]))
print("  step                      ", q(t[:, 4] - t[:, 29]))
print("    template issue          ", q(t[:, 24] - t[:, 29]))
print("    world_step              ", q(t[:, 25] - t[:, 24]))
print("    reward/rules/out        ", q(t[:, 26] - t[:, 25]))
print("    world init              ", q(t[:, 27] - t[:, 26]))
print("    record + fence + ticket ", q(t[:, 4] - t[:, 27]))
print("  step end -> exit          ", q(t[:, 30] - t[:, 4]))
print("  step end after first entry", q(t[:, 4] - t0))
print("  exit after first entry    ", q(t[:, 30] - t0))
env.close()
