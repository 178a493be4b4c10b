import sys, time
sys.path.insert(0, ".")
import numpy as np
from libzombsole_b200.gym_env import ZombsoleGymEnvDiscreteAction
env = ZombsoleGymEnvDiscreteAction("extermination", ["terminator", "terminator"], "bridge", 0, initial_zombies=10, minimum_zombies=0)
env.reset()
rs = np.random.RandomState(0)
for i in range(200):
    o, r, te, tr, _ = env.step(int(rs.randint(6)))
    if te or tr: env.reset()
t0 = time.perf_counter(); n = 3000
for i in range(n):
    o, r, te, tr, _ = env.step(int(rs.randint(6)))
    if te or tr: env.reset()
dt = time.perf_counter() - t0
print("single-env drop-in class: %.1f us per step (%.0f steps/s), obs %s %s" % (dt / n * 1e6, n / dt, type(o).__name__, o.dtype))

from libzombsole_b200.gym.multiagent_env import MultiagentZombsoleEnvDiscreteAction
menv = MultiagentZombsoleEnvDiscreteAction("evacuation", [], "city_for_evacuation", ["0", "1", "2", "3"], initial_zombies=20, minimum_zombies=0)
menv.reset()
def acts():
    return {aid: int(rs.randint(7)) for aid in menv.env.agents}
for i in range(100):
    o, r, d, t, _ = menv.step(acts())
    if (d and all(d.values())) or (t and all(t.values())) or not menv.env.agents: menv.reset()
t0 = time.perf_counter(); n = 1500
for i in range(n):
    o, r, d, t, _ = menv.step(acts())
    if (d and all(d.values())) or (t and all(t.values())) or not menv.env.agents: menv.reset()
dt = time.perf_counter() - t0
print("single-env multi-agent drop-in class (config 3's game): %.1f us per step (%.0f steps/s)" % (dt / n * 1e6, n / dt))
