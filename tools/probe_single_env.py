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
