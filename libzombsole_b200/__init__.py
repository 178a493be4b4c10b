"""libzombsole_b200 — B200-native batched zombsole simulator.

The reference's Gymnasium / multi-agent env API (reset / step / observation / reward) with the
world state of N independent games resident in HBM and every transition computed by hand-written
sm_100a CUDA kernels behind a C ABI (include/zs_b200.h).  Importing the package does not need a
GPU; constructing an environment does — there is no CPU fallback.
"""
__version__ = "0.1.0"

from .maps import Map, stock_map_names  # noqa: F401


def __getattr__(name):
    # env classes import torch; keep `import libzombsole_b200` light
    if name in ("ZombsoleGymEnv", "ZombsoleGymEnvDiscreteAction", "ZombsoleVectorEnv", "make_vector"):
        from . import gym_env
        return getattr(gym_env, name)
    if name in ("MultiagentZombsoleEnv", "MultiagentZombsoleEnvDiscreteAction", "MultiagentZombsoleVectorEnv"):
        from .gym import multiagent_env
        return getattr(multiagent_env, name)
    if name == "ZsEngine":
        from .engine import ZsEngine
        return ZsEngine
    raise AttributeError(name)
