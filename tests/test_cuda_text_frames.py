"""Text frames of a world (libzombsole_b200/renderer.py) against the reference's TerminalRenderer._draw on the same
game under the same draws (tests/golden/text_frames.json, status column cut off — see make_text_frames.py)."""
import json
import os

import numpy as np
import pytest

import parity_util as pu
from libzombsole_b200 import abi

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def cut_status(frame):
    return "\n".join(l.split(">: ")[0] + ">" if ">: " in l else l for l in frame.split("\n"))


def test_text_frames_match_reference_renderer():
    from libzombsole_b200.gym_env import ZombsoleVectorEnv
    with open(os.path.join(GOLDEN, "text_frames.json")) as f:
        g = json.load(f)
    c = pu.CONFIGS[g["config"]]
    env = ZombsoleVectorEnv(c["rules_name"], c["player_names"], c["map_name"], c["agent_ids"][0],
                            initial_zombies=c["initial_zombies"], minimum_zombies=c["minimum_zombies"],
                            observation_scope=c["observation_scope"],
                            observation_position_encoding=c["observation_position_encoding"], agent_weapon=c["agent_weapons"],
                            num_envs=1, seed=g["seed"], env_index_base=g["env_index"], auto_reset=False)
    frames = {0: cut_status(env.render_text(0))}
    acts = np.asarray(g["actions"], np.int32)
    for t in range(g["steps"]):
        env.step(acts[t].reshape(1, 3))
        if (t + 1) in g["frame_after_steps"]:
            frames[t + 1] = cut_status(env.render_text(0))
    for n, want in zip(g["frame_after_steps"], g["frames"]):
        assert frames[n] == want, "frame after %d steps differs:\n%s\n--- reference ---\n%s" % (n, frames[n], want)
    env.close()
