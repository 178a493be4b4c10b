"""Image frames of a world (libzombsole_b200/renderer.py: ImageRenderer) against the reference's OpencvRenderer on the same
game under the same draws (tests/golden/image_frames.bin, made by tests/golden/make_image_frames.py): the map region and
the life bars, pixel for pixel.  (The text lines are not compared: they print each player's `status`, which the device
does not keep.)  Dead bodies are painted as zombie remains; the golden game is cut before a player dies."""
import os

import numpy as np
import pytest

import parity_util as pu

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_image_frames_match_reference_renderer():
    pytest.importorskip("PIL")
    from libzombsole_b200.gym_env import ZombsoleVectorEnv
    with open(os.path.join(GOLDEN, "image_frames.bin"), "rb") as f:
        g = np.load(f)
        want_map, want_bars, acts, after = g["map"], g["bars"], g["actions"], g["frame_after_steps"].tolist()
        seed, env_index, steps, h, n = (int(v) for v in g["meta"])
    c = pu.CONFIGS["c1_bridge_ext"]
    env = ZombsoleVectorEnv(c["rules_name"], c["player_names"], c["map_name"], c["agent_ids"][0],
                            initial_zombies=c["initial_zombies"], minimum_zombies=c["minimum_zombies"],
                            observation_scope=c["observation_scope"],
                            observation_position_encoding=c["observation_position_encoding"], agent_weapon=c["agent_weapons"],
                            num_envs=1, seed=seed, env_index_base=env_index, auto_reset=False)
    frames = {0: env.render_image(0)}
    for t in range(steps):
        env.step(np.asarray(acts[t], np.int32).reshape(1, 3))
        if (t + 1) in after:
            frames[t + 1] = env.render_image(0)
    for k, s in enumerate(after):
        img = frames[s]
        assert img.shape == (10 * (h + 2 + n), want_map.shape[2], 3) and img.dtype == np.uint8
        assert np.array_equal(img[: 10 * h], want_map[k]), "map region differs after %d steps" % s
        assert np.array_equal(img[10 * (h + 2): 10 * (h + 2 + n) + 1, : 211], want_bars[k]), "life bars differ after %d steps" % s
    env.close()
