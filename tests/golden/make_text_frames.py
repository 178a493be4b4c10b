"""Golden text frames: the reference's TerminalRenderer._draw (zombsole/renderer.py:45-88, basic icons) on the states of
a reference game driven by the injected draws, after the constructor and after a few steps.  The status column of the
player lines (text set inside the reference's next_step implementations) is cut off: the device does not keep it."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from oracle import ref_harness  # noqa: E402
import parity_util  # noqa: E402

SEED, ENV_INDEX, NAME, STEPS = 77, 5, "c1_bridge_ext", 25


def cut_status(frame):
    return "\n".join(l.split(">: ")[0] + ">" if ">: " in l else l for l in frame.split("\n"))


def main():
    cfgd = parity_util.CONFIGS[NAME]
    tape = parity_util.action_tape(cfgd, STEPS, 4321)
    frames = []
    with ref_harness.injected_draws(SEED) as rng:
        runner = ref_harness.RefRunner(cfgd, ENV_INDEX, rng)
        import zombsole.renderer as rr
        tr = rr.TerminalRenderer(True)
        game = runner.env.game

        def frame():
            allplayers = sorted(game.agents, key=lambda x: x.agent_id) + sorted(game.players, key=lambda x: x.name)
            return cut_status(tr._draw(game.world, allplayers))
        frames.append(frame())
        for t in range(STEPS):
            rec = runner.step(tape[t])
            assert not (rec["terminated"] or rec["truncated"])
            if t % 6 == 5 or t == STEPS - 1:
                frames.append(frame())
    with open(os.path.join(HERE, "text_frames.json"), "w") as f:
        json.dump({"config": NAME, "seed": SEED, "env_index": ENV_INDEX, "steps": STEPS, "actions": np.asarray(tape).tolist(),
                   "frame_after_steps": [0] + [t + 1 for t in range(STEPS) if t % 6 == 5 or t == STEPS - 1], "frames": frames}, f)
    print(frames[-1])


if __name__ == "__main__":
    main()
