"""Golden image frames: the reference's OpencvRenderer.render (zombsole/renderer.py:241-277) on the states of a reference
game driven by the injected draws, the frame caught where the reference hands it to cv2.imshow.  Only the regions that
do not depend on text are kept (the map, and the life bars): the player lines print each player's `status`, which the
device does not keep, and the text depends on the font the image's Pillow build ships."""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from oracle import ref_harness  # noqa: E402
import parity_util  # noqa: E402

SEED, ENV_INDEX, NAME, STEPS = 78, 6, "c1_bridge_ext", 18


def main():
    cfgd = parity_util.CONFIGS[NAME]
    tape = parity_util.action_tape(cfgd, STEPS, 999, wild=0.0)
    frames = []
    with ref_harness.injected_draws(SEED) as rng:
        runner = ref_harness.RefRunner(cfgd, ENV_INDEX, rng)
        import zombsole.renderer as rr
        caught = []
        rr.cv2 = types.SimpleNamespace(imshow=lambda name, img: caught.append(img[:, :, ::-1].copy()), waitKey=lambda ms: None)
        game = runner.env.game
        w, h = game.map.size
        n = len(game.players) + len(game.agents)
        ren = rr.OpencvRenderer(w, h + 2 + n)

        def frame():
            allplayers = sorted(game.agents, key=lambda x: x.agent_id) + sorted(game.players, key=lambda x: x.name)
            ren.render(game.world, allplayers)
            return caught.pop()
        frames.append(frame())
        steps_done = [0]
        for t in range(STEPS):
            rec = runner.step(tape[t])
            assert not (rec["terminated"] or rec["truncated"])
            if t % 6 == 5:
                frames.append(frame())
                steps_done.append(t + 1)
    frames = np.stack(frames)
    cw = ch = 10
    keep_map = frames[:, : h * ch]                                            # the map
    keep_bars = frames[:, (h + 2) * ch: (h + 2 + n) * ch + 1, : (1 + 20) * cw + 1]  # the life bars (left of the text)
    path = os.path.join(HERE, "image_frames.bin")  # (an .npz by content; not by name: *.npz here are state traces)
    with open(path, "wb") as f:
        np.savez_compressed(f, map=keep_map, bars=keep_bars, actions=np.asarray(tape),
                            frame_after_steps=np.asarray(steps_done), meta=np.asarray([SEED, ENV_INDEX, STEPS, h, n]))
    print("frames", frames.shape, "kept", keep_map.shape, keep_bars.shape, "%.1f KiB" % (os.path.getsize(path) / 1024.0))


if __name__ == "__main__":
    main()
