#!/usr/bin/env python
"""Generate the committed golden fixtures from the UNMODIFIED reference.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
For every parity configuration (tests/parity_util.py:CONFIGS) it runs E reference envs
(global env indices BASE..BASE+E-1, seed SEED) through oracle/ref_harness.py under the
injected Philox draw stream with a seeded action tape, with a same-tick reset whenever an
episode ends, and stores everything the harness dumps (slot state, dict order, static
lives, dead-body cells, counters, observation, float64 reward bits, flags, draws consumed)
after the constructor, after every step and after every reset.

The fixtures are what the GPU box checks the CUDA path against (the reference tree does
not exist there); tests/test_oracle_golden.py pins the C oracle to the same files.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_harness  # noqa: E402
import parity_util  # noqa: E402

SEED = 20260118
BASE = 11
#: name -> (E envs, T ticks, max_episode_steps)
PLAN = {
    "c1_bridge_ext": (3, 160, 0),
    "c5_bridge_channels": (2, 80, 0),
    "gym_v0_alone": (2, 80, 30),
    "gym_surroundings": (2, 80, 0),
    "surroundings_channels": (2, 80, 0),
    "c3_city_evac": (2, 120, 0),
    "village_evac_mixed": (2, 100, 0),
    "c4_maze_safehouse": (2, 50, 0),
    "safehouse_small": (2, 100, 0),
    "multi_boxed_2p": (2, 80, 0),
    "multi_fort_32p": (1, 40, 0),
    "survival_minz": (2, 100, 40),
    "minz_allcells": (2, 80, 0),
    "bots_mixed": (2, 60, 0),
    "bots_hamsters": (2, 100, 0),
    "bots_randoman": (2, 120, 0),
    "randoman_crowd": (2, 60, 0),
    "box_arena": (2, 60, 0),
}


def main():
    total = 0
    only = set(sys.argv[1:])  # optional: regenerate just the named fixtures
    for name, (E, T, mes) in PLAN.items():
        if only and name not in only:
            continue
        cfgd = parity_util.CONFIGS[name]
        out = {"meta": np.array([SEED, BASE, E, T, mes], np.int64)}
        resets = 0
        for e in range(E):
            tr = ref_harness.run_trace(cfgd, BASE + e, SEED, parity_util.action_tape(cfgd, T, 1000 * (e + 1) + len(name)),
                                       max_episode_steps=mes)
            resets += int(tr["did_reset"].sum())
            for k, v in tr.items():
                out["e%d_%s" % (e, k)] = v
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        size = os.path.getsize(path)
        total += size
        print("%-24s E=%d T=%d resets=%d  %.1f KiB" % (name, E, T, resets, size / 1024.0))
    print("total %.1f KiB" % (total / 1024.0))


if __name__ == "__main__":
    main()
