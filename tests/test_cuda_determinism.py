"""Repeat determinism: the same launches on fresh engines give the same bytes, run after run.  A data race in the
kernels (a missing sync between lanes, a store overtaken by a bulk copy, a mailbox read too early) shows up as a run
that differs — this is the check that stands in for a race detector (compute-sanitizer is closed on this pool); the
cross-proxy ordering bug of round 2 was of this kind."""
import hashlib

import pytest
import torch

from libzombsole_b200 import abi
from test_cuda_properties import engine

pytestmark = pytest.mark.gpu


def _digest(name, N, K, slots):
    eng, cfg, _ = engine(name, N, seed=41, base=5)
    obs = eng.new_obs(slots)
    rew, term, trunc = eng.new_outputs(K)
    tape = torch.zeros((K, N, eng.A), dtype=torch.int32, device=eng.device)
    eng.fill_synthetic_tape(3, tape)
    eng.rollout(K, 3, tape, abi.ACTIONS_DISCRETE, obs, rew, term, trunc)          # fused
    for s in range(4):                                                               # single steps from the parked images
        eng.step(tape[s], abi.ACTIONS_DISCRETE, obs[0], rew[s], term[s], trunc[s])
    eng.rollout(5, 0, None, abi.ACTIONS_DISCRETE, obs, None, None, None)           # short, in-kernel actions
    torch.cuda.synchronize()
    h = hashlib.sha256()
    for t in (obs, rew, term, trunc, eng.state):
        h.update(t.cpu().numpy().tobytes())
    eng.close()
    return h.hexdigest()


@pytest.mark.parametrize("name,N,K,slots,lanes", [
    ("c1_bridge_ext", 4096, 24, 3, ""), ("c1_bridge_ext", 4096, 24, 1, ""), ("c1_bridge_ext", 600, 24, 2, "16"),
    ("c5_bridge_channels", 3000, 16, 1, ""), ("c3_city_evac", 512, 16, 1, ""), ("c4_maze_safehouse", 96, 12, 1, ""),
    ("bots_randoman", 512, 24, 2, "")])
def test_runs_repeat_bit_for_bit(monkeypatch, name, N, K, slots, lanes):
    if lanes:
        monkeypatch.setenv("ZS_LANES_PER_ENV", lanes)
    first = _digest(name, N, K, slots)
    for run in range(7):
        assert _digest(name, N, K, slots) == first, "run %d differs from the first" % (run + 1)


def test_producer_variant_repeats(monkeypatch):
    monkeypatch.setenv("ZS_PRODUCER", "1")
    first = _digest("c1_bridge_ext", 4096, 24, 1)
    for run in range(5):
        assert _digest("c1_bridge_ext", 4096, 24, 1) == first
