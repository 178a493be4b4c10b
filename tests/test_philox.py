"""Philox4x32-10 known-answer vectors (Random123's kat_vectors) on the host implementation and
the oracle's; the device implementation is pinned to the same vectors in test_cuda_parity.py."""
import numpy as np

from libzombsole_b200 import philox
from oracle import oracle as orc

KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_host_philox_kat():
    for ctr, key, want in KAT:
        assert philox.philox4x32_10(ctr, key) == want
        got = philox.philox4x32_10_np(*ctr, *key)
        assert tuple(int(g) for g in got) == want


def test_oracle_philox_kat():
    L = orc.lib()
    for ctr, key, want in KAT:
        c = np.array(ctr, np.uint32)
        k = np.array(key, np.uint32)
        out = np.zeros(4, np.uint32)
        L.zso_philox4x32_10(c.ctypes.data, k.ctypes.data, out.ctypes.data)
        assert tuple(int(v) for v in out) == want


def test_randbelow_is_mulhi():
    assert philox.randbelow(0xFFFFFFFF, 6) == 5
    assert philox.randbelow(0, 6) == 0
    assert philox.randbelow(0x80000000, 51) == 25


def test_synthetic_actions_match_oracle():
    import parity_util as pu
    for name in ("c1_bridge_ext", "c3_city_evac"):
        cfg, m = pu.build(pu.CONFIGS[name], 9, seed=5, env_index_base=1000)
        eng = orc.OracleEnv(cfg, m)
        for step in (0, 1, 77):
            want = philox.synthetic_actions(5, 1000, 9, cfg.n_agents, step, 7 if cfg.obs_per_agent else 6)
            assert np.array_equal(eng.synthetic_actions(step), want)
        eng.close()
