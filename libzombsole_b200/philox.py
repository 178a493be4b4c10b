"""Host-side Philox4x32-10 and the zombsole draw contract.

The reference takes all of its randomness from CPython's ``random`` module
(``shuffle``/``randint``/``choice``, every one of which reduces to
``Random._randbelow(n)``; call sites: zombsole/core.py:54,76,180,198,
zombsole/things.py:62,103,116, zombsole/weapons.py:43).  The batched simulator
replaces that stream with a counter-based one so that every draw is a pure
function of *where* it is asked for:

    u = Philox4x32-10(key = (seed_lo, seed_hi),
                      ctr = (env_index, episode, t_word, k >> 2))[k & 3]
    randbelow(n) = (u * n) >> 32

``env_index`` is the GLOBAL environment index (so results do not depend on how
environments are sharded over GPUs), ``episode`` counts world initialisations
(0 = the one done by the constructor), ``t_word`` is 0 for the draws made while
(re)building the world and ``World.t + 1`` for the draws of a step, and ``k``
counts ``_randbelow`` calls inside that (episode, t_word) cell in the
reference's own call order.

This module is host logic (used by the env classes for synthetic action tapes
and by the tests); the device implementation lives in csrc/zs_philox.cuh and the
oracle's in oracle/zs_oracle.c.  All three are pinned to the same
known-answer vectors in tests/test_philox.py.
"""
import numpy as np

PHILOX_M0 = 0xD2511F53
PHILOX_M1 = 0xCD9E8D57
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
MASK32 = 0xFFFFFFFF

#: XORed into key word 1 for the synthetic-action stream (bench / rollouts), so
#: actions never alias the world's own draws.
ACTION_STREAM_KEY1_XOR = 0xAC710115


def philox4x32_10(ctr, key):
    """One Philox4x32-10 block on Python ints. ctr: 4 words, key: 2 words."""
    c0, c1, c2, c3 = (int(c) & MASK32 for c in ctr)
    k0, k1 = (int(k) & MASK32 for k in key)
    for _ in range(10):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> 32, p0 & MASK32
        hi1, lo1 = p1 >> 32, p1 & MASK32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK32, lo1, (hi0 ^ c3 ^ k1) & MASK32, lo0
        k0 = (k0 + PHILOX_W0) & MASK32
        k1 = (k1 + PHILOX_W1) & MASK32
    return c0, c1, c2, c3


def philox4x32_10_np(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 over numpy uint32 arrays (broadcastable)."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK32 for c in (c0, c1, c2, c3))
    k0 = np.uint64(int(k0) & MASK32)
    k1 = np.uint64(int(k1) & MASK32)
    m0, m1 = np.uint64(PHILOX_M0), np.uint64(PHILOX_M1)
    w0, w1 = np.uint64(PHILOX_W0), np.uint64(PHILOX_W1)
    mask, s32 = np.uint64(MASK32), np.uint64(32)
    for _ in range(10):
        p0 = m0 * c0
        p1 = m1 * c2
        hi0, lo0 = p0 >> s32, p0 & mask
        hi1, lo1 = p1 >> s32, p1 & mask
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & mask, lo1, (hi0 ^ c3 ^ k1) & mask, lo0
        k0 = (k0 + w0) & mask
        k1 = (k1 + w1) & mask
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def seed_key(seed):
    """64-bit seed -> (key0, key1)."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return seed & MASK32, seed >> 32


def world_draw(seed, env_index, episode, t_word, k):
    """The k-th raw 32-bit draw of cell (env_index, episode, t_word)."""
    block = philox4x32_10((env_index, episode, t_word, k >> 2), seed_key(seed))
    return block[k & 3]


def randbelow(u, n):
    """Map a raw 32-bit draw to [0, n) the way the device does (mulhi)."""
    return (int(u) * int(n)) >> 32


def synthetic_actions(seed, env_base, num_envs, num_agents, step_index, n_actions):
    """Uniform discrete action ids for one step, shape [num_envs, num_agents].

    Counter: (global env index, step_index, 0, agent >> 2), word agent & 3, key
    word 1 XOR ACTION_STREAM_KEY1_XOR.  Matches zs_fill_synthetic_actions on the
    device (csrc/zs_b200.cu) and zso_synthetic_actions in the oracle.
    """
    k0, k1 = seed_key(seed)
    k1 ^= ACTION_STREAM_KEY1_XOR
    env = (np.arange(num_envs, dtype=np.uint64) + np.uint64(env_base))[:, None]
    agent = np.arange(num_agents, dtype=np.uint64)[None, :]
    blocks = philox4x32_10_np(env, np.uint64(step_index), np.uint64(0), agent >> np.uint64(2), k0, k1)
    words = np.stack(np.broadcast_arrays(*blocks), axis=-1)  # [N, A, 4]
    sel = (agent & np.uint64(3)).astype(np.int64)
    sel = np.broadcast_to(sel, words.shape[:2])
    u = np.take_along_axis(words, sel[..., None], axis=-1)[..., 0].astype(np.uint64)
    return ((u * np.uint64(n_actions)) >> np.uint64(32)).astype(np.int32)
