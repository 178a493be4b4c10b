"""The general step kernel (csrc/zs_world.cuh: world_step) does not run random.shuffle swap by swap: it reads every
element's final position off chains of the iterations that hit each position.  This is the host-side model of exactly
that procedure (same data structures: head[] = the largest iteration that hits a position, next[] = the next smaller
iteration with the same partner, built 32 iterations at a time), checked against CPython's own shuffle loop
(Lib/random.py: `for i in reversed(range(1, len(x))): j = randbelow(i + 1); x[i], x[j] = x[j], x[i]`) — the loop the
reference runs at core.py:76.  The CUDA side is checked against the oracle by the parity tests of the many-slot configs."""
import random

import pytest


def shuffle_sequential(L, J):
    x = list(range(L))
    for i in range(L - 1, 0, -1):
        j = J[i]
        x[i], x[j] = x[j], x[i]
    fin = [0] * L
    for pos, el in enumerate(x):
        fin[el] = pos
    return fin


def shuffle_chains(L, J, lanes=32):
    head, nxt = [0] * L, [0] * L
    for i0 in range(1, L, lanes):  # one warp round: lanes with the same partner are linked in lane order
        rnd = [i for i in range(i0, min(i0 + lanes, L)) if J[i] != i]
        prev = {}
        for i in rnd:
            lower = [k for k in rnd if k < i and J[k] == J[i]]
            prev[i] = max(lower) if lower else head[J[i]]
        for i in rnd:
            nxt[i] = prev[i]
            if not any(k > i and J[k] == J[i] for k in rnd):
                head[J[i]] = i
    fin = []
    for q in range(L):
        h = head[q]
        while h == 0 and q != 0:
            nq = J[q]
            if nq == q:
                break
            h = nxt[q]
            q = nq
        fin.append(h if h else q)
    return fin


@pytest.mark.parametrize("seed", range(8))
def test_chain_follow_equals_sequential_shuffle(seed):
    rs = random.Random(seed)
    for _ in range(1500):
        L = rs.randint(1, 256)
        J = [0] + [rs.randint(0, i) for i in range(1, L)]
        assert shuffle_chains(L, J) == shuffle_sequential(L, J)


def test_chain_follow_extremes():
    for L in (1, 2, 3, 33, 64, 65, 256):
        assert shuffle_chains(L, list(range(L))) == list(range(L))            # every swap with itself
        J0 = [0] * L                                                          # everything swaps with position 0
        assert shuffle_chains(L, J0) == shuffle_sequential(L, J0)
        Jm = [0] + [i - 1 for i in range(1, L)]                               # a rotation
        assert shuffle_chains(L, Jm) == shuffle_sequential(L, Jm)
