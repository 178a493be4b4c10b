"""Action / observation spaces.  gymnasium's classes are used when gymnasium is installed;
otherwise these minimal stand-ins with the same attributes (shape, dtype, n, low, high,
sample(), contains()) are used so the env classes work in an image without gymnasium."""
import numpy as np

try:  # pragma: no cover - depends on the image
    from gymnasium.spaces import Box, Dict, Text  # noqa: F401
    from gymnasium.spaces.discrete import Discrete  # noqa: F401
    HAVE_GYMNASIUM = True
except ImportError:
    HAVE_GYMNASIUM = False

    class Space(object):
        shape = None
        dtype = None

        def contains(self, x):
            return True

        def __contains__(self, x):
            return self.contains(x)

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)
            self._rng = np.random.default_rng(0)

        def sample(self):
            return self._rng.integers(self.low, self.high + 1, size=self.shape).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low)) and bool(np.all(x <= self.high))

    class Text(Space):
        def __init__(self, max_length, **kwargs):
            self.max_length = max_length

        def contains(self, x):
            return isinstance(x, str) and len(x) <= self.max_length

    class Dict(Space):
        def __init__(self, spaces=None, **kwargs):
            self.spaces = dict(spaces or {}, **kwargs)

        def __getitem__(self, k):
            return self.spaces[k]

    class Discrete(Space):
        def __init__(self, n, start=0):
            self.n, self.start = int(n), int(start)
            self.shape, self.dtype = (), np.dtype(np.int64)
            self._rng = np.random.default_rng(0)

        def sample(self):
            return int(self.start + self._rng.integers(self.n))

        def contains(self, x):
            return self.start <= int(x) < self.start + self.n
