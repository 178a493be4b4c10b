import random as _random


class Discrete(object):
    def __init__(self, n, start=0):
        self.n, self.start = int(n), int(start)
        self.shape, self.dtype = (), int
        self._rng = _random.Random(0)  # private: must not touch the (injected) global stream

    def sample(self):
        return self.start + self._rng.randrange(self.n)

    def contains(self, x):
        return self.start <= int(x) < self.start + self.n
