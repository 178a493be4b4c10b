import numpy as np


class Space(object):
    shape = None
    dtype = None


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)


class Text(Space):
    def __init__(self, max_length, **kwargs):
        self.max_length = max_length


class Dict(Space):
    def __init__(self, spaces=None, **kwargs):
        self.spaces = dict(spaces or {}, **kwargs)


class Sequence(Space):
    def __init__(self, space, **kwargs):
        self.feature_space = space


from .discrete import Discrete  # noqa: E402,F401
