"""Discrete / Box / Dict spaces: shape + dtype carriers, `sample()` for scripts."""
import numpy as np


class Space(object):
    def __init__(self, shape=None, dtype=None):
        self.shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)


class Discrete(Space):
    def __init__(self, n):
        self.n = int(n)
        Space.__init__(self, (), np.int64)

    def sample(self):
        return int(np.random.randint(self.n))

    def contains(self, x):
        return 0 <= int(x) < self.n

    def __repr__(self):
        return "Discrete(%d)" % self.n

    def __eq__(self, other):
        return isinstance(other, Discrete) and other.n == self.n


class Box(Space):
    def __init__(self, low=None, high=None, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.shape(low)
        Space.__init__(self, shape, dtype)
        self.low = np.full(self.shape, low, dtype=self.dtype) if np.isscalar(low) else np.asarray(low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype) if np.isscalar(high) else np.asarray(high, dtype=self.dtype)

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(self.dtype)

    def __repr__(self):
        return "Box%s" % (self.shape,)

    def __eq__(self, other):
        return isinstance(other, Box) and other.shape == self.shape and other.dtype == self.dtype


class Dict(Space):
    def __init__(self, spaces=None):
        self.spaces = dict(spaces or {})
        Space.__init__(self, None, None)
