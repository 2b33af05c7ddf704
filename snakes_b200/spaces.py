"""The two space types the reference's env exposes (gym.spaces.Discrete / Box), for images that
ship without gym.  Real gym spaces are used instead when gym is importable (registration.py)."""
import numpy as np


class Discrete(object):
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)

    def sample(self):
        return int(np.random.randint(self.n))

    def contains(self, x):
        return 0 <= int(x) < self.n

    def __repr__(self):
        return "Discrete(%d)" % self.n

    def __eq__(self, other):
        return getattr(other, "n", None) == self.n


class Box(object):
    def __init__(self, low, high, shape, dtype=np.uint8):
        self.low, self.high = low, high
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)

    def __repr__(self):
        return "Box(%s, %s, %s, %s)" % (self.low, self.high, self.shape, self.dtype)

    def __eq__(self, other):
        return getattr(other, "shape", None) == self.shape and getattr(other, "dtype", None) == self.dtype
