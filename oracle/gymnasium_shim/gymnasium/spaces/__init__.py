"""Space stand-ins: attributes only, plus Discrete.contains/sample (see package docstring)."""
import numpy as np


class Space:
    shape = None
    dtype = None

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)

    @property
    def np_random(self):
        if not hasattr(self, "_rng"):
            self._rng = np.random.default_rng()
        return self._rng


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.shape = tuple(shape) if shape is not None else np.shape(low)
        self.dtype = np.dtype(dtype)
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))


class Discrete(Space):
    def __init__(self, n, start=0):
        self.n = int(n)
        self.start = int(start)
        self.shape = ()
        self.dtype = np.dtype(np.int64)

    def contains(self, x):
        if isinstance(x, (int, np.integer)):
            v = int(x)
        elif isinstance(x, np.ndarray) and x.shape == () and np.issubdtype(x.dtype, np.integer):
            v = int(x)
        else:
            return False
        return self.start <= v < self.start + self.n

    def sample(self):
        return int(self.start + self.np_random.integers(self.n))


class Dict(Space, dict):
    def __init__(self, spaces=None, **kw):
        dict.__init__(self, spaces or {}, **kw)
        self.spaces = self
