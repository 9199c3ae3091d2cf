"""Minimal stand-ins for the gym spaces the reference inspects (class name, .shape, .n, .low/.high).

The reference dispatches on ``space.__class__.__name__`` (``mappo/utils/util.py:29-51``), so the
class names must be exactly ``Box`` and ``Discrete``.
"""
import numpy as np


class Box:
    def __init__(self, low, high, dtype=np.float32):
        self.low = np.asarray(low, dtype=dtype)
        self.high = np.asarray(high, dtype=dtype)
        self.shape = self.low.shape
        self.dtype = np.dtype(dtype)

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


class Discrete:
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)

    def sample(self):
        return int(np.random.randint(self.n))

    def contains(self, x):
        return 0 <= int(x) < self.n

    def __repr__(self):
        return f"Discrete({self.n})"


class MultiAgentSpace(list):
    """list of per-agent spaces with ma-gym's ``sample``/``contains`` helpers."""

    def sample(self):
        return [s.sample() for s in self]

    def contains(self, xs):
        return all(s.contains(x) for s, x in zip(self, xs))
