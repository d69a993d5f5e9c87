"""Minimal observation/action space types (the reference uses ``gym.spaces``, which is optional here).

``is_discrete`` accepts these classes and any gym / gymnasium ``Discrete`` by duck typing, so datasets
built with the reference's own spaces keep working."""
from __future__ import annotations

import numpy as np


class Space:
    pass


class Discrete(Space):
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.int64

    def contains(self, x):
        return 0 <= int(x) < self.n

    def __repr__(self):
        return f"Discrete({self.n})"


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low, self.high, self.dtype = low, high, dtype
        self.shape = tuple(shape) if shape is not None else np.shape(low)

    def __repr__(self):
        return f"Box({self.low}, {self.high}, {self.shape})"


def is_discrete(space) -> bool:
    return isinstance(space, Discrete) or (type(space).__name__ == "Discrete" and hasattr(space, "n"))
