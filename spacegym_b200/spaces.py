"""Minimal Box / Discrete descriptions of the action and observation spaces.

Same fields as the gym spaces the reference builds (spaceship_env.py:102-111,183-214;
kepler.py:158-170); `gym` itself is not a dependency.
"""
from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Box:
    low: np.ndarray
    high: np.ndarray
    dtype: np.dtype = np.dtype(np.float32)

    @property
    def shape(self):
        return self.low.shape

    def contains(self, x):
        x = np.asarray(x)
        return bool(np.can_cast(x.dtype, self.dtype) and x.shape == self.shape
                    and np.all(x >= self.low) and np.all(x <= self.high))


@dataclass(frozen=True)
class Discrete:
    n: int
    shape: tuple = ()
    dtype: np.dtype = np.dtype(np.int64)

    def contains(self, x):
        return isinstance(x, (int, np.integer)) and 0 <= int(x) < self.n
