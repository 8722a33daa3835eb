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

    def sample(self, rng=None):
        """gym.spaces.Box.sample for the bounded case (the action space); unbounded entries draw a standard normal."""
        rng = np.random if rng is None else rng
        lo, hi = np.asarray(self.low, dtype=np.float64), np.asarray(self.high, dtype=np.float64)
        bounded = np.isfinite(lo) & np.isfinite(hi)
        u = rng.uniform(size=self.shape)
        x = np.where(bounded, np.where(bounded, lo, 0.0) + u * np.where(bounded, hi - lo, 0.0), rng.standard_normal(self.shape))
        return x.astype(self.dtype)


@dataclass(frozen=True)
class Discrete:
    n: int
    shape: tuple = ()
    dtype: np.dtype = np.dtype(np.int64)

    def contains(self, x):
        return isinstance(x, (int, np.integer)) and 0 <= int(x) < self.n

    def sample(self, rng=None):
        rng = np.random if rng is None else rng
        return int(rng.randint(self.n))
