"""Box / Discrete with the `contains` semantics of gym 0.21 (test infrastructure only)."""
import numpy as np


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        low = np.asarray(low)
        high = np.asarray(high)
        if shape is None:
            shape = low.shape
        self.shape = tuple(shape)
        self.low = np.broadcast_to(low, self.shape).astype(self.dtype)
        self.high = np.broadcast_to(high, self.shape).astype(self.dtype)

    def contains(self, x):
        if not isinstance(x, np.ndarray):
            x = np.asarray(x)
        return bool(
            np.can_cast(x.dtype, self.dtype)
            and x.shape == self.shape
            and np.all(x >= self.low)
            and np.all(x <= self.high)
        )

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(self.dtype)


class Discrete:
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)

    def contains(self, x):
        if isinstance(x, (int, np.integer)):
            return 0 <= int(x) < self.n
        if isinstance(x, np.ndarray) and x.shape == () and np.issubdtype(x.dtype, np.integer):
            return 0 <= int(x) < self.n
        return False

    def sample(self):
        return int(np.random.randint(self.n))
