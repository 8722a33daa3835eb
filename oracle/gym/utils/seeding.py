"""gym<=0.21 seeding: a legacy numpy RandomState (test infrastructure only).

The reference calls .uniform/.randint/.choice/.standard_normal on it
(gym_space/hexagonal_tiling.py:69-89, gym_space/envs/goal.py:140-143).
"""
import os
import numpy as np


def np_random(seed=None):
    if seed is None:
        seed = int.from_bytes(os.urandom(4), "little")
    rng = np.random.RandomState()
    rng.seed(int(seed) % (2 ** 32))
    return rng, seed
