"""Minimal stand-in for the old `gym` (<=0.21) surface the reference touches.

TEST INFRASTRUCTURE ONLY (see oracle/README.md): it lets the unmodified reference
package under /root/reference be imported in the authoring container, where neither
`gym` nor `gymnasium` is installed, so that golden vectors can be generated
(tests/golden/make_golden.py) and the C oracle can be pinned against the real
reference.  Nothing in the product package imports it.

Surface covered (reference call sites):
  gym.Env                              gym_space/envs/spaceship_env.py:15
  gym.spaces.Box / Discrete            gym_space/envs/spaceship_env.py:5,111,187,208
  gym.utils.seeding.np_random          gym_space/envs/spaceship_env.py:93, hexagonal_tiling.py:51
  gym.envs.registration.register/make  gym_space/__init__.py:1-146
"""
from . import spaces, utils, envs  # noqa: F401
from .envs.registration import register, make  # noqa: F401


class Env:
    """Attribute bag; the reference only subclasses it."""

    metadata = {}
    observation_space = None
    action_space = None


class Wrapper(Env):
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        return getattr(self.env, name)
