"""SpaceEnv — ONE env of the batched engine behind the reference's own `gym.Env` interface.

The reference's callers (keyboard_agent.py, notebooks/run_agent.ipynb, BASELINE configs[0]) drive a single env:

    env = gym.make("GoalContinuous2P-v0"); env.seed(0)
    obs = env.reset()
    obs, reward, done, info = env.step(action)      # numpy in, numpy / python scalars out
    if done: obs = env.reset()

`SpaceEnv` runs that loop unchanged on the CUDA engine (a one-env `BatchedSpaceEnv` with `auto_reset=0`): same ids and
constructor kwargs (gym_space/__init__.py), `reset` / `step` / `seed` / `vector_field` / `render` / `close` with the
argument meaning and error behaviour of `SpaceshipEnv` (spaceship_env.py:59-111) under gym's `TimeLimit`
(`info["TimeLimit.truncated"]` only appears on the step that hits `max_episode_steps`; `step()` before the first `reset()` is
the AssertionError gym's TimeLimit raises; a finished env keeps stepping from its terminal state until the caller resets
it, as the bare reference env does).  This is the drop-in for the reference's
single-env call sites, not a fast path: one env per launch uses one lane of the GPU.
"""
import numpy as np
import torch

from .replay import Renderer
from .vector_env import BatchedSpaceEnv


class SpaceEnv:
    metadata = {"render.modes": ["rgb_array"]}

    def __init__(self, env_id, device="cuda:0", seed=0, **kwargs):
        kwargs.setdefault("auto_reset", 0)
        self._b = BatchedSpaceEnv(env_id, 1, device=device, seed=seed, copy=True, check_actions=True, **kwargs)
        self.observation_space = self._b.single_observation_space
        self.action_space = self._b.single_action_space
        self._started = False  # gym TimeLimit: "Cannot call env.step() before calling reset()"
        self._renderer = None
        self._last_action = None

    @property
    def unwrapped(self):
        return self

    def seed(self, seed=None):
        """spaceship_env.py:92-94: re-seeds the env's random stream, returns [seed]."""
        return self._b.seed(seed)

    def reset(self):
        """spaceship_env.py:59-66: task reset, returns the observation."""
        obs = self._b.reset()
        self._started = True
        if self._renderer is not None:
            self._sync_renderer(reset=True)
        return obs[0].cpu().numpy()

    def step(self, action):
        """spaceship_env.py:68-78 under gym's TimeLimit: (observation, reward, done, info)."""
        assert self._started, "Cannot call env.step() before calling reset()"
        disc = bool(self._b.cfg.discrete)
        if disc:
            a = np.zeros((1, 2), dtype=np.float32)
            assert np.ndim(action) == 0 and int(action) == action and 0 <= int(action) <= 5, "action outside Discrete(6)"
            a[0, 0] = int(action)
        else:
            a = np.asarray(action, dtype=np.float32).reshape(1, 2)
        self._last_action = a[0].copy()
        obs, rew, done, info = self._b.step(torch.as_tensor(a, device=self._b.device))
        d = bool(done[0])
        out_info = {}
        if bool(info["TimeLimit.truncated"][0]):
            out_info["TimeLimit.truncated"] = True
        return obs[0].cpu().numpy(), float(rew[0]), d, out_info

    def vector_field(self, action):
        """spaceship_env.py:96-100."""
        a = np.zeros((1, 2), dtype=np.float32)
        a[0, : np.size(action)] = np.asarray(action, dtype=np.float32).ravel()
        return self._b.vector_field(torch.as_tensor(a, device=self._b.device))[0].cpu().numpy()

    # ---- rendering (spaceship_env.py:80-90; software rasteriser of replay.py instead of the pyglet window) ----
    def _sync_renderer(self, reset=False):
        st = self._b.get_state()
        c = self._b.cfg
        if c.task == 0:
            planets = st["planets"][0].cpu().numpy().reshape(-1, 2)
            goal = st["goal"][0].cpu().numpy()
        else:
            planets, goal = np.zeros((1, 2)), None
        if reset:
            self._renderer.reset(planets, goal)
        else:
            self._renderer.move_goal(goal)
        return st["ship"][0].cpu().numpy()

    def render(self, mode="rgb_array"):
        if mode != "rgb_array":
            raise NotImplementedError("only mode='rgb_array' (no window system on the GPU box)")
        c = self._b.cfg
        if self._renderer is None:
            radius = c.planets_radius if c.task == 0 else c.kepler_planet_radius
            n_pl = c.n_planets if c.task == 0 else 1
            self._renderer = Renderer([radius] * n_pl, c.world_size)
            self._sync_renderer(reset=True)
        ship = self._sync_renderer()
        return self._renderer.render(ship[:3], self._last_action)

    def get_state(self):
        return self._b.get_state()

    def close(self):
        self._b.close()


def make_single(env_id, **kwargs):
    """`gym.make(env_id, **kwargs)` for one env on the engine (SpaceEnv)."""
    return SpaceEnv(env_id, **kwargs)
