"""register()/make() + TimeLimit with gym 0.21 semantics (test infrastructure only)."""
import importlib

registry = {}


def register(id, entry_point=None, max_episode_steps=None, kwargs=None, **_):
    registry[id] = dict(entry_point=entry_point, max_episode_steps=max_episode_steps, kwargs=dict(kwargs or {}))


class TimeLimit:
    """gym/wrappers/time_limit.py (0.21): elapsed+=1; truncate at max_episode_steps."""

    def __init__(self, env, max_episode_steps):
        self.env = env
        self._max_episode_steps = max_episode_steps
        self._elapsed_steps = None

    def __getattr__(self, name):
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env

    def step(self, action):
        obs, reward, done, info = self.env.step(action)
        self._elapsed_steps += 1
        if self._elapsed_steps >= self._max_episode_steps:
            info["TimeLimit.truncated"] = not done
            done = True
        return obs, reward, done, info

    def reset(self, **kwargs):
        self._elapsed_steps = 0
        return self.env.reset(**kwargs)


def make(id, **overrides):
    spec = registry[id]
    mod_name, cls_name = spec["entry_point"].split(":")
    cls = getattr(importlib.import_module(mod_name), cls_name)
    kwargs = dict(spec["kwargs"])
    kwargs.update(overrides)
    env = cls(**kwargs)
    if spec["max_episode_steps"] is not None:
        env = TimeLimit(env, spec["max_episode_steps"])
    return env
