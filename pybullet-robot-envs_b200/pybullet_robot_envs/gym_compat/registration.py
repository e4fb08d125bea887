"""``gym.envs.registration.register`` / ``gym.make`` for the ids of reference __init__.py:7-80."""
import importlib

from .core import TimeLimit


class EnvSpec:
    def __init__(self, id, entry_point=None, max_episode_steps=None, kwargs=None):
        self.id = id
        self.entry_point = entry_point
        self.max_episode_steps = max_episode_steps
        self._kwargs = {} if kwargs is None else kwargs

    def make(self, **kwargs):
        kw = dict(self._kwargs)
        kw.update(kwargs)
        if callable(self.entry_point):
            env = self.entry_point(**kw)
        else:
            mod_name, attr = self.entry_point.split(":")
            env = getattr(importlib.import_module(mod_name), attr)(**kw)
        env.spec = self
        if self.max_episode_steps is not None:
            env = TimeLimit(env, max_episode_steps=self.max_episode_steps)
        return env


class Registry:
    def __init__(self):
        self.env_specs = {}

    def register(self, id, **kw):
        if id in self.env_specs:
            raise ValueError("Cannot re-register id: {}".format(id))
        self.env_specs[id] = EnvSpec(id, **kw)

    def spec(self, id):
        if id not in self.env_specs:
            raise KeyError("No registered env with id: {}".format(id))
        return self.env_specs[id]

    def make(self, id, **kwargs):
        return self.spec(id).make(**kwargs)

    def all(self):
        return self.env_specs.values()


registry = Registry()


def register(id, **kw):
    return registry.register(id, **kw)


def make(id, **kwargs):
    return registry.make(id, **kwargs)


def spec(id):
    return registry.spec(id)
