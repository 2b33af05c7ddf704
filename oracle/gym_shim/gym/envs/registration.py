"""id -> 'module:Class' registry with `make`, as far as the reference's __init__ needs."""
import importlib

from gym import error


class EnvSpec(object):
    def __init__(self, id, entry_point=None, kwargs=None, **_ignored):
        self.id = id
        self.entry_point = entry_point
        self.kwargs = dict(kwargs or {})

    def make(self, **kwargs):
        if callable(self.entry_point):
            cls = self.entry_point
        else:
            mod_name, attr = self.entry_point.split(":")
            cls = getattr(importlib.import_module(mod_name), attr)
        kw = dict(self.kwargs)
        kw.update(kwargs)
        env = cls(**kw)
        try:
            env.spec = self
        except AttributeError:
            pass
        return env


class EnvRegistry(object):
    def __init__(self):
        self.env_specs = {}

    def register(self, id, **kwargs):
        self.env_specs[id] = EnvSpec(id, **kwargs)

    def spec(self, id):
        try:
            return self.env_specs[id]
        except KeyError:
            raise error.UnregisteredEnv("No registered env with id: %s" % id)

    def make(self, id, **kwargs):
        return self.spec(id).make(**kwargs)

    def all(self):
        return self.env_specs.values()


registry = EnvRegistry()


def register(id, **kwargs):
    return registry.register(id, **kwargs)


def make(id, **kwargs):
    return registry.make(id, **kwargs)


def spec(id):
    return registry.spec(id)
