"""Stand-in for gym==0.19 surface used by the reference: Env, Wrapper, spaces."""
from . import spaces  # noqa: F401


class Env:
    pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)
