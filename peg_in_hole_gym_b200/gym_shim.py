"""Minimal gym-compatible surface (gym / gymnasium are not installed in this image).

Provides ``spaces.Box``, ``Env``, ``register`` and ``make`` with the old 4-tuple step API the reference uses
(peg_in_hole_gym/__init__.py:1-11).  If a real ``gym`` or ``gymnasium`` is importable the ids are registered there
too, so ``gym.make('peg-in-hole-mp-v0', ...)`` keeps working unchanged."""
import importlib

import numpy as np


class Space:
    shape = None

    def sample(self):
        raise NotImplementedError

    def contains(self, x):
        raise NotImplementedError


class Box(Space):
    """gym.spaces.Box(low, high) with array bounds, float32 (envs/peg_in_hole.py:12-13)."""

    def __init__(self, low, high, shape=None, dtype=np.float32):
        low, high = np.asarray(low, dtype=dtype), np.asarray(high, dtype=dtype)
        if shape is not None:
            low, high = np.broadcast_to(low, shape).copy(), np.broadcast_to(high, shape).copy()
        self.low, self.high, self.shape, self.dtype = low, high, low.shape, np.dtype(dtype)
        self._rng = np.random.default_rng()

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)

    def sample(self):
        return self._rng.uniform(self.low, self.high).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


class spaces:  # namespace mirroring `from gym import spaces`
    Space = Space
    Box = Box

    class space:  # gym.spaces.space.Space (envs/utils.py:135)
        Space = Space


class Env:
    metadata = {"render.modes": ["human", "rgb_array"]}
    action_space = None
    observation_space = None

    def step(self, action):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    def render(self, mode="rgb_array"):
        raise NotImplementedError

    def close(self):
        pass


_REGISTRY = {}


def register(id, entry_point, **kwargs):
    _REGISTRY[id] = (entry_point, kwargs)
    for modname in ("gym", "gymnasium"):
        try:  # mirror into a real gym when present (never required)
            mod = importlib.import_module(modname + ".envs.registration")
            if id not in getattr(mod, "registry", {}):
                mod.register(id=id, entry_point=entry_point, **kwargs)
        except Exception:
            pass


def make(id, **kwargs):
    if id not in _REGISTRY:
        raise KeyError(f"No registered env with id: {id}")
    entry, defaults = _REGISTRY[id]
    if isinstance(entry, str):
        modname, cls = entry.split(":")
        entry = getattr(importlib.import_module(modname), cls)
    kw = dict(defaults)
    kw.update(kwargs)
    return entry(**kw)


# pybullet connection-mode enum values used as the `client` argument (README.md:38) [3P]
GUI, DIRECT = 1, 2
