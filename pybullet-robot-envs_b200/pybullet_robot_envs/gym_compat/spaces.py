"""``gym.spaces.Box`` / ``Dict`` as used by the reference (panda_push_gym_env.py:95-101,
panda_push_gym_goal_env.py:52-56)."""
from collections import OrderedDict

import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None):
        self.shape = None if shape is None else tuple(shape)
        self.dtype = None if dtype is None else np.dtype(dtype)
        self.np_random = np.random.RandomState()

    def seed(self, seed=None):
        self.np_random.seed(seed)


class Box(Space):
    def __init__(self, low=None, high=None, shape=None, dtype=np.float32):
        dtype = np.dtype(dtype)
        if shape is None:
            low = np.asarray(low)
            high = np.asarray(high)
            assert low.shape == high.shape
            shape = low.shape
        else:
            assert np.isscalar(low) and np.isscalar(high)
            low = np.full(shape, low)
            high = np.full(shape, high)
        self.low = low.astype(dtype)
        self.high = high.astype(dtype)
        super().__init__(shape, dtype)

    def sample(self):
        return self.np_random.uniform(low=self.low, high=self.high, size=self.shape).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and np.all(x >= self.low) and np.all(x <= self.high)

    def __repr__(self):
        return "Box" + str(self.shape)

    def __eq__(self, other):
        return isinstance(other, Box) and np.allclose(self.low, other.low) and np.allclose(self.high, other.high)


class Dict(Space):
    def __init__(self, spaces=None, **kw):
        if spaces is None:
            spaces = kw
        if isinstance(spaces, dict) and not isinstance(spaces, OrderedDict):
            spaces = OrderedDict(sorted(list(spaces.items())))
        self.spaces = spaces
        super().__init__(None, None)

    def sample(self):
        return OrderedDict([(k, s.sample()) for k, s in self.spaces.items()])

    def __getitem__(self, k):
        return self.spaces[k]

    def __repr__(self):
        return "Dict(" + ", ".join(k + ":" + str(s) for k, s in self.spaces.items()) + ")"
