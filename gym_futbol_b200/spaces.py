"""Observation/action space objects.

Uses ``gym.spaces`` (or ``gymnasium.spaces``) when one of them is installed, so the drop-in classes
expose the very same space types as the reference (gym_futbol/envs/futbol_env.py:156-179,
gym_futbol/envs_v1/futbol_env.py:78-91); otherwise minimal stand-ins with the attributes policies
read (``n``, ``nvec``, ``shape``, ``dtype``, ``low``, ``high``, ``sample()``, ``contains()``).
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - depends on the environment
    from gym import spaces as _sp
    Discrete, Box, Tuple, MultiDiscrete = _sp.Discrete, _sp.Box, _sp.Tuple, _sp.MultiDiscrete
    BACKEND = "gym"
except Exception:  # noqa: BLE001
    try:  # pragma: no cover
        from gymnasium import spaces as _sp
        Discrete, Box, Tuple, MultiDiscrete = _sp.Discrete, _sp.Box, _sp.Tuple, _sp.MultiDiscrete
        BACKEND = "gymnasium"
    except Exception:  # noqa: BLE001
        BACKEND = "builtin"

        class _Space:
            shape = ()
            dtype = None
            _rng = np.random.RandomState()

            def seed(self, seed=None):
                self._rng = np.random.RandomState(seed)

        class Discrete(_Space):
            def __init__(self, n):
                self.n, self.shape, self.dtype = int(n), (), np.int64

            def sample(self):
                return int(self._rng.randint(self.n))

            def contains(self, x):
                return isinstance(x, (int, np.integer)) and 0 <= int(x) < self.n

            def __repr__(self):
                return "Discrete(%d)" % self.n

        class MultiDiscrete(_Space):
            def __init__(self, nvec):
                self.nvec = np.asarray(nvec, dtype=np.int64)
                self.shape, self.dtype = self.nvec.shape, np.int64

            def sample(self):
                return (self._rng.random_sample(self.nvec.shape) * self.nvec).astype(np.int64)

            def contains(self, x):
                x = np.asarray(x)
                return x.shape == self.nvec.shape and bool(((x >= 0) & (x < self.nvec)).all())

            def __repr__(self):
                return "MultiDiscrete(%s)" % self.nvec.tolist()

        class Box(_Space):
            def __init__(self, low, high, shape=None, dtype=np.float32):
                if shape is None:
                    low, high = np.asarray(low), np.asarray(high)
                    shape = low.shape
                self.low = np.broadcast_to(np.asarray(low, dtype=dtype), shape).copy()
                self.high = np.broadcast_to(np.asarray(high, dtype=dtype), shape).copy()
                self.shape, self.dtype = tuple(shape), np.dtype(dtype)

            def sample(self):
                return self._rng.uniform(self.low, self.high).astype(self.dtype)

            def contains(self, x):
                x = np.asarray(x)
                return x.shape == self.shape and bool(((x >= self.low) & (x <= self.high)).all())

            def __repr__(self):
                return "Box%s" % (self.shape,)

        class Tuple(_Space):
            def __init__(self, spaces):
                self.spaces = tuple(spaces)

            def sample(self):
                return tuple(s.sample() for s in self.spaces)

            def contains(self, x):
                return len(x) == len(self.spaces) and all(s.contains(v) for s, v in zip(self.spaces, x))

            def __repr__(self):
                return "Tuple%s" % (self.spaces,)
