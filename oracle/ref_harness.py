"""Runs the UNMODIFIED reference v0 ``FutbolEnv`` with an injected RNG.

TEST INFRASTRUCTURE ONLY.  Used (a) in the build container to generate the golden
fixtures under tests/golden/ (tests/golden/make_golden_v0.py), and (b) by
``bench.py --impl reference`` to time the reference's own CPU path when a copy of
the reference package is available (``/root/reference`` in the build container,
the git-ignored pip ``--target`` install ``baseline/_ref`` elsewhere).
The product path never imports this file.

What is injected and why:
  * ``gym`` and ``matplotlib.pyplot`` are not installed here: minimal stand-ins are
    placed in ``sys.modules`` *before* the reference is imported.  Only what the
    reference touches at import/constructor time exists (gym.Env, gym.spaces.*,
    gym.envs.registration.register).
  * the module-level names ``random`` (futbol_env.py:12, easy_agent.py:4) and ``np``
    (futbol_env.py:9) are rebound to objects that forward every draw to the
    ``oracle.philox.DrawStream`` of the env being stepped.  This is mandatory on
    Python >= 3.12 anyway: un-patched, futbol_env.py:306 raises TypeError
    (``random.randint`` with float bounds).
No reference source is modified or copied.
"""
from __future__ import annotations

import importlib
import os
import random as _stdlib_random
import sys
import types

import numpy as _np

from . import philox

_REF_CANDIDATES = (
    os.environ.get("FUTBOL_REFERENCE_ROOT", ""),
    "/root/reference",
    os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref"),
)


def find_reference_root():
    for root in _REF_CANDIDATES:
        if root and os.path.isfile(os.path.join(root, "gym_futbol", "envs", "futbol_env.py")):
            return root
    return None


# --------------------------------------------------------------------------- stubs
def _install_stubs():
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")

        class Env:  # gym.Env of the 0.17 era: no behaviour the reference relies on
            metadata = {}

        class _Space:
            def __init__(self, **kw):
                self.__dict__.update(kw)

        class Discrete(_Space):
            def __init__(self, n):
                super().__init__(n=n)

        class Box(_Space):
            def __init__(self, low, high, shape=None, dtype=_np.float32):
                low = _np.asarray(low)
                super().__init__(low=low, high=_np.asarray(high), shape=shape if shape is not None else low.shape,
                                 dtype=dtype)

        class Tuple(_Space):
            def __init__(self, spaces):
                super().__init__(spaces=tuple(spaces))

        class MultiDiscrete(_Space):
            def __init__(self, nvec):
                super().__init__(nvec=_np.asarray(nvec), sampler=None)

            def sample(self):        # envs_v1/futbol_env.py:307; the v1 harness installs the draw stream here
                if self.sampler is None:
                    raise RuntimeError("gym stand-in: MultiDiscrete.sample() needs an injected sampler")
                return self.sampler()

        spaces = types.ModuleType("gym.spaces")
        spaces.Discrete, spaces.Box, spaces.Tuple, spaces.MultiDiscrete = Discrete, Box, Tuple, MultiDiscrete
        error = types.ModuleType("gym.error")
        utils = types.ModuleType("gym.utils")
        seeding = types.ModuleType("gym.utils.seeding")          # envs_v1/futbol_env.py:3 imports it, never uses it
        utils.seeding = seeding
        envs = types.ModuleType("gym.envs")
        registration = types.ModuleType("gym.envs.registration")
        registration.registry = {}

        def register(id, entry_point=None, kwargs=None, **_):
            registration.registry[id] = (entry_point, kwargs or {})

        registration.register = register
        envs.registration = registration
        gym.Env, gym.spaces, gym.error, gym.utils, gym.envs = Env, spaces, error, utils, envs
        sys.modules.update({"gym": gym, "gym.spaces": spaces, "gym.error": error, "gym.utils": utils, "gym.utils.seeding": seeding,
                            "gym.envs": envs, "gym.envs.registration": registration})
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt})


# ------------------------------------------------------------------- RNG injection
class _RandomShim:
    """Stands in for the stdlib ``random`` module inside the reference."""

    def __init__(self):
        self.stream = None

    def randint(self, a, b):
        return self.stream.randint(a, b)

    def random(self):
        return self.stream.random()

    def uniform(self, a, b):
        return self.stream.uniform(a, b)


class _NpRandomShim:
    def __init__(self, shim):
        self._shim = shim

    def normal(self, mu, sd, n):
        return self._shim.stream.normal(mu, sd, n)


class _NpProxy:
    """numpy, except ``.random`` (futbol_env.py:103 is the only use)."""

    def __init__(self, shim):
        self.random = _NpRandomShim(shim)

    def __getattr__(self, name):
        return getattr(_np, name)


class _MTStream:
    """stdlib-Mersenne-Twister stream with the DrawStream interface (timing runs only)."""

    total = 0

    def __init__(self, seed):
        self._r = _stdlib_random.Random(seed)
        self._n = _np.random.RandomState(seed & 0x7FFFFFFF)

    def begin_step(self, t):
        pass

    def randint(self, a, b):
        return self._r.randint(int(a), int(b))

    def random(self):
        return self._r.random()

    def uniform(self, a, b):
        return self._r.uniform(a, b)

    def normal(self, mu, sd, n):
        return self._n.normal(mu, sd, n)


class _ConstStream:
    """The constant RNG of SURVEY.md Appendix A (RNG-free known-answer trace)."""

    ctr = 0
    total = 0

    def begin_step(self, t):
        pass

    def randint(self, a, b):
        return int(a)

    def random(self):
        return 0.5

    def uniform(self, a, b):
        return (a + b) / 2

    def normal(self, mu, sd, n):
        return _np.zeros(n)


_loaded = None


def load_reference():
    """Import the reference v0 modules (once) and inject the RNG shim."""
    global _loaded
    if _loaded is not None:
        return _loaded
    root = find_reference_root()
    if root is None:
        raise RuntimeError("reference package not found (looked in %s)" % (_REF_CANDIDATES,))
    _install_stubs()
    if root not in sys.path:
        sys.path.insert(0, root)
    fe = importlib.import_module("gym_futbol.envs.futbol_env")
    ea = importlib.import_module("gym_futbol.envs.easy_agent")
    shim = _RandomShim()
    fe.random = shim
    ea.random = shim
    fe.np = _NpProxy(shim)
    _loaded = (fe, ea, shim)
    return _loaded


class RefEnvV0:
    """One reference ``FutbolEnv`` bound to its own draw stream.

    Protocol fixed by SURVEY.md Q1: construct, ``reset()``, then ``step()``.
    """

    OWNER = {"AI_1": 0, "AI_2": 1, "OPP_1": 2, "OPP_2": 3, "NOONE": 4}

    def __init__(self, seed=0, env_id=0, random_opp=True, rng="philox", **kwargs):
        self.fe, _, self.shim = load_reference()
        if rng == "philox":
            self.stream = philox.DrawStream(seed, env_id)
        elif rng == "const":
            self.stream = _ConstStream()
        else:
            self.stream = _MTStream(seed * 1000003 + env_id)
        self.shim.stream = self.stream
        self.env = self.fe.FutbolEnv(random_opp=random_opp, **kwargs)
        self.obs = self.env.reset()
        self.t_total = 0   # steps since construction: the Philox step index (not cleared by reset)

    def reset(self):
        self.shim.stream = self.stream
        self.obs = self.env.reset()
        return self.obs

    def step(self, action):
        self.shim.stream = self.stream
        self.stream.begin_step(self.t_total)
        self.t_total += 1
        self.obs, reward, done, info = self.env.step(action)
        return self.obs, reward, done, info

    @property
    def owner(self):
        return self.env.ball_owner.value

    @property
    def last_owner(self):
        return self.env.last_ball_owner.value


def rollout_v0(seed, env_id, steps, random_opp, actions=None, reset_on_done=True, rng="philox", **kwargs):
    """Step one reference env ``steps`` times; returns a dict of per-step arrays.

    actions: optional sequence of ints (len steps); default = Philox action stream.
    Every integer output and the full (6,5) float64 obs are recorded AFTER each step and
    BEFORE the harness-level reset that follows a ``done``.
    """
    env = RefEnvV0(seed=seed, env_id=env_id, random_opp=random_opp, rng=rng, **kwargs)
    out = {
        "action": _np.zeros(steps, _np.uint8), "obs": _np.zeros((steps, 6, 5), _np.float64),
        "reward": _np.zeros(steps, _np.float64), "done": _np.zeros(steps, _np.uint8),
        "owner": _np.zeros(steps, _np.uint8), "last_owner": _np.zeros(steps, _np.uint8),
        "ai_score": _np.zeros(steps, _np.int32), "opp_score": _np.zeros(steps, _np.int32),
        "draws": _np.zeros(steps, _np.int64),
    }
    for t in range(steps):
        a = int(actions[t]) if actions is not None else philox.action_for(seed, env_id, t)
        obs, r, d, _ = env.step(a)
        out["action"][t] = a
        out["obs"][t] = obs
        out["reward"][t] = r
        out["done"][t] = d
        out["owner"][t] = env.owner
        out["last_owner"][t] = env.last_owner
        out["ai_score"][t] = env.env.ai_score
        out["opp_score"][t] = env.env.opp_score
        out["draws"][t] = env.stream.total
        if d and reset_on_done:
            env.reset()
    return out
