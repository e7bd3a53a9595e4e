"""ctypes binding of the v1 oracle (oracle/futbol_v1_oracle.c).  TEST INFRASTRUCTURE ONLY.  Game logic pinned to traces of
the reference's own Python (tests/golden/v1_golden.npz), contact physics unpinned (see the header of the C file)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import v0 as _v0

MAX_N = 10
MAX_BODIES = 2 * MAX_N + 1
NSEG = 12
MAX_PAIRS = MAX_BODIES * (MAX_BODIES - 1) // 2 + MAX_BODIES * NSEG

CFG_DTYPE = np.dtype([("seed", np.uint64), ("n_players", np.int32), ("ep_limit", np.int32), ("damping_dt", np.float64),
                      ("bias_coef", np.float64), ("slop", np.float64), ("arith", np.int32), ("pad_", np.int32), ("form_x", np.float64, (2 * MAX_N,)), ("form_y", np.float64, (2 * MAX_N,))],
                     align=True)
ENV_DTYPE = np.dtype([("p", np.float64, (MAX_BODIES, 2)), ("v", np.float64, (MAX_BODIES, 2)), ("vb", np.float64, (MAX_BODIES, 2)),
                      ("jn", np.float64, (MAX_PAIRS,)), ("age", np.uint8, (MAX_PAIRS,)), ("t_total", np.uint64),
                      ("ep_step", np.int32), ("owner_side", np.int32), ("env_id", np.uint32), ("step_draws", np.uint32),
                      ("goals_left", np.int32), ("goals_right", np.int32), ("flags", np.int32), ("contacts", np.int32),
                      ("overflow", np.int32), ("pad_", np.int32)],
                     align=True)

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = _v0.lib()
        L.futbol_v1_oracle_env_bytes.restype = C.c_size_t
        L.futbol_v1_oracle_cfg_bytes.restype = C.c_size_t
        assert L.futbol_v1_oracle_env_bytes() == ENV_DTYPE.itemsize, (L.futbol_v1_oracle_env_bytes(), ENV_DTYPE.itemsize)
        assert L.futbol_v1_oracle_cfg_bytes() == CFG_DTYPE.itemsize
        L.futbol_v1_oracle_config.restype = C.c_int
        L.futbol_v1_oracle_config.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_double, C.c_int]
        L.futbol_v1_oracle_init.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.futbol_v1_oracle_reset.argtypes = [C.c_void_p, C.c_void_p]
        L.futbol_v1_oracle_obs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.futbol_v1_oracle_step.restype = C.c_int
        L.futbol_v1_oracle_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        L.futbol_v1_oracle_rollout.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 4
        L.futbol_v1_oracle_rollout_vs.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 4
        L.futbol_v1_oracle_team_actions.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleV1:
    """A batch of v1 oracle envs with global ids env_id0 .. env_id0+n-1 (constructed and reset, like Futbol())."""

    def __init__(self, n=1, seed=0, env_id0=0, number_of_player=2, total_time=30.0, arith=0):
        self.lib = lib()
        self.n, self.N = int(n), int(number_of_player)
        self.obs_dim = 4 + 8 * self.N
        self.cfg = np.zeros(1, dtype=CFG_DTYPE)
        if self.lib.futbol_v1_oracle_config(_ptr(self.cfg), int(seed), self.N, float(total_time), int(arith)) != 0:
            raise ValueError("number_of_player must be 1..10")
        self.envs = np.zeros(self.n, dtype=ENV_DTYPE)
        for i in range(self.n):
            self.lib.futbol_v1_oracle_init(_ptr(self.cfg), C.c_void_p(self.envs[i:i + 1].ctypes.data), env_id0 + i)

    def reset(self, idx=None):
        for i in (range(self.n) if idx is None else idx):
            self.lib.futbol_v1_oracle_reset(_ptr(self.cfg), C.c_void_p(self.envs[i:i + 1].ctypes.data))

    def obs(self, i=0):
        out = np.zeros(self.obs_dim)
        self.lib.futbol_v1_oracle_obs(_ptr(self.cfg), C.c_void_p(self.envs[i:i + 1].ctypes.data), _ptr(out))
        return out

    def step_one(self, i, action):
        a = np.ascontiguousarray(action, dtype=np.uint8).reshape(2 * self.N)
        r = C.c_double()
        d = self.lib.futbol_v1_oracle_step(_ptr(self.cfg), C.c_void_p(self.envs[i:i + 1].ctypes.data), _ptr(a), C.byref(r))
        return self.obs(i), r.value, bool(d)

    def rollout(self, steps, actions=None, autoreset=2, n_threads=1, record=True, right_actions=None):
        n, D = self.n, self.obs_dim
        out = {}
        if record:
            out = {"obs": np.zeros((steps, n, D)), "reward": np.zeros((steps, n)), "done": np.zeros((steps, n), np.uint8),
                   "flags": np.zeros((steps, n), np.uint8)}
        if actions is not None:
            actions = np.ascontiguousarray(actions, dtype=np.uint8).reshape(steps, n, 2 * self.N)
        if right_actions is not None:
            right_actions = np.ascontiguousarray(right_actions, dtype=np.uint8).reshape(steps, n, 2 * self.N)
        g = out.get
        self.lib.futbol_v1_oracle_rollout_vs(_ptr(self.cfg), _ptr(self.envs), n, int(steps), _ptr(actions), _ptr(right_actions), int(autoreset),
                                          int(n_threads), _ptr(g("obs")), _ptr(g("reward")), _ptr(g("done")), _ptr(g("flags")))
        return out


def team_actions(seed, env_id, stream, t, n_players):
    out = np.zeros(2 * n_players, dtype=np.uint8)
    lib().futbol_v1_oracle_team_actions(int(seed), int(env_id), int(stream), int(t), int(n_players), _ptr(out))
    return out
