"""ctypes binding of the C oracle (oracle/futbol_v0_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libfutbol_oracle.so")


class OracleV0Config(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("random_opp", C.c_int32), ("one_goal_end", C.c_int32),
                ("only_reward_goal", C.c_int32), ("arith", C.c_int32), ("rng_const", C.c_int32),
                ("pad_", C.c_int32), ("game_time", C.c_double),
                ("player_speed", C.c_double), ("shoot_speed", C.c_double)]


class OracleV0Env(C.Structure):
    _fields_ = [("obs", (C.c_double * 5) * 6), ("kick", (C.c_double * 2) * 4), ("time", C.c_double),
                ("draw_ctr", C.c_uint64), ("t_total", C.c_uint64), ("step_draws", C.c_uint32),
                ("normal_calls", C.c_uint32), ("env_id", C.c_uint32),
                ("owner", C.c_int32), ("last_owner", C.c_int32), ("ai_score", C.c_int32),
                ("opp_score", C.c_int32), ("flags", C.c_int32)]


ENV_DTYPE = np.dtype([("obs", np.float64, (6, 5)), ("kick", np.float64, (4, 2)), ("time", np.float64),
                      ("draw_ctr", np.uint64), ("t_total", np.uint64), ("step_draws", np.uint32),
                      ("normal_calls", np.uint32), ("env_id", np.uint32),
                      ("owner", np.int32), ("last_owner", np.int32), ("ai_score", np.int32),
                      ("opp_score", np.int32), ("flags", np.int32)], align=True)

_lib = None


def build(force=False):
    if force or not os.path.exists(_LIB_PATH) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_LIB_PATH)
            for f in ("futbol_v0_oracle.c", "futbol_v1_oracle.c", "Makefile")):
        import fcntl
        with open(_LIB_PATH + ".lock", "w") as lock:          # several processes may get here at once (torchrun, xdist)
            fcntl.flock(lock, fcntl.LOCK_EX)
            try:
                tmp = "libfutbol_oracle.so.tmp.%d" % os.getpid()
                subprocess.run(["make", "-C", _HERE, "-B", "CC=gcc", "TARGET=" + tmp, tmp], check=True, capture_output=True)
                os.replace(os.path.join(_HERE, tmp), _LIB_PATH)
            finally:
                fcntl.flock(lock, fcntl.LOCK_UN)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.futbol_v0_oracle_env_bytes.restype = C.c_size_t
        _lib.futbol_v0_oracle_cfg_bytes.restype = C.c_size_t
        assert _lib.futbol_v0_oracle_env_bytes() == ENV_DTYPE.itemsize == C.sizeof(OracleV0Env)
        assert _lib.futbol_v0_oracle_cfg_bytes() == C.sizeof(OracleV0Config)
        _lib.futbol_v0_oracle_step.restype = C.c_int
        _lib.futbol_v0_oracle_step.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_double)]
        _lib.futbol_v0_oracle_rollout_vs.restype = None
        _lib.futbol_v0_oracle_rollout_vs.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                                     C.c_int] + [C.c_void_p] * 9
        _lib.futbol_v0_oracle_rollout.restype = None
        _lib.futbol_v0_oracle_rollout.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                                  C.c_int] + [C.c_void_p] * 9
        _lib.futbol_oracle_action.restype = C.c_int
        _lib.futbol_oracle_action.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_int]
        _lib.futbol_oracle_philox.argtypes = [C.c_void_p] * 3
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class OracleV0:
    """A batch of oracle envs with global ids env_id0 .. env_id0+n-1."""

    def __init__(self, n=1, seed=0, env_id0=0, random_opp=True, one_goal_end=False, only_reward_goal=False,
                 game_time=40.0, player_speed=12.0, shoot_speed=20.0, arith=1, rng_const=False):
        self.lib = lib()
        self.n = int(n)
        self.cfg = OracleV0Config(seed, int(random_opp), int(one_goal_end), int(only_reward_goal), int(arith),
                                  int(rng_const), 0, float(game_time), float(player_speed), float(shoot_speed))
        self.envs = np.zeros(self.n, dtype=ENV_DTYPE)
        for i in range(self.n):
            self.lib.futbol_v0_oracle_init(C.c_void_p(self.envs[i:i + 1].ctypes.data), C.c_uint32(env_id0 + i))

    def reset(self, idx=None):
        for i in (range(self.n) if idx is None else idx):
            self.lib.futbol_v0_oracle_reset(C.c_void_p(self.envs[i:i + 1].ctypes.data))

    def step_one(self, i, action):
        r = C.c_double()
        d = self.lib.futbol_v0_oracle_step(C.byref(self.cfg), C.c_void_p(self.envs[i:i + 1].ctypes.data),
                                           int(action), C.byref(r))
        return self.envs["obs"][i].copy(), r.value, bool(d)

    def rollout(self, steps, actions=None, autoreset=1, n_threads=1, record=True, opp_actions=None):
        n = self.n
        out = {}
        if record:
            out = {"obs": np.zeros((steps, n, 30)), "reward": np.zeros((steps, n)),
                   "done": np.zeros((steps, n), np.uint8), "owner": np.zeros((steps, n), np.uint8),
                   "last_owner": np.zeros((steps, n), np.uint8), "ai_score": np.zeros((steps, n), np.int32),
                   "opp_score": np.zeros((steps, n), np.int32), "draws": np.zeros((steps, n), np.uint64),
                   "flags": np.zeros((steps, n), np.uint8)}
        if actions is not None:
            actions = np.ascontiguousarray(actions, dtype=np.uint8).reshape(steps, n)
        if opp_actions is not None:
            opp_actions = np.ascontiguousarray(opp_actions, dtype=np.uint8).reshape(steps, n)
        g = out.get
        self.lib.futbol_v0_oracle_rollout_vs(C.byref(self.cfg), _ptr(self.envs), n, int(steps), _ptr(actions), _ptr(opp_actions),
                                          int(autoreset), int(n_threads), _ptr(g("obs")), _ptr(g("reward")),
                                          _ptr(g("done")), _ptr(g("owner")), _ptr(g("last_owner")),
                                          _ptr(g("ai_score")), _ptr(g("opp_score")), _ptr(g("draws")),
                                          _ptr(g("flags")))
        return out


def action_for(seed, env_id, t, n_actions=16):
    return lib().futbol_oracle_action(seed, env_id, t, n_actions)
