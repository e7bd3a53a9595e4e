"""The kernel's step logic (gym_futbol_b200/csrc/v0_step.cuh, the DEVICE header) compiled for the host with
the CUDA intrinsics shimmed (tests/host_shim/v0_step_host.cpp) and compared BIT for bit with the oracle
in kernel-arithmetic mode.  This is a development aid for machines without a GPU: it checks the logic of
the branch-lean step (predicated player turns, deferred kick, squared-distance thresholds); the GPU parity
tests (test_v0_gpu.py) remain the gate for the compiled sm_100a code.  Nothing in the package uses it.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

from gym_futbol_b200._lib import V0_ENV_STATE
from oracle import philox
from oracle.v0 import OracleV0

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_shim", "v0_step_host.cpp")
SO = os.path.join(HERE, "host_shim", "_v0_step_host.so")
CSRC = os.path.join(os.path.dirname(HERE), "gym_futbol_b200", "csrc")


@pytest.fixture(scope="module")
def host_lib():
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("v0_step.cuh", "philox.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                        "-o", SO, SRC], check=True)
    lib = C.CDLL(SO)
    lib.host_v0_rollout.restype = None
    lib.host_v0_rollout.argtypes = [C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_double, C.c_double, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def ep_limit(game_time):
    t, k = 0.0, 0
    while not t >= game_time:
        t += 0.1
        k += 1
    return k


def sqrt_less_than_bound(r):
    """Largest double s with sqrt(s) < r (mirror of capi.cu sqrt_less_than_bound)."""
    if not r > 0.0:
        return -1.0
    s = r * r
    while math.sqrt(s) >= r:
        s = math.nextafter(s, -math.inf)
    while math.sqrt(math.nextafter(s, math.inf)) < r:
        s = math.nextafter(s, math.inf)
    return s


def kickoff_records(n):
    rec = np.zeros(n, dtype=V0_ENV_STATE)
    rec["rows"][:] = np.array([[43.5, 39, 0, 0, 0], [43.5, 29, 0, 0, 0], [61.5, 39, 0, 0, 0], [61.5, 29, 0, 0, 0],
                               [52.5, 34, 0, 0, 0]])
    rec["owner"] = 4
    rec["last_owner"] = 4
    return rec


def run_host(lib, n, steps, seed, off, random_opp, game_time=40.0, player_speed=12.0, one_goal_end=False,
             only_reward_goal=False, actions=None):
    rec = kickoff_records(n)
    obs = np.zeros((steps, n, 30))
    rew = np.zeros((steps, n))
    done = np.zeros((steps, n), np.uint8)
    flags = np.zeros((steps, n), np.uint8)
    lib.host_v0_rollout(seed, off, int(random_opp), int(one_goal_end), int(only_reward_goal), 1, ep_limit(game_time), 20,
                        player_speed, sqrt_less_than_bound(0.1 * player_speed), rec.ctypes.data, n, steps,
                        actions.ctypes.data, obs.ctypes.data, rew.ctypes.data, done.ctypes.data, flags.ctypes.data)
    return rec, obs, rew, done, flags


@pytest.mark.parametrize("random_opp", [True, False])
@pytest.mark.parametrize("flags", [dict(), dict(one_goal_end=True), dict(only_reward_goal=True)])
def test_device_step_logic_bit_exact_on_host(host_lib, random_opp, flags):
    n, steps, seed, off = 2048, 600, 5, 9000
    acts = philox.actions_table(seed, np.arange(off, off + n), 0, steps)
    orc = OracleV0(n, seed=seed, env_id0=off, random_opp=random_opp, game_time=25.0, arith=0, **flags)
    want = orc.rollout(steps, actions=acts, autoreset=2, n_threads=8)
    rec, obs, rew, done, fl = run_host(host_lib, n, steps, seed, off, random_opp, game_time=25.0, actions=acts, **flags)
    assert np.array_equal(done, want["done"])
    assert np.array_equal(fl & 3, want["flags"] & 3)     # goal / out-of-field bits (the done bit is compared via `done`)
    assert np.array_equal(rew, want["reward"])
    assert np.array_equal(obs, want["obs"])                 # bit-exact float64
    assert np.array_equal(rec["owner"], orc.envs["owner"].astype(np.uint8))
    assert np.array_equal(rec["last_owner"], orc.envs["last_owner"].astype(np.uint8))
    assert np.array_equal(rec["ai_score"], orc.envs["ai_score"]) and np.array_equal(rec["opp_score"], orc.envs["opp_score"])
    assert np.array_equal(rec["rows"].reshape(n, 25), orc.envs["obs"][:, :5].reshape(n, 25))
    assert want["done"].sum() > 0 and (want["flags"] & 1).sum() > 0 and (want["flags"] & 2).sum() > 0


def test_other_player_speed(host_lib):
    """`reach_sq_max` (the squared-distance form of `< 0.1 * player_speed`) for a non-default speed."""
    n, steps, seed = 1024, 300, 2
    acts = philox.actions_table(seed, np.arange(n), 0, steps)
    orc = OracleV0(n, seed=seed, random_opp=False, player_speed=9.5, arith=0)
    want = orc.rollout(steps, actions=acts, autoreset=2, n_threads=8)
    rec, obs, rew, done, fl = run_host(host_lib, n, steps, seed, 0, False, player_speed=9.5, actions=acts)
    assert np.array_equal(obs, want["obs"]) and np.array_equal(rew, want["reward"]) and np.array_equal(done, want["done"])


def test_at_most_one_kick_per_step():
    """The deferred kick relies on at most one normal() call per step (v0_step.cuh PendingShot)."""
    n, steps = 4096, 400
    for random_opp in (True, False):
        orc = OracleV0(n, seed=3, random_opp=random_opp, arith=0)
        seen = 0
        for _ in range(steps // 50):
            orc.rollout(50, actions=None, autoreset=2, n_threads=8, record=False)
            seen = max(seen, int(orc.envs["normal_calls"].max()))
        assert seen <= 1
