"""The v1 kernel's step logic (gym_futbol_b200/csrc/v1_step.cuh, the DEVICE header) compiled for the host and
compared BIT for bit with the v1 oracle (same specification, independent code).  Development aid for machines
without a GPU; the GPU parity tests (test_v1_gpu.py) gate the compiled sm_100a code."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle.v1 import OracleV1

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_shim", "v1_step_host.cpp")
SO = os.path.join(HERE, "host_shim", "_v1_step_host.so")
CSRC = os.path.join(os.path.dirname(HERE), "gym_futbol_b200", "csrc")


@pytest.fixture(scope="module")
def host_lib():
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("v1_step.cuh", "philox.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                        "-o", SO, SRC], check=True)
    lib = C.CDLL(SO)
    lib.host_v1_rollout.restype = None
    lib.host_v1_rollout.argtypes = [C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                                    C.c_int, C.c_int] + [C.c_void_p] * 6
    return lib


def run_host(lib, orc_cfg, n, steps, seed, off, N, actions=None):
    D = 4 + 8 * N
    obs, rew = np.zeros((steps, n, D)), np.zeros((steps, n))
    done, flags, contacts = np.zeros((steps, n), np.uint8), np.zeros((steps, n), np.uint8), np.zeros((steps, n), np.int32)
    cfg = orc_cfg[0]
    fx, fy = np.ascontiguousarray(cfg["form_x"]), np.ascontiguousarray(cfg["form_y"])
    lib.host_v1_rollout(seed, off, N, int(cfg["ep_limit"]), float(cfg["damping_dt"]), float(cfg["bias_coef"]), fx.ctypes.data,
                        fy.ctypes.data, n, steps, None if actions is None else actions.ctypes.data, obs.ctypes.data,
                        rew.ctypes.data, done.ctypes.data, flags.ctypes.data, contacts.ctypes.data)
    return obs, rew, done, flags, contacts


@pytest.mark.parametrize("N,n,steps", [(1, 256, 700), (2, 1024, 700), (3, 256, 400), (5, 512, 400), (7, 128, 350), (10, 128, 350)])
def test_v1_device_step_logic_bit_exact_on_host(host_lib, N, n, steps):
    seed, off = 11, 4000
    orc = OracleV1(n, seed=seed, env_id0=off, number_of_player=N)
    want = orc.rollout(steps, actions=None, autoreset=2, n_threads=8)
    obs, rew, done, flags, contacts = run_host(host_lib, orc.cfg, n, steps, seed, off, N)
    assert np.array_equal(done, want["done"]) and np.array_equal(flags, want["flags"])
    assert np.array_equal(rew, want["reward"])
    assert np.array_equal(obs, want["obs"])                  # bit-exact float64
    assert contacts.sum() > 0 and (flags & 1).sum() > 0 and orc.envs["overflow"].sum() == 0


def test_v1_given_left_actions(host_lib):
    N, n, steps, seed = 2, 300, 320, 5
    rng = np.random.default_rng(0)
    acts = rng.integers(0, 5, (steps, n, 2 * N), dtype=np.uint8)
    orc = OracleV1(n, seed=seed, number_of_player=N)
    want = orc.rollout(steps, actions=acts, autoreset=2, n_threads=8)
    obs, rew, done, flags, _ = run_host(host_lib, orc.cfg, n, steps, seed, 0, N, actions=acts)
    assert np.array_equal(obs, want["obs"]) and np.array_equal(rew, want["reward"]) and np.array_equal(done, want["done"])
