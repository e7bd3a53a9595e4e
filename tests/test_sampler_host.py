"""The row rule of futbol_sample_actions (gym_futbol_b200/csrc/sampler.cuh, the DEVICE header) compiled for the host and
checked against a float64 restatement with the same Philox uniforms (oracle/philox.py) -- the CPU counterpart of
tests/test_glue_gpu.py::test_sample_actions_follows_its_rule, which gates the compiled sm_100a kernel."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import philox

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_shim", "sampler_host.cpp")
SO = os.path.join(HERE, "host_shim", "_sampler_host.so")
CSRC = os.path.join(os.path.dirname(HERE), "gym_futbol_b200", "csrc")


@pytest.fixture(scope="module")
def host_lib():
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("sampler.cuh", "philox.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-o", SO, SRC], check=True)
    lib = C.CDLL(SO)
    lib.host_sample_actions.restype = None
    lib.host_sample_actions.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]
    return lib


def uniforms(seed, t, n):
    k0, k1 = philox.seed_key(seed)
    w = philox.philox4x32_10_np(np.full(n, t & 0xFFFFFFFF, np.uint64), np.full(n, t >> 32, np.uint64), np.arange(n, dtype=np.uint64),
                                np.full(n, 4, np.uint64), k0, k1)[0]
    return (w >> np.uint64(8)).astype(np.float64) / 16777216.0


def sample(lib, logits, seed, t):
    logits = np.ascontiguousarray(logits, np.float32)
    n, A = logits.shape
    act, logp = np.zeros(n, np.uint8), np.zeros(n, np.float32)
    lib.host_sample_actions(logits.ctypes.data, n, A, seed, t, act.ctypes.data, logp.ctypes.data)
    return act.astype(np.int64), logp


@pytest.mark.parametrize("A", [1, 5, 16, 32])
def test_row_rule(host_lib, A):
    n, seed, t = 20000, 77, (5 << 32) + 123
    rng = np.random.default_rng(A)
    logits = (rng.standard_normal((n, A)) * 3.0).astype(np.float32)
    if A > 3:
        logits[:50, 3] = -200.0                                 # a term that underflows to zero is never picked
        logits[50:60] = 0.0
    a, logp = sample(host_lib, logits, seed, t)
    l = logits.astype(np.float64)
    e = np.exp(l - l.max(1, keepdims=True))
    cum, tot = np.cumsum(e, 1), e.sum(1)
    target = uniforms(seed, t, n) * tot
    rows = np.arange(n)
    lo = np.where(a > 0, cum[rows, np.maximum(a - 1, 0)], 0.0)
    eps = 1e-5 * tot                                            # the rule sums in float32
    assert np.all(lo <= target + eps) and np.all(cum[rows, a] > target - eps)
    assert np.mean(np.argmax(cum > target[:, None], 1) == a) > 0.999
    if A > 3:
        assert not np.any(a[:50] == 3)
    want = l[rows, a] - l.max(1) - np.log(tot)
    assert np.allclose(logp, want, atol=2e-5, rtol=0)
    a2, _ = sample(host_lib, logits, seed, t + 1)
    assert A == 1 or not np.array_equal(a, a2)


def test_distribution_and_last_nonzero_fallback(host_lib):
    row = np.array([0.3, -1.0, 2.0, 0.0, -4.0, 1.5, 0.7, -0.2, 1.1, -2.5, 0.0, 0.9, -0.6, 2.2, -1.7, 0.4], np.float32)
    n = 200000
    a, _ = sample(host_lib, np.tile(row, (n, 1)), 5, 9)
    p = np.exp(row.astype(np.float64)); p /= p.sum()
    freq = np.bincount(a, minlength=16) / n
    assert np.all(np.abs(freq - p) < 5 * np.sqrt(p * (1 - p) / n))
    # only one action has a non-zero term: it is picked whatever the uniform is
    lone = np.full((1000, 16), -1000.0, np.float32)
    lone[:, 11] = 3.0
    a, logp = sample(host_lib, lone, 1, 2)
    assert np.all(a == 11) and np.allclose(logp, 0.0, atol=1e-6)
