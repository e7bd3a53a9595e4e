"""Philox4x32-10 and the draw->value maps shared by every implementation.

TEST INFRASTRUCTURE ONLY (see oracle/README.md): imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs.
The product path (gym_futbol_b200/) never imports this module.

The reference (yc2454/gym-futbol) draws from the *unseeded* stdlib ``random``
and ``numpy.random`` generators (gym_futbol/envs/futbol_env.py:103,107,306,353,
367,416,459,641; gym_futbol/envs/easy_agent.py:83), so there is no reference
random stream to reproduce.  Parity is therefore defined on an injected
generator: this counter-based Philox stream.  The same specification is
implemented three times, independently:

  * here (pure Python ints / numpy)            -> drives the unmodified reference
  * oracle/futbol_v0_oracle.c  (plain C)       -> the CPU restatement
  * gym_futbol_b200/csrc/philox.cuh (CUDA)     -> the product

RNG specification
-----------------
Philox4x32-10 (Salmon et al., SC'11), multipliers 0xD2511F53 / 0xCD9E8D57,
Weyl constants 0x9E3779B9 / 0xBB67AE85.

  key      = (seed & 0xffffffff, seed >> 32)
  counter  = (t & 0xffffffff, ((t >> 32) & 0xffff) | (block << 16), global_env_id, stream)

``t`` is the env's total step index (number of ``step`` calls since creation; it is NOT
cleared by ``reset``), ``block`` a 16-bit block index inside that step.  A step's
randomness therefore depends only on (seed, env id, t): no RNG state is carried.

Streams: 0 = environment dynamics draws.  The reference consumes a variable number of
draws per step in a state-dependent order (SURVEY.md "RNG ledger"); draw ``j`` (0-based,
in the reference's call order, restarting at 0 every step) is word ``j & 3`` of block
``j >> 2``.  1 = synthetic AI actions (block 0, word 0).  2 = v1 opponent actions,
3 = v1 dynamics draws (same per-step sequential scheme as stream 0).

Draw -> value maps (``w`` = one 32-bit draw):

  random()        = (w >> 8) * 2**-24                        in [0, 1)
  randint(a, b)   = a + ((w * (b - a + 1)) >> 32)            a..b inclusive
  uniform(a, b)   = a + (b - a) * random()                   (CPython's formula)
  normal(mu,sd,n) = consumes NO sequential draws: the c-th normal() call of a step takes slot k
                    (k < n <= 16) from block 0x8000 + 8c + (k >> 1), words (w0, w1) =
                    (2(k&1), 2(k&1)+1) of that block (so a consumer that needs a single slot
                    evaluates one Philox block):
                    u1 = ((w0 >> 8) + 1) * 2**-24   in (0, 1]
                    u2 =  (w1 >> 8)      * 2**-24   in [0, 1)
                    z  = sqrt(-2 ln u1) * cos(2 pi u2);   value = mu + sd * z
  action(t)       = (w * n_actions) >> 32
"""
from __future__ import annotations

import math

import numpy as np

M0 = 0xD2511F53
M1 = 0xCD9E8D57
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = 0xFFFFFFFF

STREAM_DYNAMICS = 0
STREAM_ACTIONS = 1
STREAM_V1_OPP = 2
STREAM_V1_DYNAMICS = 3
NORMAL_BLOCK0 = 0x8000

TWO_PI = 6.283185307179586
INV_2_24 = 1.0 / 16777216.0


def philox4x32_10(ctr, key):
    """One Philox4x32-10 block.  ctr: 4 ints, key: 2 ints -> 4 ints."""
    c0, c1, c2, c3 = (int(x) & MASK for x in ctr)
    k0, k1 = (int(x) & MASK for x in key)
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


def philox4x32_10_np(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox: uint64 numpy arrays holding 32-bit values."""
    c0 = np.asarray(c0, dtype=np.uint64) & MASK
    c1 = np.asarray(c1, dtype=np.uint64) & MASK
    c2 = np.asarray(c2, dtype=np.uint64) & MASK
    c3 = np.asarray(c3, dtype=np.uint64) & MASK
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint64(int(k0) & MASK)
    k1 = np.uint64(int(k1) & MASK)
    m0, m1, mask, s32 = np.uint64(M0), np.uint64(M1), np.uint64(MASK), np.uint64(32)
    for _ in range(10):
        p0 = m0 * c0
        p1 = m1 * c2
        c0, c1, c2, c3 = ((p1 >> s32) ^ c1 ^ k0) & mask, p1 & mask, ((p0 >> s32) ^ c3 ^ k1) & mask, p0 & mask
        k0 = (k0 + np.uint64(W0)) & mask
        k1 = (k1 + np.uint64(W1)) & mask
    return c0, c1, c2, c3


def seed_key(seed: int):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return seed & MASK, seed >> 32


def u32_to_random(w: int) -> float:
    return (w >> 8) * INV_2_24


def u32_to_randint(w: int, a: int, b: int) -> int:
    return a + ((w * (b - a + 1)) >> 32)


def step_counter(t: int, block: int, env_id: int, stream: int):
    return (t & MASK, ((t >> 32) & 0xFFFF) | ((block & 0xFFFF) << 16), env_id & MASK, stream)


def action_for(seed: int, env_id: int, t: int, n_actions: int = 16) -> int:
    """Synthetic AI action of global env ``env_id`` at its total step ``t``."""
    w = philox4x32_10(step_counter(t, 0, env_id, STREAM_ACTIONS), seed_key(seed))[0]
    return (w * n_actions) >> 32


def actions_table(seed: int, env_ids, t0: int, steps: int, n_actions: int = 16) -> np.ndarray:
    """[steps, n_envs] uint8 table of synthetic actions (vectorised)."""
    env_ids = np.asarray(env_ids, dtype=np.uint64)
    t = np.arange(t0, t0 + steps, dtype=np.uint64)[:, None]
    k0, k1 = seed_key(seed)
    w, _, _, _ = philox4x32_10_np(t & np.uint64(MASK), (t >> np.uint64(32)) & np.uint64(0xFFFF), env_ids[None, :],
                                  np.uint64(STREAM_ACTIONS), k0, k1)
    return ((w * np.uint64(n_actions)) >> np.uint64(32)).astype(np.uint8)


class DrawStream:
    """Per-step sequential draw stream of one env (stream 0 or 3).

    The harness calls ``begin_step(t)`` before every ``env.step``; ``ctr`` is the number of
    draws consumed so far inside the current step, ``total`` since construction.
    """

    __slots__ = ("key", "env_id", "stream", "t", "ctr", "total", "normal_calls", "_blk_idx", "_blk", "log")

    def __init__(self, seed: int, env_id: int, stream: int = STREAM_DYNAMICS):
        self.key = seed_key(seed)
        self.env_id = int(env_id)
        self.stream = stream
        self.t = 0
        self.ctr = 0
        self.total = 0
        self.normal_calls = 0
        self._blk_idx = None
        self._blk = None
        self.log = None  # set to a list to record (kind, value) per call

    def begin_step(self, t: int):
        self.t = int(t)
        self.ctr = 0
        self.normal_calls = 0
        self._blk_idx = None

    def word_at(self, j: int) -> int:
        b = j >> 2
        if b != self._blk_idx:
            self._blk = philox4x32_10(step_counter(self.t, b, self.env_id, self.stream), self.key)
            self._blk_idx = b
        return self._blk[j & 3]

    def next_u32(self) -> int:
        w = self.word_at(self.ctr)
        self.ctr += 1
        self.total += 1
        return w

    # --- the four entry points the reference calls -------------------------------
    def random(self) -> float:
        v = u32_to_random(self.next_u32())
        if self.log is not None:
            self.log.append(("random", v))
        return v

    def randint(self, a, b) -> int:
        # the reference passes floats (futbol_env.py:306: randint(32.0, 36.0)); CPython <= 3.11
        # accepted integral floats, so coerce exactly like it did.
        a, b = int(a), int(b)
        v = u32_to_randint(self.next_u32(), a, b)
        if self.log is not None:
            self.log.append(("randint", v))
        return v

    def uniform(self, a, b) -> float:
        v = a + (b - a) * u32_to_random(self.next_u32())
        if self.log is not None:
            self.log.append(("uniform", v))
        return v

    def normal(self, mu, sd, n):
        n = int(n)
        assert n <= 16
        out = np.empty(n, dtype=np.float64)
        base = NORMAL_BLOCK0 + 8 * self.normal_calls
        self.normal_calls += 1
        for k in range(n):
            blk = philox4x32_10(step_counter(self.t, base + (k >> 1), self.env_id, self.stream), self.key)
            w0, w1 = blk[2 * (k & 1)], blk[2 * (k & 1) + 1]
            u1 = ((w0 >> 8) + 1) * INV_2_24
            u2 = (w1 >> 8) * INV_2_24
            z = math.sqrt(-2.0 * math.log(u1)) * math.cos(TWO_PI * u2)
            out[k] = mu + sd * z
        if self.log is not None:
            self.log.append(("normal", out.copy()))
        return out
