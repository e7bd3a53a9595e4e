#!/usr/bin/env python
"""Generates tests/golden/v0_golden.npz by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_v0.py

The reference (gym_futbol/envs/futbol_env.py::FutbolEnv) is imported from
/root/reference through oracle/ref_harness.py (gym/matplotlib stand-ins + injected
Philox draw stream, see that file); nothing of it is copied.  Every array is recorded
after each ``env.step`` and before the harness-level ``env.reset()`` that follows a
``done``.  Keys are ``<case>/<field>``; ``<case>/meta`` is a JSON string with the
constructor arguments, seed, env id and action source.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import philox  # noqa: E402
from oracle.ref_harness import rollout_v0  # noqa: E402

APPENDIX_A_ACTIONS = [0] * 9 + [5] + [0] * 6 + [10] + [0] * 3 + [15, 4, 1, 0, 0, 2, 8, 0, 0, 0]

FIELDS = ("action", "obs", "reward", "done", "owner", "last_owner", "ai_score", "opp_score", "draws")


def main():
    store = {}

    def add(case, seed, env_id, steps, random_opp, rng="philox", actions=None, **kw):
        out = rollout_v0(seed, env_id, steps, random_opp, actions=actions, rng=rng, **kw)
        for f in FIELDS:
            store["%s/%s" % (case, f)] = out[f]
        meta = dict(seed=seed, env_id=env_id, steps=steps, random_opp=random_opp, rng=rng,
                    actions="given" if actions is not None else "philox-stream-1", kwargs=kw)
        store["%s/meta" % case] = np.array(json.dumps(meta))
        print(case, "goals", int(out["ai_score"].max()), int(out["opp_score"].max()),
              "dones", int(out["done"].sum()), "draws", int(out["draws"][-1]))

    # A: BASELINE.json config #1 -- single env, 1000 steps (episodes of 401 + 401 + 198)
    for ro in (True, False):
        for seed, env_id in ((0, 0), (1, 7), (2, 4095)):
            add("trace_ro%d_s%d_e%d" % (ro, seed, env_id), seed, env_id, 1000, ro)
    # B: members of a batch (env ids 1000..1015 of seed 3) -- "env #k inside a batch"
    for ro in (True, False):
        for env_id in range(1000, 1016):
            add("batch_ro%d_s3_e%d" % (ro, env_id), 3, env_id, 100, ro)
    # C: constructor flags
    add("one_goal_end_ro0", 4, 5, 600, False, one_goal_end=True)
    add("one_goal_end_ro1", 4, 6, 600, True, one_goal_end=True)
    add("only_reward_goal_ro1", 5, 6, 600, True, only_reward_goal=True)
    add("game_time5_ro0", 6, 9, 200, False, game_time=5)
    # D: RNG-free known-answer traces + longer constant-RNG runs
    add("kat_appendix_a", 0, 0, 30, False, rng="const", actions=APPENDIX_A_ACTIONS)
    for ro in (True, False):
        acts = [philox.action_for(7, 0, t) for t in range(300)]
        add("const_rng_ro%d" % ro, 7, 0, 300, ro, rng="const", actions=acts)

    # libm fingerprint: numpy-scalar ``x**2`` is libm pow(x, 2.0), which is not correctly rounded, so
    # bit-exact float comparison against these vectors is only meaningful on a libm whose pow agrees
    # with the one that produced them.  Record inputs where pow(x,2.0) != x*x here.
    rng = np.random.RandomState(12345)
    xs = rng.uniform(-100, 100, 400000)
    ys = np.array([float(np.float64(x) ** 2) for x in xs])
    odd = np.flatnonzero(ys != xs * xs)[:128]
    store["libm_fingerprint/x"] = xs[odd]
    store["libm_fingerprint/y"] = ys[odd]
    store["libm_fingerprint/meta"] = np.array(json.dumps({"n_probed": 400000, "n_pow_ne_sq": int((ys != xs * xs).sum())}))
    print("libm fingerprint:", int((ys != xs * xs).sum()), "of 400000 pow(x,2) != x*x")

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "v0_golden.npz")
    np.savez_compressed(path, **store)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
