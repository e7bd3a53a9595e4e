#!/usr/bin/env python
"""Generates tests/golden/v1_golden.npz by running the UNMODIFIED reference v1 ``Futbol``.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_v1.py

The reference (gym_futbol/envs_v1/{futbol_env,team,player,ball}.py) is imported from /root/reference through
oracle/ref_harness_v1.py: gym / matplotlib stand-ins, the pure-Python pymunk stand-in oracle/pymunk_standin.py
(pymunk itself is absent: the physics under the reference's game logic is that restatement, see its header) and
the injected Philox streams.  Nothing of the reference is copied.  Every array is recorded after each
``env.step`` and before the harness-level ``env.reset()`` that follows a ``done``.  Keys are ``<case>/<field>``;
``<case>/meta`` is a JSON string with the constructor arguments, seed, env id, action source and coverage counts.

Fields: action u8 [T, 2N]; obs f64 [T, 4+8N] (the reference's return value); reward f64 [T]; done u8 [T];
flags u8 [T] (1 goal, 2 out of bounds, 4 done, 8 the goal was the left team's); owner_side u8 [T];
draws i32 [T] (sequential stream-3 draws of the step); contacts i32 [T] (contacts solved in the 0.1 s step);
obs0 f64 [4+8N] (observation after construction); bodies f64 [T, 2N+1, 6] (x, y, vx, vy, v_bias_x, v_bias_y of
team A, team B, ball -- kept for the short cases, last step only for the 1000-step traces).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ref_harness_v1 import chase_and_kick_policy, dribble_policy, rollout_v1  # noqa: E402

FIELDS = ("action", "obs", "reward", "done", "flags", "owner_side", "draws", "contacts", "obs0", "bodies")


def main():
    store = {}
    total = {"pass_arrows": [0] * 5, "out_walls": [0] * 6, "goals_left": 0, "goals_right": 0, "steps": 0,
             "arbiters": {"warm_started": 0, "inherited_not_warm": 0, "new": 0}, "two_draw_steps": 0}

    def add(case, seed, env_id, steps, N, full_bodies, policy=None, **kw):
        out = rollout_v1(seed, env_id, steps, N, policy=policy, **kw)
        cov = out.pop("coverage")
        if not full_bodies:
            out["bodies"] = out["bodies"][-1:]
        for f in FIELDS:
            store["%s/%s" % (case, f)] = out[f]
        meta = dict(seed=seed, env_id=env_id, steps=steps, number_of_player=N, kwargs=kw, full_bodies=full_bodies,
                    actions=getattr(policy, "__name__", "dribble_policy") if policy is not None else "philox-stream-1", coverage=cov)
        store["%s/meta" % case] = np.array(json.dumps(meta))
        for k in range(5):
            total["pass_arrows"][k] += cov["pass_arrows"][k]
        for k in range(6):
            total["out_walls"][k] += cov["out_walls"][k]
        for k in total["arbiters"]:
            total["arbiters"][k] += cov["arbiters"][k]
        total["goals_left"] += cov["goals_left"]
        total["goals_right"] += cov["goals_right"]
        total["steps"] += steps
        total["two_draw_steps"] += int((out["draws"] >= 2).sum())
        print("%-28s goals %d/%d out %s passes %s contacts %d (max %d) arbiters %s" % (
            case, cov["goals_left"], cov["goals_right"], cov["out_walls"], cov["pass_arrows"], cov["contacts"], cov["max_contacts"], cov["arbiters"]))

    # A: 1000-step traces (episodes of 300 + 300 + 300 + 100) under uniform random actions, both teams
    for N, runs in ((1, ((0, 0), (1, 7))), (2, ((0, 0), (1, 7))), (5, ((0, 0), (1, 7))), (10, ((1, 7),))):
        for seed, env_id in runs:
            add("trace_n%d_s%d_e%d" % (N, seed, env_id), seed, env_id, 1000, N, False)
    for N, steps in ((3, 500), (7, 350)):
        add("trace_n%d_s2_e4095" % N, 2, 4095, steps, N, False)
    # B: members of a batch (env ids 2000..2007 of seed 3): "env #k inside a batch"
    for N in (2, 5):
        for env_id in range(2000, 2008):
            add("batch_n%d_s3_e%d" % (N, env_id), 3, env_id, 100, N, True)
    # C: directed play (scripted left team: press, pass with every arrow, shoot): goals, out-of-bounds fixes at
    #    every wall, two-draw passes, pairs that re-touch after one or two steps apart
    for N, seed in ((1, 11), (2, 12), (3, 13), (5, 14), (10, 15)):
        add("directed_n%d_s%d" % (N, seed), seed, 31 + N, 1000 if N < 10 else 500, N, N <= 2, policy=chase_and_kick_policy)
    for seed in (21, 22, 23):
        add("directed_n2_s%d" % seed, seed, 77, 600, 2, False, policy=chase_and_kick_policy)
    for k, target in enumerate(((-5, 12), (-5, 56), (50, 75), (110, 12), (110, 56), (50, -5))):   # one per boundary segment
        add("dribble_wall%d_n2" % k, 30 + k, 500 + k, 300, 2, False, policy=dribble_policy(target))
    # D: constructor arguments
    add("total_time5_n2", 6, 9, 200, 2, True, total_time=5)
    add("total_time0p35_n4", 7, 3, 40, 4, True, total_time=0.35)

    # libm fingerprint (as v0): Python float ``x**2`` is libm pow(x, 2.0), not always x*x
    rng = np.random.RandomState(54321)
    xs = rng.uniform(-100, 100, 400000)
    ys = np.array([float(x) ** 2 for x in xs])
    odd = np.flatnonzero(ys != xs * xs)[:128]
    store["libm_fingerprint/x"], store["libm_fingerprint/y"] = xs[odd], ys[odd]
    store["libm_fingerprint/meta"] = np.array(json.dumps({"n_probed": 400000, "n_pow_ne_sq": int((ys != xs * xs).sum())}))
    store["coverage/meta"] = np.array(json.dumps(total))
    print("coverage over all cases:", json.dumps(total))
    assert min(total["pass_arrows"]) > 0 and min(total["out_walls"]) > 0 and total["goals_left"] > 0 and total["goals_right"] > 0
    assert total["arbiters"]["inherited_not_warm"] > 0 and total["two_draw_steps"] > 0

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "v1_golden.npz")
    np.savez_compressed(path, **store)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
