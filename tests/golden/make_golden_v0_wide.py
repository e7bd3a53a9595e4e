#!/usr/bin/env python
"""Generates tests/golden/v0_wide_golden.npz: 64 envs x 1000 steps per opponent mode from the UNMODIFIED reference.

    python tests/golden/make_golden_v0_wide.py          (build container only: needs /root/reference)

A wide, integers-only companion of v0_golden.npz (same harness: oracle/ref_harness.py).  Its purpose is statistical: the
kernel squares with x*x where numpy-scalar x**2 is libm pow(x, 2.0), so a trajectory can leave the reference's at a
last-bit compare; this set measures how often over 128,000 reference steps (tests assert the rate).  Per case
(``ro0`` / ``ro1``): action u8 [T, E]; done, owner, last_owner u8 [T, E]; ai_score, opp_score i16 [T, E];
reward f32 [T, E]; draws u8 [T, E] (draws of the step); obs_last f64 [E, 6, 5].  Env ids 5000..5063, seed 17.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ref_harness import rollout_v0  # noqa: E402

SEED, ENV0, ENVS, STEPS = 17, 5000, 64, 1000


def main():
    store = {}
    for ro in (0, 1):
        cols = {k: [] for k in ("action", "done", "owner", "last_owner", "ai_score", "opp_score", "reward", "draws", "obs_last")}
        for e in range(ENVS):
            out = rollout_v0(SEED, ENV0 + e, STEPS, bool(ro))
            for k in ("action", "done", "owner", "last_owner", "ai_score", "opp_score", "reward"):
                cols[k].append(out[k])
            cols["draws"].append(np.diff(np.concatenate([[0], out["draws"]])))
            cols["obs_last"].append(out["obs"][-1])
        case = "ro%d" % ro
        for k, dt in (("action", np.uint8), ("done", np.uint8), ("owner", np.uint8), ("last_owner", np.uint8), ("ai_score", np.int16),
                      ("opp_score", np.int16), ("reward", np.float32), ("draws", np.uint8)):
            store["%s/%s" % (case, k)] = np.stack(cols[k], axis=1).astype(dt)
        store["%s/obs_last" % case] = np.stack(cols["obs_last"])
        store["%s/meta" % case] = np.array(json.dumps(dict(seed=SEED, env_id0=ENV0, envs=ENVS, steps=STEPS, random_opp=bool(ro))))
        print(case, "goals", int(store[case + "/ai_score"].max()), int(store[case + "/opp_score"].max()), "dones", int(store[case + "/done"].sum()),
              "mean draws", float(store[case + "/draws"].mean()), "max draws", int(store[case + "/draws"].max()))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "v0_wide_golden.npz")
    np.savez_compressed(path, **store)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
