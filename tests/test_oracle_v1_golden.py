"""Pins the v1 C oracle (oracle/futbol_v1_oracle.c) to the reference's own Python.

tests/golden/v1_golden.npz holds traces of the UNMODIFIED gym_futbol/envs_v1/{futbol_env,team,player,ball}.py run by
oracle/ref_harness_v1.py (tests/golden/make_golden_v1.py) over the pymunk stand-in (oracle/pymunk_standin.py:
pymunk / Chipmunk2D itself is absent, so the physics under the reference's game logic is that second restatement
of the DESIGN.md section 10 specification -- rows b3 / b6 of SURVEY.md section 8 stay "unpinned").  Bar: every
integer output, the draw count and the contact count exact; observation, reward and body state BIT-exact
(arith = 1: the oracle calls the same libm pow for the reference's Python-level ``x**2``); in kernel arithmetic
(arith = 0, what the CUDA kernels compute) integers exact and floats within 1e-9.
"""
import json
import math
import os

import numpy as np
import pytest

from oracle.v1 import OracleV1

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_REFERENCE = os.path.isfile("/root/reference/gym_futbol/envs_v1/futbol_env.py")


@pytest.fixture(scope="module")
def golden_v1():
    z = np.load(os.path.join(ROOT, "tests", "golden", "v1_golden.npz"))
    cases = {}
    for key in z.files:
        case, field = key.split("/")
        cases.setdefault(case, {})[field] = z[key]
    for case in cases.values():
        case["meta"] = json.loads(str(case["meta"]))
    fp = cases.pop("libm_fingerprint")
    coverage = cases.pop("coverage")["meta"]
    same_libm = all(float(x) ** 2 == float(y) for x, y in zip(fp["x"], fp["y"]))
    return {"cases": cases, "same_libm": same_libm, "coverage": coverage}


def run_oracle_case(case, arith):
    """Replays a golden case through the oracle; returns per-step arrays in the fixture's layout."""
    m = case["meta"]
    N, T = m["number_of_player"], m["steps"]
    o = OracleV1(1, seed=m["seed"], env_id0=m["env_id"], number_of_player=N, total_time=m["kwargs"].get("total_time", 30), arith=arith)
    B = 2 * N + 1
    out = {"obs0": o.obs(0), "obs": np.zeros((T, 4 + 8 * N)), "reward": np.zeros(T), "done": np.zeros(T, np.uint8),
           "flags": np.zeros(T, np.uint8), "owner_side": np.zeros(T, np.uint8), "draws": np.zeros(T, np.int32),
           "contacts": np.zeros(T, np.int32), "bodies": np.zeros((T, B, 6)), "overflow": 0}
    for t in range(T):
        obs, r, d = o.step_one(0, case["action"][t])
        e = o.envs[0]
        out["obs"][t], out["reward"][t], out["done"][t] = obs, r, d
        out["flags"][t], out["owner_side"][t], out["draws"][t], out["contacts"][t] = e["flags"], e["owner_side"], e["step_draws"], e["contacts"]
        out["bodies"][t] = np.concatenate([e["p"][:B], e["v"][:B], e["vb"][:B]], axis=1)
        if d:
            o.reset()
    out["overflow"] = int(o.envs[0]["overflow"])
    return out


INT_FIELDS = ("done", "flags", "owner_side", "draws", "contacts")


def test_golden_set_covers_the_game(golden_v1):
    cov = golden_v1["coverage"]
    assert len(golden_v1["cases"]) >= 40 and cov["steps"] >= 15000
    assert min(cov["pass_arrows"]) >= 100          # passes with every arrow key (team.py:148-178)
    assert min(cov["out_walls"]) >= 3              # every boundary segment of check_and_fix_out_bounds (:247-287)
    assert cov["goals_left"] >= 10 and cov["goals_right"] >= 10
    assert cov["two_draw_steps"] >= 100            # steps with two or more sequential draws
    assert cov["arbiters"]["inherited_not_warm"] >= 100 and cov["arbiters"]["warm_started"] >= 1000
    assert {c["meta"]["number_of_player"] for c in golden_v1["cases"].values()} >= {1, 2, 3, 4, 5, 7, 10}


def test_oracle_matches_reference_golden_bit_exact(golden_v1):
    cases, same_libm = golden_v1["cases"], golden_v1["same_libm"]
    for name, case in cases.items():
        out = run_oracle_case(case, arith=1)
        assert out["overflow"] == 0, name
        for f in INT_FIELDS:
            assert np.array_equal(out[f], case[f]), (name, f)
        bodies = out["bodies"] if case["meta"]["full_bodies"] else out["bodies"][-1:]
        if same_libm:
            assert np.array_equal(out["obs0"], case["obs0"]) and np.array_equal(out["obs"], case["obs"]), name
            assert np.array_equal(out["reward"], case["reward"]) and np.array_equal(bodies, case["bodies"]), name
        else:
            for a, b in ((out["obs"], case["obs"]), (out["reward"], case["reward"]), (bodies, case["bodies"])):
                assert (np.abs(a - b) <= 1e-12 * np.maximum(1.0, np.abs(b))).all(), name


def test_oracle_kernel_arithmetic_mode_within_tolerance(golden_v1):
    """arith = 0 (x*x: what the CUDA kernels compute) against the reference traces: integers exact, floats 1e-9."""
    worst = 0.0
    for name, case in golden_v1["cases"].items():
        out = run_oracle_case(case, arith=0)
        for f in INT_FIELDS:
            assert np.array_equal(out[f], case[f]), (name, f)
        for a, b in ((out["obs"], case["obs"]), (out["reward"], case["reward"])):
            worst = max(worst, float((np.abs(a - b) / np.maximum(1.0, np.abs(b))).max()))
    assert worst <= 1e-9


def test_chipmunk_default_constants():
    """cpSpace.c writes its defaults as C float literals; the oracle, the stand-in and the CUDA library's literals agree."""
    from oracle import pymunk_standin
    o = OracleV1(1, number_of_player=2)
    sp = pymunk_standin.Space()
    slop, f09 = float(np.float32(0.1)), float(np.float32(1.0) - np.float32(0.1))
    assert sp.collision_slop == slop == float(o.cfg["slop"][0]) and slop.hex() == "0x1.99999a0000000p-4"
    assert sp.collision_bias == math.pow(f09, 60.0) and f09.hex() == "0x1.cccccc0000000p-1"
    bias_coef, damping = 1.0 - math.pow(sp.collision_bias, 0.1), math.pow(0.95, 0.1)
    assert float(o.cfg["bias_coef"][0]) == bias_coef and float(o.cfg["damping_dt"][0]) == damping
    src = open(os.path.join(ROOT, "gym_futbol_b200", "csrc", "capi.cu")).read()
    assert "Q.bias_coef = %s;" % bias_coef.hex() in src and "Q.damping_dt = %s;" % damping.hex() in src
    assert "kSlop = (double)0.1f" in open(os.path.join(ROOT, "gym_futbol_b200", "csrc", "v1_step.cuh")).read()


def test_standin_warm_start_rule():
    """A pair that re-touches after one step apart inherits its accumulated impulse but is not warm-started
    (cpArbiterUpdate marks a cached arbiter FIRST_COLLISION; cpArbiterApplyCachedImpulse returns early for it)."""
    from oracle import pymunk_standin as pm
    sp = pm.Space()
    a, b = pm.Body(20, 22.5), pm.Body(10, 5)
    sa, sb = pm.Circle(a, 1.5), pm.Circle(b, 1.0)
    sp.add(a, sa, b, sb)
    a.position, b.position = (0, 0), (2.4, 0)
    a.velocity = (1, 0)
    sp.step(0.1)                                     # touching: new arbiter
    assert sp.counters == {"warm_started": 0, "inherited_not_warm": 0, "new": 1}
    jn = sp.cached_arbiters[(id(sa), id(sb))].contacts[0].jn_acc
    assert jn > 0
    b.position = (10, 0)
    sp.step(0.1)                                     # apart for one step: the arbiter stays cached
    b.position, b.velocity, a.velocity = (a.position.x + 2.4, 0), (0, 0), (0, 0)
    va = a.velocity.x
    sp.step(0.1)
    assert sp.counters["inherited_not_warm"] == 1
    # at rest and overlapping by exactly the slop: no warm-start impulse was applied, and the inherited impulse is
    # taken back by the clamp (jnAcc = max(jnOld + jn, 0) with jn = 0 keeps it; nothing pushed the bodies apart)
    assert a.velocity.x == va == 0.0 and b.velocity.x == 0.0
    sp.step(0.1)
    assert sp.counters["warm_started"] == 1          # touching in consecutive steps: now it is warm-started


@pytest.mark.skipif(not HAVE_REFERENCE, reason="needs /root/reference (build container only)")
@pytest.mark.parametrize("N", list(range(1, 11)))
def test_oracle_matches_live_reference(N):
    """Fresh seeds, every team size: the unmodified reference executed now against the oracle (arith = 1), bit for bit."""
    from oracle.ref_harness_v1 import chase_and_kick_policy, rollout_v1
    for seed, policy, steps in ((100 + N, None, 330), (200 + N, chase_and_kick_policy, 200)):
        ref = rollout_v1(seed, 9000 + N, steps, N, policy=policy)
        ref.pop("coverage")
        case = dict(ref, meta={"seed": seed, "env_id": 9000 + N, "steps": steps, "number_of_player": N, "kwargs": {}})
        out = run_oracle_case(case, arith=1)
        for f in INT_FIELDS + ("obs", "reward", "bodies", "obs0"):
            assert np.array_equal(out[f], ref[f]), (N, seed, f)


@pytest.mark.skipif(not HAVE_REFERENCE, reason="needs /root/reference (build container only)")
def test_golden_fixture_regenerates(golden_v1):
    """The committed fixture is what the committed generator produces from the reference today."""
    from oracle.ref_harness_v1 import chase_and_kick_policy, rollout_v1
    for name in ("trace_n2_s0_e0", "directed_n3_s13", "batch_n5_s3_e2003"):
        case = golden_v1["cases"][name]
        m = case["meta"]
        ref = rollout_v1(m["seed"], m["env_id"], m["steps"], m["number_of_player"],
                         policy=chase_and_kick_policy if m["actions"] == "chase_and_kick_policy" else None, **m["kwargs"])
        for f in ("action", "obs", "reward", "done", "flags", "owner_side", "draws", "contacts"):
            assert np.array_equal(ref[f], case[f]), (name, f)
