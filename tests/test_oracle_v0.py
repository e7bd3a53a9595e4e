"""Pins the C oracle (oracle/futbol_v0_oracle.c) to the reference.

Golden vectors come from the unmodified reference (tests/golden/make_golden_v0.py); the
known-answer numbers of test_appendix_a_* are the ones printed in SURVEY.md Appendix A.
Bar: every integer output and the draw count exact; floats BIT-exact (the oracle computes
in the reference's operation order with the same libm).
"""
import numpy as np
import pytest

from oracle import philox
from oracle.v0 import OracleV0, lib

INT_FIELDS = ("done", "owner", "last_owner", "ai_score", "opp_score")


def _run_case(case, arith=1):
    m = case["meta"]
    kw = m["kwargs"]
    o = OracleV0(1, seed=m["seed"], env_id0=m["env_id"], random_opp=m["random_opp"], arith=arith,
                 rng_const=(m["rng"] == "const"), one_goal_end=kw.get("one_goal_end", False),
                 only_reward_goal=kw.get("only_reward_goal", False), game_time=kw.get("game_time", 40.0))
    return o.rollout(m["steps"], actions=case["action"], autoreset=1)


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert philox.philox4x32_10((0, 0, 0, 0), (0, 0)) == (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)
    assert philox.philox4x32_10((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2) == (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)
    assert philox.philox4x32_10((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0)) == \
        (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)
    ctr = np.array([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], np.uint32)
    key = np.array([0xA4093822, 0x299F31D0], np.uint32)
    out = np.zeros(4, np.uint32)
    lib().futbol_oracle_philox(ctr.ctypes.data, key.ctypes.data, out.ctypes.data)
    assert out.tolist() == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_action_stream_matches_python():
    from oracle.v0 import action_for
    tab = philox.actions_table(11, np.arange(5, 9), 3, 50)
    for j, e in enumerate(range(5, 9)):
        for t in range(3, 53):
            a = philox.action_for(11, e, t)
            assert a == action_for(11, e, t) == tab[t - 3, j]


def test_oracle_matches_reference_golden_bit_exact(golden_v0):
    """Floats are required BIT-exact when this box's libm pow() reproduces the fingerprint stored with
    the vectors (numpy-scalar x**2 is libm pow, which is not correctly rounded); otherwise 1e-12."""
    cases, same_libm = golden_v0["cases"], golden_v0["same_libm"]
    assert len(cases) >= 40
    for name, case in cases.items():
        out = _run_case(case)
        for f in INT_FIELDS:
            assert np.array_equal(out[f][:, 0], case[f]), (name, f)
        if case["meta"]["rng"] == "philox":
            assert np.array_equal(out["draws"][:, 0].astype(np.int64), case["draws"]), name
        steps = case["meta"]["steps"]
        ref = case["obs"].reshape(steps, 30)
        if same_libm:
            assert np.array_equal(out["obs"][:, 0], ref), name
        else:
            assert (np.abs(out["obs"][:, 0] - ref) <= 1e-12 * np.maximum(1.0, np.abs(ref))).all(), name
        assert np.array_equal(out["reward"][:, 0], case["reward"]), name


def test_oracle_kernel_arithmetic_mode_within_tolerance(golden_v0):
    """arith=0 (x*x, what the CUDA kernel computes) vs the reference: ints exact, floats 1e-9."""
    worst = 0.0
    for name, case in golden_v0["cases"].items():
        out = _run_case(case, arith=0)
        for f in INT_FIELDS:
            assert np.array_equal(out[f][:, 0], case[f]), (name, f)
        steps = case["meta"]["steps"]
        ref = case["obs"].reshape(steps, 30)
        err = np.abs(out["obs"][:, 0] - ref) / np.maximum(1.0, np.abs(ref))
        worst = max(worst, float(err.max()))
    assert worst <= 1e-9


def test_appendix_a_known_answers():
    """Numbers transcribed from SURVEY.md Appendix A (constant RNG, random_opp=False)."""
    acts = [0] * 9 + [5] + [0] * 6 + [10] + [0] * 3 + [15, 4, 1, 0, 0, 2, 8, 0, 0, 0]
    o = OracleV0(1, random_opp=False, rng_const=True)
    out = o.rollout(30, actions=np.array(acts, np.uint8), autoreset=1)
    rewards = [2, 2, 2, 2, 2, 2, 2, 2, 2, 18, 13, 13, 13, 13, 13, 13, -1, 2, 2, 2, -2, 20, 2, 13, 4, 1, 1, 2, 2, 2]
    N, O2, A1, A2 = 4, 3, 0, 1
    owners = [N] * 8 + [O2] + [A2] * 7 + [N] * 3 + [O2, O2, A1, A2, A2] + [O2] * 6
    assert out["reward"][:, 0].tolist() == [float(r) for r in rewards]
    assert out["owner"][:, 0].tolist() == owners
    assert not out["done"].any() and out["ai_score"].max() == 0 and out["opp_score"].max() == 0
    hexes = ["0x1.ca5cdc1ec1dfep+5", "0x1.154bc5377189ap+5", "-0x1.ad32ede641040p+0", "0x1.ba78baaf8fb80p+0", "0x1.8p+3",
             "0x1.c50f5a3f72476p+5", "0x1.1ad221e829e96p+5", "-0x1.07e9df0117d00p+0", "0x1.05518015d2600p+0", "0x1.8p+3"]
    assert out["obs"][29, 0, :10].tolist() == [float.fromhex(h) for h in hexes]
    # step 9: both opponents intercept in one step (Q11); last owner = OPP_1
    assert out["last_owner"][8, 0] == 2 and out["owner"][8, 0] == 3


def test_episode_length_and_reset():
    """First done=True on step 401 (SURVEY.md a13 / BASELINE.md section 2)."""
    o = OracleV0(3, seed=9, env_id0=100, random_opp=True)
    out = o.rollout(900, autoreset=1)
    for i in range(3):
        assert np.flatnonzero(out["done"][:, i]).tolist() == [400, 801]
    # the reset obs after construction: kickoff rows, owner row all zeros
    o2 = OracleV0(1)
    obs = o2.envs["obs"][0]
    assert obs[4].tolist() == [52.5, 34.0, 0, 0, 0] and obs[0].tolist() == [43.5, 39.0, 0, 0, 0]
    assert obs[3].tolist() == [61.5, 29.0, 0, 0, 0] and not obs[5].any()


def test_vec_semantics_and_threads_agree():
    a = OracleV0(64, seed=5, env_id0=10, random_opp=False, game_time=3.0)
    b = OracleV0(64, seed=5, env_id0=10, random_opp=False, game_time=3.0)
    oa = a.rollout(100, autoreset=2, n_threads=1)
    ob = b.rollout(100, autoreset=2, n_threads=4)
    for k in oa:
        assert np.array_equal(oa[k], ob[k]), k
    t, i = np.argwhere(oa["done"])[0]
    assert oa["obs"][t, i, 20:25].tolist() == [52.5, 34.0, 0, 0, 0]   # reset obs in the done slot
    assert not oa["obs"][t, i, 25:].any()


def test_trajectory_independent_of_batch_position():
    big = OracleV0(32, seed=3, env_id0=1000, random_opp=False).rollout(200)
    one = OracleV0(1, seed=3, env_id0=1017, random_opp=False).rollout(200)
    assert np.array_equal(big["obs"][:, 17], one["obs"][:, 0])
    assert np.array_equal(big["reward"][:, 17], one["reward"][:, 0])


def test_fm_functions_track_libm():
    """The specified log/sin/cos of the kernel arithmetic stay within ~1 ulp of libm on their domains."""
    import ctypes
    import math
    fm = lib().futbol_oracle_fm
    fm.argtypes = [ctypes.c_double, ctypes.c_void_p]
    out = np.zeros(3)
    rng = np.random.RandomState(7)
    xs = np.concatenate([rng.uniform(-6.3, 6.3, 20000), (rng.randint(1, 1 << 24, 20000)) / 16777216.0])
    for x in xs:
        fm(float(x), out.ctypes.data)
        assert abs(out[1] - math.sin(x)) <= 4.5e-16 and abs(out[2] - math.cos(x)) <= 4.5e-16
        if x > 0:
            assert abs(out[0] - math.log(x)) <= 2.3e-16 * max(abs(math.log(x)), 1e-300) or x == 1.0


def test_libm_mode_vs_kernel_mode_divergence_is_knife_edge_only():
    """arith=1 (== Python reference) vs arith=0 (== CUDA kernel) on 2048 envs x 1000 steps.

    Integers agree for nearly every env (measured on 8192 envs x 1000 steps: 0.5 % of the trajectories
    split with random opponents, 3.2 % with the hard-coded ones, i.e. 5e-6 / 3e-5 per env-step); where an
    env does split, the floats were still within 1e-9 (measured 4e-12) on the step before: the split is decided by a
    last-bit difference in a compare (the game has structural knife edges, e.g. an opponent chasing its
    own shot closes by exactly 0.1 per step and is tested with ``distance <= 1``), not by logic.
    """
    n, steps = 2048, 1000
    for random_opp in (True, False):
        a = OracleV0(n, seed=17, random_opp=random_opp, arith=1).rollout(steps, autoreset=2, n_threads=8)
        b = OracleV0(n, seed=17, random_opp=random_opp, arith=0).rollout(steps, autoreset=2, n_threads=8)
        mism = np.zeros((steps, n), bool)
        for f in INT_FIELDS:
            mism |= a[f] != b[f]
        err = np.abs(a["obs"] - b["obs"]).max(-1) / np.maximum(1.0, np.abs(a["obs"]).max(-1))
        split = mism | (err > 1e-6)           # first step at which the two trajectories visibly part
        diverged = split.any(0)
        first = np.where(diverged, split.argmax(0), steps)
        assert diverged.mean() <= 0.06, "more than 6 % of the 1000-step trajectories split"
        for i in np.flatnonzero(diverged):
            if first[i] > 0:
                assert err[first[i] - 1, i] <= 1e-9, (i, first[i])
        assert err[:, ~diverged].max() <= 1e-7   # (bimodal: measured <= 3e-9 for these, O(1) for the split ones)


def _wide():
    import json
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "v0_wide_golden.npz"))
    return {c: dict({k.split("/")[1]: z[k] for k in z.files if k.startswith(c + "/") and not k.endswith("meta")},
                    meta=json.loads(str(z[c + "/meta"]))) for c in ("ro0", "ro1")}


def wide_split_report(got, case):
    """Per (env, episode): does any integer output of the episode differ from the reference's?  -> (split episodes, episodes,
    mismatching env-steps)."""
    mism = np.zeros(case["done"].shape, bool)
    for f in INT_FIELDS:
        mism |= got[f].astype(np.int64) != case[f].astype(np.int64)
    ends = np.flatnonzero(case["done"][:, 0])            # the time limit: the same steps for every env
    bounds = [0] + [int(e) + 1 for e in ends] + ([mism.shape[0]] if (not len(ends) or ends[-1] + 1 < mism.shape[0]) else [])
    split = sum(int(mism[a:b].any(0).sum()) for a, b in zip(bounds[:-1], bounds[1:]))
    return split, (len(bounds) - 1) * mism.shape[1], int(mism.sum())


def test_wide_reference_set_libm_mode_is_exact(golden_v0):
    """128,000 steps of the unmodified reference (64 envs x 1000 steps x both opponent modes, tests/golden/
    make_golden_v0_wide.py): the oracle in libm mode reproduces every integer of every step, the draw counts, the rewards
    and (same libm) the last observation bit for bit."""
    for name, case in _wide().items():
        m = case["meta"]
        o = OracleV0(m["envs"], seed=m["seed"], env_id0=m["env_id0"], random_opp=m["random_opp"], arith=1)
        out = o.rollout(m["steps"], actions=case["action"], autoreset=1, n_threads=8)
        for f in INT_FIELDS:
            assert np.array_equal(out[f].astype(np.int64), case[f].astype(np.int64)), (name, f)
        draws = np.diff(np.concatenate([np.zeros((1, m["envs"]), np.uint64), out["draws"]]).astype(np.int64), axis=0)
        assert np.array_equal(draws, case["draws"].astype(np.int64)), name
        assert np.array_equal(out["reward"].astype(np.float32), case["reward"]), name
        if golden_v0["same_libm"]:
            assert np.array_equal(out["obs"][-1].reshape(-1, 6, 5), case["obs_last"]), name


def test_wide_reference_set_kernel_arithmetic_split_rate():
    """Kernel arithmetic (x*x for numpy-scalar x**2) against the same 128,000 reference steps: an episode may leave the
    reference's at a last-bit compare.  Measured here: 0 of 192 episodes with random opponents, 2 of 192 with the
    hard-coded ones; the bound is 6 % of the episodes, and every episode re-joins the reference at its reset."""
    for name, case in _wide().items():
        m = case["meta"]
        o = OracleV0(m["envs"], seed=m["seed"], env_id0=m["env_id0"], random_opp=m["random_opp"], arith=0)
        out = o.rollout(m["steps"], actions=case["action"], autoreset=1, n_threads=8)
        split, episodes, steps = wide_split_report(out, case)
        print("v0 wide %s: %d of %d episodes split (%d of %d env-steps differ)" % (name, split, episodes, steps, case["done"].size))
        assert split <= 0.06 * episodes, (name, split, episodes)
        first = np.flatnonzero(case["done"][:, 0])[0] + 1           # the first step of the second episode: all envs agree again
        for f in ("owner", "last_owner", "ai_score", "opp_score"):
            assert np.array_equal(out[f][first].astype(np.int64), case[f][first].astype(np.int64)), (name, f)
