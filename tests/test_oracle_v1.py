"""The v1 oracle (CPU restatement of Futbol.step + the Chipmunk subset) against every external pin the
reference offers (the traces of the reference's own Python are in tests/test_oracle_v1_golden.py).  The physics at the pymunk
boundary is UNPINNED: pymunk cannot be run here and the reference has no test at this boundary; what CAN
be pinned from outside is checked: episode length 300
(gym_futbol/envs_v1/2v2/logs/evaluations.npz), observation/action shapes (saved-model JSON), kick-off
formations (team.py:52-112 re-derived independently below), single-body closed forms and restitution.
"""
import math

import numpy as np
import pytest

from oracle.v1 import OracleV1, team_actions

W, H = 105, 68


def formation_py(n, side):
    """team.py:52-112, transcribed as Python arithmetic (independent of the C restatement)."""
    if n <= 3:
        xs = [W * 0.25 if side == "left" else W * 0.75] * n
        ys = [H / (n + 1) * (i + 1) for i in range(n)]
    elif n <= 6:
        xs = ([W * 1 / 6] * 3 + [W * 2 / 6] * (n - 3)) if side == "left" else ([W * 5 / 6] * 3 + [W * 4 / 6] * (n - 3))
        ys = [H / 4 * (i + 1) for i in range(3)] + [H / (n - 3 + 1) * (i + 1) for i in range(n - 3)]
    else:
        xs = ([W * 1 / 8] * 4 + [W * 2 / 8] * 3 + [W * 3 / 8] * (n - 7)) if side == "left" else \
             ([W * 7 / 8] * 4 + [W * 6 / 8] * 3 + [W * 5 / 8] * (n - 7))
        ys = [H / 5 * (i + 1) for i in range(4)] + [H / 4 * (i + 1) for i in range(3)] + [H / (n - 7 + 1) * (i + 1) for i in range(n - 7)]
    return xs, ys


@pytest.mark.parametrize("n", range(1, 11))
def test_kickoff_formation_and_shapes(n):
    o = OracleV1(1, number_of_player=n)
    lx, ly = formation_py(n, "left")
    rx, ry = formation_py(n, "right")
    e = o.envs[0]
    assert np.array_equal(e["p"][:2 * n, 0], np.array(lx + rx)) and np.array_equal(e["p"][:2 * n, 1], np.array(ly + ry))
    assert tuple(e["p"][2 * n]) == (52.5, 34.0) and not e["v"].any()
    obs = o.obs()
    assert obs.shape == (4 + 8 * n,)                 # 2v2 -> 20, 5v5 -> 44, 10v10 -> 84 (saved-model JSON: Box(20,))
    assert np.all(np.abs(obs) <= 1.0) and not obs[:4].any()
    assert e["owner_side"] in (0, 1)


def test_episode_length_is_300_steps():
    """evaluations.npz: ep_lengths == 300 for all 1725 logged episodes (current_time > 30 after 300 x 0.1)."""
    o = OracleV1(3, seed=1, number_of_player=2)
    out = o.rollout(650, actions=None, autoreset=2)
    first = out["done"].argmax(0)
    assert (first == 299).all()
    assert (out["done"].sum(0) == 2).all() and out["done"][599].all()


def test_single_body_closed_forms():
    """noop + RIGHT gives delta v = 20 / 20 = 1; positions integrate the undamped velocity; v decays by 0.95^0.1."""
    o = OracleV1(1, number_of_player=1)
    e = o.envs[0]
    x0 = e["p"][0, 0]
    obs, r, d = o.step_one(0, [2, 0])
    assert e["p"][0, 0] == x0 + 1.0 * 0.1
    assert e["v"][0, 0] == 1.0 * 0.95 ** 0.1 and e["v"][0, 1] == 0.0
    # dash DOWN: delta v = 40 / 20 = 2
    y0, vx = e["p"][0, 1], e["v"][0, 0]
    o.step_one(0, [3, 1])
    assert e["p"][0, 1] == y0 + (-2.0) * 0.1 and e["v"][0, 1] == -2.0 * 0.95 ** 0.1
    assert e["v"][0, 0] == vx * 0.95 ** 0.1


def test_speed_clamps():
    o = OracleV1(1, number_of_player=1)
    e = o.envs[0]
    for _ in range(12):
        o.step_one(0, [2, 1])                        # dash right repeatedly
        assert math.hypot(*e["v"][0]) <= 10.0 + 1e-12
    assert abs(math.hypot(*e["v"][0]) - 10.0) < 1e-9 or e["p"][0, 0] > 100   # clamped at PLAYER_MAX_VELOCITY
    e["v"][2] = (300.0, 400.0)                        # ball far too fast: clamped to 25 in one velocity update
    e["p"][2] = (30.0, 10.0)
    o.step_one(0, [0, 0])
    assert abs(math.hypot(*e["v"][2]) - 25.0) < 1e-9


def test_two_body_restitution():
    """Head-on player/ball contact: relative normal velocity after = -e_a e_b (0.04) x before."""
    o = OracleV1(1, number_of_player=1)
    e = o.envs[0]
    e["p"][0] = (40.0, 34.0); e["v"][0] = (4.0, 0.0)
    e["p"][2] = (42.7, 34.0); e["v"][2] = (-3.0, 0.0)       # distance 2.7 -> 2.0 after integration: touching (< 2.5)
    e["p"][1] = (90.0, 60.0)
    before = e["v"][2, 0] - e["v"][0, 0]
    o.step_one(0, [0, 3])                                   # press with... arrow 0: player 0 would run to the ball; use key 3 arrow 1
    # redo with a neutral action (press + arrow pressed = nothing happens, futbol_env.py:390-391)
    o = OracleV1(1, number_of_player=1)
    e = o.envs[0]
    e["p"][0] = (40.0, 34.0); e["v"][0] = (4.0, 0.0)
    e["p"][2] = (42.7, 34.0); e["v"][2] = (-3.0, 0.0)
    e["p"][1] = (90.0, 60.0)
    o.lib.futbol_v1_oracle_step  # noqa: B018
    import ctypes as C
    act = np.array([1, 3], np.uint8)
    r = C.c_double()
    # right-team action is random; keep it away from the ball: it is at (90, 60)
    o.lib.futbol_v1_oracle_step(o.cfg.ctypes.data_as(C.c_void_p), C.c_void_p(o.envs[0:1].ctypes.data), act.ctypes.data_as(C.c_void_p), C.byref(r))
    after = e["v"][2, 0] - e["v"][0, 0]
    assert e["contacts"] == 1
    assert abs(after - (-(0.2 * 0.2) * before)) < 1e-12
    # momentum along x is conserved by the impulse pair (damping applied to both before the solve)
    d = 0.95 ** 0.1
    assert abs((20 * e["v"][0, 0] + 10 * e["v"][2, 0]) - (20 * 4.0 * d + 10 * -3.0 * d)) < 1e-9


def test_wall_contact_keeps_a_pushing_player_in():
    """A player dashing into the bottom wall: every step the position integrates the fresh 2.0 of velocity
    (0.2 inwards) before the solve, and the bias impulse removes bias_coef * (penetration - slop); the
    penetration therefore settles at slop + 0.2 / bias_coef, bias_coef = 1 - 0.9f**6 and slop = 0.1f (Chipmunk's
    defaults, written as C float literals in cpSpace.c)."""
    o = OracleV1(1, number_of_player=1)
    e = o.envs[0]
    e["p"][0] = (30.0, 6.0)
    for _ in range(60):
        o.step_one(0, [3, 1])
    f09, slop = float(np.float32(1.0) - np.float32(0.1)), float(np.float32(0.1))
    assert f09.hex() == "0x1.ccccccp-1".replace("p", "0000000p") and slop.hex() == "0x1.99999a0000000p-4"
    bias_coef = 1.0 - (f09 ** 60) ** 0.1
    assert abs(float(o.cfg["bias_coef"][0]) - bias_coef) < 1e-15 and float(o.cfg["slop"][0]) == slop
    assert abs((2.5 - e["p"][0, 1]) - (slop + 0.2 / bias_coef)) < 1e-6     # r_player + r_segment = 2.5 above y = 0
    assert abs(e["v"][0, 1]) < 1e-9                                       # the normal velocity is removed every step
    q_wall = 3 * 2 // 2 + 0 * 12 + 5
    assert e["age"][q_wall] == 0 and e["jn"][q_wall] > 0.0


def test_goal_rekickoff_and_out_of_bounds_happen_and_are_consistent():
    n_env, steps, N = 512, 600, 2
    o = OracleV1(n_env, seed=5, number_of_player=N)
    out = o.rollout(steps, actions=None, autoreset=2, n_threads=8)
    fl = out["flags"]
    goals, outs = (fl & 1) > 0, (fl & 2) > 0
    assert goals.sum() > 0 and outs.sum() > 0
    lx, ly = formation_py(N, "left")
    rx, ry = formation_py(N, "right")
    kick = np.concatenate([[0, 0, 0, 0]] + [[(x - 52.5) / 55.5, (y - 34.0) / 34.0, 0, 0] for x, y in zip(lx + rx, ly + ry)])
    t, i = np.argwhere(goals)[0]
    assert np.allclose(out["obs"][t, i], kick, atol=1e-3)      # the goal step returns the kick-off observation (+ v_bias * 1e-4)
    assert abs(abs(out["reward"][t, i]) - 1000) < 400           # +-1000 plus the shaped terms
    assert np.isfinite(out["obs"]).all() and np.abs(out["obs"]).max() < 1.6
    # the shaped reward is suppressed on out-of-bounds steps (futbol_env.py:463)
    assert (out["reward"][outs & ~goals] == 0).all()


def test_opponent_actions_are_uniform_over_25_pairs():
    a = np.stack([team_actions(0, i, 2, t, 5) for i in range(40) for t in range(50)])
    assert a.min() == 0 and a.max() == 4
    counts = np.bincount(a.reshape(-1), minlength=5) / a.size
    assert np.abs(counts - 0.2).max() < 0.02


def test_sharding_invariance_and_determinism():
    a = OracleV1(8, seed=3, env_id0=100, number_of_player=5).rollout(120)
    b = OracleV1(4, seed=3, env_id0=104, number_of_player=5).rollout(120, n_threads=2)
    assert np.array_equal(a["obs"][:, 4:], b["obs"]) and np.array_equal(a["reward"][:, 4:], b["reward"])
