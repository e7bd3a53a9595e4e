"""Runs the UNMODIFIED reference v1 ``Futbol`` (gym_futbol/envs_v1/*.py) over stand-ins for its absent imports.

TEST INFRASTRUCTURE ONLY.  Used in the build container to generate the golden fixtures
tests/golden/v1_golden.npz (tests/golden/make_golden_v1.py); the product path never imports this file.

What is injected (no reference source is modified or copied):
  * ``gym`` / ``matplotlib.pyplot``: the stand-ins of oracle/ref_harness.py.
  * ``pymunk``: oracle/pymunk_standin.py -- a pure-Python restatement of the Chipmunk2D subset the reference
    drives, per the specification of DESIGN.md section 10.  The game logic that runs on top of it is the
    reference's own code; the physics under it is ours (see that file's header for what this does and does
    not pin).
  * randomness (the reference is unseeded): the module-level name ``random`` of envs_v1/futbol_env.py (:11) and
    envs_v1/team.py (:5) is rebound to a shim backed by the Philox streams of oracle/philox.py, and the
    ``action_space`` object's ``sample`` (:307) draws the right team's actions from stream 2:
      - ``random.choice(["left", "right"])`` inside ``reset()`` (:147): stream 3, block 0x4000, word 0 at the
        env's current total step t: index = w * 2 >> 32;
      - inside ``step`` the calls are sequential draws j = 0, 1, ... of stream 3 at step t:
        ``random.choice(seq)`` (:275, :278, :475): index = w * len(seq) >> 32;
        ``random.choices(pop, weights)`` (team.py:141-178): the ((w >> 8) * c >> 24)-th of the c entries with a
        non-zero weight, in order;
      - ``action_space.sample()`` (:429): entry j = word (j & 3) of block (j >> 2) of stream 2 at step t, * 5 >> 32.
    Synthetic left-team actions come from stream 1 with the same layout.
Protocol: construct (``Futbol.__init__`` calls ``reset()`` itself, :127), then ``step``; after a ``done`` the
harness calls ``reset()``.
"""
from __future__ import annotations

import importlib
import sys

import numpy as _np

from . import philox, pymunk_standin, ref_harness

_loaded = None


class _RandomShimV1:
    """Stands in for the stdlib ``random`` module inside envs_v1/futbol_env.py and envs_v1/team.py."""

    def __init__(self):
        self.env = None      # the RefEnvV1 being driven

    def choice(self, seq):
        e = self.env
        if not e.in_step:
            if list(seq) != ["left", "right"]:
                raise RuntimeError("unexpected random.choice outside step()")
            w = philox.philox4x32_10(philox.step_counter(e.t_total, 0x4000, e.env_id, philox.STREAM_V1_DYNAMICS), e.stream.key)[0]
            return seq[(w * 2) >> 32]
        w = e.stream.next_u32()
        return seq[(w * len(seq)) >> 32]

    def choices(self, population, weights=None, k=1):
        if k != 1 or weights is None:
            raise RuntimeError("unexpected random.choices signature")
        elig = [i for i, wt in enumerate(weights) if wt]
        w = self.env.stream.next_u32()
        return [population[elig[((w >> 8) * len(elig)) >> 24]]]


def team_actions(seed, env_id, stream, t, n_players):
    """MultiDiscrete([5, 5] * N).sample() for one team at total step t (numpy int array of 2N)."""
    key = philox.seed_key(seed)
    out = _np.zeros(2 * n_players, dtype=_np.int64)
    blk = None
    for j in range(2 * n_players):
        if j & 3 == 0:
            blk = philox.philox4x32_10(philox.step_counter(t, j >> 2, env_id, stream), key)
        out[j] = (blk[j & 3] * 5) >> 32
    return out


def load_reference_v1():
    """Import the reference v1 modules (once) over the stand-ins and inject the RNG shim."""
    global _loaded
    if _loaded is not None:
        return _loaded
    root = ref_harness.find_reference_root()
    if root is None:
        raise RuntimeError("reference package not found")
    ref_harness._install_stubs()
    pymunk_standin.install()
    if root not in sys.path:
        sys.path.insert(0, root)
    fe = importlib.import_module("gym_futbol.envs_v1.futbol_env")
    team = importlib.import_module("gym_futbol.envs_v1.team")
    shim = _RandomShimV1()
    fe.random = shim
    team.random = shim
    _loaded = (fe, team, shim)
    return _loaded


class RefEnvV1:
    """One reference ``Futbol`` bound to its own Philox streams."""

    def __init__(self, seed=0, env_id=0, number_of_player=2, total_time=30):
        self.fe, _, self.shim = load_reference_v1()
        self.seed, self.env_id, self.N = int(seed), int(env_id), int(number_of_player)
        self.stream = philox.DrawStream(seed, env_id, philox.STREAM_V1_DYNAMICS)
        self.t_total = 0
        self.in_step = False
        self.shim.env = self
        self.env = self.fe.Futbol(number_of_player=number_of_player, total_time=total_time)
        self.env.action_space.sampler = lambda: team_actions(self.seed, self.env_id, philox.STREAM_V1_OPP, self.t_total, self.N)
        # instrumentation (instance attributes wrapping bound methods; the class is untouched)
        self.last_out = self.last_goal = False
        self.goal_ball_x = None
        self._orig_out, self._orig_goal = self.env.check_and_fix_out_bounds, self.env.ball_contact_goal

        def out_wrapper():
            self.last_out = bool(self._orig_out())
            return self.last_out

        def goal_wrapper():
            self.last_goal = bool(self._orig_goal())
            self.goal_ball_x = self.env.ball.get_position()[0]
            return self.last_goal

        self.env.check_and_fix_out_bounds = out_wrapper
        self.env.ball_contact_goal = goal_wrapper
        self.pass_arrows = []        # arrow key of every pass that happened (team.py:136), for coverage reports
        self.out_walls = []          # wall index of every out-of-bounds fix (:247-254)
        for team in (self.env.team_A, self.env.team_B):
            def pass_wrapper(player, arrow_keys, _orig=team.get_pass_target_teammate):
                self.pass_arrows.append(int(arrow_keys))
                return _orig(player, arrow_keys=arrow_keys)
            team.get_pass_target_teammate = pass_wrapper
        _orig_wall = self.env.ball_contact_wall

        def wall_wrapper():
            hit, idx = _orig_wall()
            if hit:
                self.out_walls.append(int(idx))
            return hit, idx

        self.env.ball_contact_wall = wall_wrapper

    def reset(self):
        self.shim.env = self
        self.in_step = False
        return self.env.reset()

    def step(self, action=None):
        """action: 2N ints (arrow, key per left player) or None = the synthetic stream-1 actions."""
        self.shim.env = self
        if action is None:
            action = team_actions(self.seed, self.env_id, philox.STREAM_ACTIONS, self.t_total, self.N)
        self.stream.begin_step(self.t_total)
        self.in_step = True
        n_log = len(self.env.space.step_log)
        try:
            obs, reward, done, info = self.env.step(_np.asarray(action))
        finally:
            self.in_step = False
        self.t_total += 1
        self.contacts = self.env.space.step_log[n_log][1]      # of the 0.1 s space step (a goal adds a 1e-4 step)
        return obs, reward, done, info

    def bodies(self):
        """[2N+1, 6] float64: x, y, vx, vy, v_bias_x, v_bias_y in the oracle's body order (A, B, ball)."""
        bl = [p.body for p in self.env.player_arr] + [self.env.ball.body]
        return _np.array([[b._p[0], b._p[1], b._v[0], b._v[1], b._v_bias[0], b._v_bias[1]] for b in bl], dtype=_np.float64)


def chase_and_kick_policy(env, t):
    """Scripted left-team actions for the directed golden cases: a player touching the ball passes (cycling through the
    five arrow keys) or shoots, a player away from it presses towards it (:371-391) or dashes along an arrow."""
    N = env.N
    ball = env.env.ball
    a = _np.zeros(2 * N, dtype=_np.int64)
    for i, pl in enumerate(env.env.team_A.player_array):
        if ball.has_contact_with(pl):
            a[2 * i], a[2 * i + 1] = ((t + i) % 5, 4) if (t // 7 + i) % 3 else (0, 2)
        elif (t + 3 * i) % 11 == 0:
            a[2 * i], a[2 * i + 1] = 1 + (t + i) % 4, 1
        else:
            a[2 * i], a[2 * i + 1] = 0, 3
    return a


def dribble_policy(target):
    """Scripted left-team actions: press to the ball, then dribble it (dash while touching, :338-341 + :300-304)
    towards ``target`` = (x, y); used to reach every boundary segment of check_and_fix_out_bounds (:247-287)."""
    def policy(env, t):
        N = env.N
        ball = env.env.ball
        a = _np.zeros(2 * N, dtype=_np.int64)
        for i, pl in enumerate(env.env.team_A.player_array):
            if ball.has_contact_with(pl):
                px, py = pl.get_position()
                dx, dy = target[0] - px, target[1] - py
                arrow = (2 if dx > 0 else 4) if abs(dx) > abs(dy) else (1 if dy > 0 else 3)
                a[2 * i], a[2 * i + 1] = arrow, 1
            else:
                a[2 * i], a[2 * i + 1] = 0, 3
        return a
    return policy


def rollout_v1(seed, env_id, steps, number_of_player, actions=None, total_time=30, reset_on_done=True, policy=None):
    """Step one reference env ``steps`` times; per-step arrays (recorded AFTER the step, BEFORE the harness reset).
    actions: [steps, 2N] or None; policy: callable(env, t) -> 2N actions (wins over ``actions``); neither = stream 1."""
    env = RefEnvV1(seed=seed, env_id=env_id, number_of_player=number_of_player, total_time=total_time)
    N, D = env.N, 4 + 8 * env.N
    out = {"action": _np.zeros((steps, 2 * N), _np.uint8), "obs": _np.zeros((steps, D)), "reward": _np.zeros(steps),
           "done": _np.zeros(steps, _np.uint8), "flags": _np.zeros(steps, _np.uint8), "owner_side": _np.zeros(steps, _np.uint8),
           "bodies": _np.zeros((steps, 2 * N + 1, 6)), "draws": _np.zeros(steps, _np.int32), "contacts": _np.zeros(steps, _np.int32),
           "obs0": _np.asarray(env.env.observation, dtype=_np.float64).copy()}
    for t in range(steps):
        if policy is not None:
            a = _np.asarray(policy(env, t))
        elif actions is not None:
            a = _np.asarray(actions[t])
        else:
            a = team_actions(seed, env_id, philox.STREAM_ACTIONS, env.t_total, N)
        obs, r, d, _ = env.step(a)
        fl = (1 if env.last_goal else 0) | (2 if env.last_out else 0) | (4 if d else 0)
        if env.last_goal and env.goal_ball_x > env.env.width - 2:
            fl |= 8
        out["action"][t], out["obs"][t], out["reward"][t], out["done"][t], out["flags"][t] = a, obs, r, d, fl
        out["owner_side"][t] = 0 if env.env.ball_owner_side == "left" else 1
        out["bodies"][t], out["draws"][t], out["contacts"][t] = env.bodies(), env.stream.ctr, env.contacts
        if d and reset_on_done:
            env.reset()
    out["coverage"] = {"pass_arrows": [env.pass_arrows.count(k) for k in range(5)], "out_walls": [env.out_walls.count(k) for k in range(6)],
                       "goals_left": int(((out["flags"] & 9) == 9).sum()), "goals_right": int(((out["flags"] & 9) == 1).sum()),
                       "arbiters": dict(env.env.space.counters), "draws": int(out["draws"].sum()), "contacts": int(out["contacts"].sum()),
                       "max_contacts": int(out["contacts"].max())}
    return out
