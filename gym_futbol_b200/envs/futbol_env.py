"""Drop-in for the reference's single-environment class ``gym_futbol.envs.FutbolEnv``.

Same constructor keywords, spaces, ``reset() -> obs`` and ``step(a) -> (obs, reward, done, {})`` as
gym_futbol/envs/futbol_env.py:132-717; the work is done by the CUDA step kernel on a 1-env batch
(float64 outputs, no auto-reset -- the reference does not auto-reset either).  For throughput use
``gym_futbol_b200.FutbolVecEnv``; this class exists so that code written against the reference
runs unchanged and so that parity tests read like the reference's own usage.

Differences, all deliberate:
  * randomness is the seeded counter-based Philox stream (``seed=``, ``env_id=``); the reference is
    unseeded.
  * ``length`` / ``width`` / ``goal_size`` must keep their defaults: the reference honours them only
    partially (reset and score() use module constants, futbol_env.py:211-219,581-582).
  * out-of-range actions raise ``ValueError`` (the reference prints and carries on).
  * the returned observation is a fresh numpy array, not a live view of internal state.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import spaces
from ..vec_env import FutbolVecEnv
from .action import Action  # noqa: F401  (re-exported like the reference module does)
from .ballowner import BallOwner

try:  # gym is optional
    import gym as _gym
    _Base = _gym.Env
except Exception:  # noqa: BLE001
    _Base = object

FIELD_LEN, FIELD_WID, GOAL_SIZE = 105, 68, 10
SHOOT_SPEED, PLARYER_SPEED_W_BALL, GAME_TIME, PRESSURE_RANGE = 20, 12, 40, 2


class FutbolEnv(_Base):
    def __init__(self, length=FIELD_LEN, width=FIELD_WID, goal_size=GOAL_SIZE, game_time=GAME_TIME,
                 player_speed=PLARYER_SPEED_W_BALL, shoot_speed=SHOOT_SPEED, Debug=False,
                 pressure_range=PRESSURE_RANGE, one_goal_end=False, action_as_int=True, only_reward_goal=False,
                 random_opp=True, seed=0, env_id=0, device="cuda:0"):
        if (length, width, goal_size) != (FIELD_LEN, FIELD_WID, GOAL_SIZE):
            raise NotImplementedError("length/width/goal_size are only supported at the reference defaults "
                                      "(105, 68, 10): the reference itself honours them only partially")
        self.length, self.width, self.goal_size = length, width, goal_size
        self.goal_up, self.goal_down = width / 2 + goal_size / 2, width / 2 - goal_size / 2
        self.game_time, self.player_speed, self.shoot_speed = game_time, player_speed, shoot_speed
        self.one_goal_end, self.Debug = one_goal_end, Debug
        self.action_as_int, self.only_reward_goal, self.random_opp = action_as_int, only_reward_goal, random_opp
        if action_as_int:
            self.action_space = spaces.Discrete(16)
        else:
            self.action_space = spaces.Tuple((spaces.Discrete(4), spaces.Discrete(4)))
        self.observation_space = spaces.Box(
            low=np.array([[0, 0, -length, -width, 0]] * 6, dtype=np.float64),
            high=np.array([[length, width, length, width, player_speed]] * 4
                          + [[length, width, length, width, shoot_speed], [10, 10, 10, 10, 10]], dtype=np.float64),
            dtype=np.float64)
        self.ai_1_index, self.ai_2_index, self.opp_1_index, self.opp_2_index = 0, 1, 2, 3
        self.ball_index, self.ball_owner_array_index = 4, 5
        self._vec = FutbolVecEnv(1, device=device, seed=seed, env_id_offset=env_id, random_opp=random_opp,
                                 one_goal_end=one_goal_end, only_reward_goal=only_reward_goal, game_time=game_time,
                                 player_speed=player_speed, shoot_speed=shoot_speed, auto_reset=False,
                                 dtype=torch.float64)
        self._act = torch.zeros(1, dtype=torch.uint8, device=self._vec.device)
        self.obs = self.reset()

    def reset(self):
        self.obs = self._vec.reset().cpu().numpy().reshape(6, 5).copy()
        return self.obs

    def _next_observation(self):
        return self.obs

    def step(self, ai_action_type):
        if self.action_as_int:
            a = int(ai_action_type)
            if not 0 <= a < 16:
                raise ValueError("action must be in 0..15, got %r" % (ai_action_type,))
        else:
            a0, a1 = (int(x) for x in ai_action_type)
            if not (0 <= a0 < 4 and 0 <= a1 < 4):
                raise ValueError("action must be a pair of ints in 0..3, got %r" % (ai_action_type,))
            a = a0 * 4 + a1
        self._act.fill_(a)
        obs, reward, done, _ = self._vec.step(self._act)
        host = torch.cat([obs.reshape(-1), reward.reshape(-1), done.to(torch.float64)]).cpu().numpy()
        self.obs = host[:30].reshape(6, 5).copy()
        return self.obs, float(host[30]), bool(host[31]), {}

    # ---- hidden state the reference keeps as attributes (futbol_env.py:235-243) ----
    def _state(self):
        return self._vec.get_state()[0]

    @property
    def ball_owner(self):
        return BallOwner(int(self._state()["owner"]))

    @property
    def last_ball_owner(self):
        return BallOwner(int(self._state()["last_owner"]))

    @property
    def ai_score(self):
        return int(self._state()["ai_score"])

    @property
    def opp_score(self):
        return int(self._state()["opp_score"])

    @property
    def time(self):
        t = 0
        for _ in range(int(self._state()["ep_step"])):
            t += 0.1
        return t

    def render(self, mode="human", close=False):
        """Same picture as futbol_env.py:253-277 (needs matplotlib)."""
        import matplotlib.pyplot as plt
        _, ax = plt.subplots()
        ax.set_xlim(0, self.length)
        ax.set_ylim(0, self.width)
        for r, colour, label, size in ((0, "red", "ai", 12), (1, "red", "ai", 12), (2, "blue", "opp", 12),
                                       (3, "blue", "opp", 12), (4, "green", "ball", 8)):
            ax.plot(self.obs[r][0], self.obs[r][1], color=colour, marker="o", markersize=size, label=label)
        ax.legend()
        return ax

    def close(self):
        self._vec.close()
