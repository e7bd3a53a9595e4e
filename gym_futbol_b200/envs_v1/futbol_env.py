"""Drop-in for the reference's single-environment class ``gym_futbol.envs_v1.Futbol``.

Same constructor keywords (``width``, ``height``, ``total_time``, ``debug``, ``number_of_player``), spaces,
``reset() -> obs`` and ``step(action) -> (obs, reward, done, {})`` as gym_futbol/envs_v1/futbol_env.py:62-483;
the work is done by the CUDA v1 step kernel on a 1-env batch (float64 outputs, no auto-reset -- the reference
does not auto-reset either).  For throughput use ``gym_futbol_b200.FutbolV1VecEnv``.

Differences, all deliberate:
  * the rigid-body physics is this repository's restatement of the Chipmunk2D subset pymunk runs for the
    reference.  The game logic around it is pinned to traces of the reference's own Python (run over a pymunk
    stand-in, tests/golden/v1_golden.npz); the contact response itself could not be checked against the real
    library (pymunk cannot be run where this was built);
  * randomness (right-team actions, pass targets, out-of-bounds receiver, side after a goal) is the seeded
    counter-based Philox stream (``seed=``, ``env_id=``); the reference is unseeded;
  * ``width`` / ``height`` must keep their defaults (the reference's walls, goals and normalisation constants
    are module constants);
  * out-of-range actions raise ``ValueError`` (the reference prints and carries on).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import spaces
from ..vec_env import FutbolV1VecEnv

try:  # gym is optional
    import gym as _gym
    _Base = _gym.Env
except Exception:  # noqa: BLE001
    _Base = object

WIDTH, HEIGHT, TOTAL_TIME, NUMBER_OF_PLAYER = 105, 68, 30, 5


class Futbol(_Base):
    def __init__(self, width=WIDTH, height=HEIGHT, total_time=TOTAL_TIME, debug=False, number_of_player=NUMBER_OF_PLAYER,
                 seed=0, env_id=0, device="cuda:0"):
        if (width, height) != (WIDTH, HEIGHT):
            raise NotImplementedError("width/height are only supported at the reference defaults (105, 68)")
        self.width, self.height, self.total_time = width, height, total_time
        self.debug, self.number_of_player = debug, number_of_player
        self.action_space = spaces.MultiDiscrete([5, 5] * number_of_player)
        self.observation_space = spaces.Box(low=np.array([-1.0] * 4 * (1 + 2 * number_of_player), dtype=np.float32),
                                            high=np.array([1.0] * 4 * (1 + 2 * number_of_player), dtype=np.float32),
                                            dtype=np.float32)
        self._vec = FutbolV1VecEnv(1, number_of_player=number_of_player, device=device, seed=seed, env_id_offset=env_id,
                                   total_time=total_time, auto_reset=False, dtype=torch.float64)
        self._act = torch.zeros((1, 2 * number_of_player), dtype=torch.uint8, device=self._vec.device)
        self.observation = self._vec.obs.cpu().numpy().reshape(-1).copy()
        self.observation = self.reset()

    def reset(self):
        self.observation = self._vec.reset().cpu().numpy().reshape(-1).copy()
        return self.observation

    def _get_observation(self):
        return self.observation

    def random_action(self):
        return self.action_space.sample()

    def step(self, left_player_action):
        a = np.asarray(left_player_action, dtype=np.int64).reshape(-1)
        if a.shape != (2 * self.number_of_player,) or a.min() < 0 or a.max() > 4:
            raise ValueError("action must be %d ints in 0..4, got %r" % (2 * self.number_of_player, left_player_action))
        self._act.copy_(torch.from_numpy(a.astype(np.uint8)).reshape(1, -1))
        obs, reward, done, _ = self._vec.step(self._act)
        host = torch.cat([obs.reshape(-1), reward.reshape(-1), done.to(torch.float64)]).cpu().numpy()
        D = self._vec.obs_dim
        self.observation = host[:D].copy()
        return self.observation, float(host[D]), bool(host[D + 1]), {}

    # ---- state the reference keeps as attributes ----
    @property
    def current_time(self):
        t = 0
        for _ in range(int(self._vec.get_state()[0]["ep_step"])):
            t += 0.1
        return t

    @property
    def ball_owner_side(self):
        return "right" if int(self._vec.get_state()[0]["owner_side"]) else "left"

    def render(self):
        """Pitch picture from the current body positions (the reference draws the pymunk space, :236-243)."""
        import matplotlib.pyplot as plt
        st = self._vec.get_state()[0]["body"]
        N = self.number_of_player
        ax = plt.axes(xlim=(-5, self.width + 5), ylim=(-5, self.height + 5))
        ax.set_aspect("equal")
        ax.scatter(st[:N, 0], st[:N, 1], c="red", s=60)
        ax.scatter(st[N:2 * N, 0], st[N:2 * N, 1], c="blue", s=60)
        ax.scatter(st[2 * N, 0], st[2 * N, 1], c="green", s=30)
        return ax

    def close(self):
        self._vec.close()
