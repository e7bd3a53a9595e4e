"""Batched front end: n environments stepped by one CUDA launch, tensors stay on the device.

Mirrors the call surface stable-baselines' ``DummyVecEnv`` gives the reference's training loop
(colab_notebook.ipynb:818-823): ``reset() -> obs[n, ...]``, ``step(actions[n]) -> (obs, rewards,
dones, infos)`` with auto-reset on done and the terminal observation exposed separately, plus
``rollout(K)``: the caller's ``for t: step(a_t)`` loop fused into one launch.

All returned tensors are views of persistent device buffers owned by this object (zero-copy for
the policy); they are overwritten by the next ``step`` / ``rollout`` call.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib, spaces

OBS_DIM_V0 = 30


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class FutbolVecEnv:
    """v0 ``FutbolEnv`` (gym_futbol/envs/futbol_env.py) x ``num_envs`` on one GPU."""

    def __init__(self, num_envs, device="cuda:0", seed=0, env_id_offset=0, random_opp=True, one_goal_end=False,
                 only_reward_goal=False, game_time=40, player_speed=12, shoot_speed=20, auto_reset=True,
                 dtype=torch.float32):
        if dtype not in (torch.float32, torch.float64):
            raise ValueError("dtype must be torch.float32 or torch.float64")
        if int(num_envs) <= 0:
            raise ValueError("num_envs must be positive")
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise _lib.FutbolError("FutbolVecEnv needs a CUDA device; there is no CPU fallback")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.num_envs = int(num_envs)
        self.dtype = dtype
        self._dt = 1 if dtype == torch.float64 else 0
        self.cfg = _lib.FutbolConfig(_lib.ABI_VERSION, _lib.VARIANT_V0, self.num_envs, int(env_id_offset), int(seed), 2,
                                     int(bool(random_opp)), int(bool(one_goal_end)), int(bool(only_reward_goal)),
                                     int(bool(auto_reset)), int(shoot_speed), float(game_time), float(player_speed))
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.futbol_create(C.byref(self.cfg), C.byref(h)))
        self._h = h
        n = self.num_envs
        self.obs_dim, self.act_shape = OBS_DIM_V0, ()
        self._alloc()
        self.observation_space = spaces.Box(low=-np.inf, high=np.inf, shape=(OBS_DIM_V0,), dtype=np.float32)
        self.action_space = spaces.Discrete(16)

    def _alloc(self):
        n, h, D = self.num_envs, self._h, self.obs_dim
        self.state = torch.zeros(self.lib.futbol_state_bytes(h), dtype=torch.uint8, device=self.device)
        self.obs = torch.zeros((n, D), dtype=self.dtype, device=self.device)
        self.rewards = torch.zeros(n, dtype=self.dtype, device=self.device)
        self.dones = torch.zeros(n, dtype=torch.uint8, device=self.device)
        self.final_obs = torch.zeros((n, D), dtype=self.dtype, device=self.device)
        self.stats = torch.zeros(_lib.STATS_DTYPE.itemsize, dtype=torch.uint8, device=self.device)
        self._roll = {}
        self.episode_steps = self.lib.futbol_draw_limit_steps(h)
        # step() is called once per environment step from Python: everything that does not change between calls is
        # looked up once (the buffers above are persistent, so are their addresses)
        self._dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._step_shape = (n,) + tuple(self.act_shape)
        self._step_args = (self.state.data_ptr(), self.obs.data_ptr(), self.rewards.data_ptr(), self.dones.data_ptr(),
                           self.final_obs.data_ptr())
        self._step_info = {"terminal_observation": self.final_obs}

    # ------------------------------------------------------------------ plumbing
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, "_h", None):
            self.lib.futbol_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    @property
    def launch_count(self):
        return int(self.lib.futbol_launch_count(self._h))

    def rollout_slices(self, K):
        """Number of time slices ``rollout(K)`` will use on this device (1 = the plain kernel)."""
        return int(self.lib.futbol_rollout_slices(self._h, int(K)))

    ROLLOUT_KERNELS = ("v0_rollout_kernel", "v0_rollout_sliced_kernel", "v0_rollout_dense_kernel")

    def rollout_kernel(self, K):
        """Name of the kernel ``rollout(K)`` launches on this device."""
        if self.act_shape:
            return ("v1_rollout_kernel", "v1_rollout_sliced_kernel")[int(self.lib.futbol_rollout_kernel(self._h, int(K)))]
        return self.ROLLOUT_KERNELS[int(self.lib.futbol_rollout_kernel(self._h, int(K)))]

    def set_rollout_variant(self, variant):
        """0 = automatic, 1 = always the standard kernel (20 warps per SM), 2 = always the dense one (28).  Results are identical."""
        _lib.check(self.lib.futbol_set_rollout_variant(self._h, int(variant)))

    def set_rollout_slices(self, slices):
        """Time slicing of ``rollout`` (include/futbol_b200.h): 0 = automatic, 1 = off, n = n slices.  Results are identical."""
        _lib.check(self.lib.futbol_set_rollout_slices(self._h, int(slices)))

    def _actions(self, actions, shape):
        if (type(actions) is torch.Tensor and actions.dtype == torch.uint8 and actions.device == self.device
                and tuple(actions.shape) == shape and actions.is_contiguous()):
            if self.act_shape and actions.data_ptr() & 1:     # v1 reads (arrow, key) pairs as 16-bit words: realign a view
                return actions.clone()                        # that starts at an odd byte of a larger buffer
            return actions                                # the usual case: already what the kernel reads
        if not torch.is_tensor(actions):
            actions = torch.as_tensor(np.asarray(actions), device=self.device)
        if tuple(actions.shape) != shape:
            raise ValueError("actions must have shape %s, got %s" % (shape, tuple(actions.shape)))
        if actions.device != self.device:
            actions = actions.to(self.device, non_blocking=True)
        if actions.dtype != torch.uint8:
            actions = actions.to(torch.uint8)
        return actions.contiguous()

    # ------------------------------------------------------------------ gym-style API
    def reset(self, mask=None):
        """Reset all envs (or those with mask != 0); returns obs [n, 30]."""
        with torch.cuda.device(self.device):
            m = None if mask is None else self._actions(mask, (self.num_envs,))
            _lib.check(self.lib.futbol_reset(self._h, _ptr(self.state), _ptr(m), _ptr(self.obs), self._dt, self._stream()))
        return self.obs

    def step(self, actions, opp_actions=None, out=None):
        """actions: [n] ints in 0..15 (ai_1 = a // 4, ai_2 = a % 4).  opp_actions: None = the reference's own opponents;
        else the opponents' actions in the same format (self-play / learned opponents; v0 needs random_opp=True).
        out: None = this object's own buffers; else ``(obs, reward, done)`` -- contiguous CUDA tensors ``[n, obs_dim]`` /
        ``[n]`` of this env's dtype / ``[n]`` uint8, e.g. row ``t`` of a PPO rollout buffer -- that the kernel writes
        instead (no copy afterwards); an entry may be None (not written)."""
        if torch.cuda.current_device() != self._dev_index:
            with torch.cuda.device(self.device):
                return self.step(actions, opp_actions, out)
        a = self._actions(actions, self._step_shape)
        o = None if opp_actions is None else self._actions(opp_actions, self._step_shape).data_ptr()
        st, ob, rw, dn, fo = self._step_args
        if out is not None:
            obs_t, rew_t, done_t = out
            n = self.num_envs
            for t, shape, dt in ((obs_t, (n, self.obs_dim), self.dtype), (rew_t, (n,), self.dtype), (done_t, (n,), torch.uint8)):
                if t is not None and (tuple(t.shape) != shape or t.dtype != dt or t.device != self.device or not t.is_contiguous()):
                    raise ValueError("out tensors must be contiguous %s tensors of shape %s on %s" % (dt, shape, self.device))
            ob, rw, dn = (None if t is None else t.data_ptr() for t in (obs_t, rew_t, done_t))
        rc = self.lib.futbol_step_vs(self._h, st, a.data_ptr(), o, ob, rw, dn, fo, self._dt,
                                     torch.cuda.current_stream().cuda_stream)
        if rc != 0:
            _lib.check(rc)
        if out is not None:
            return out[0], out[1], out[2], self._step_info
        return self.obs, self.rewards, self.dones, self._step_info

    def rollout(self, K, actions=None, obs=True, reward=True, done=True, out=None, opp_actions=None):
        """K fused steps.  actions: uint8 [K, n] or None (uniform random actions drawn in-kernel).

        Returns (obs [K, n, 30] f32, reward [K, n] f32, done [K, n] u8).  By default the buffers are owned by
        this object and cached per K; ``out=(obs, reward, done)`` writes into caller-provided contiguous CUDA
        tensors of those shapes and dtypes instead (e.g. slices of a PPO rollout buffer); an entry may be None.
        ``opp_actions``: uint8 [K, n(, 2N)] the opponents' actions supplied by the caller (open-loop self-play replay).
        """
        K = int(K)
        n = self.num_envs
        if out is not None:
            o, r, d = out
            for t, shape, dt in ((o, (K, n, self.obs_dim), torch.float32), (r, (K, n), torch.float32), (d, (K, n), torch.uint8)):
                if t is not None and (tuple(t.shape) != shape or t.dtype != dt or t.device != self.device or not t.is_contiguous()):
                    raise ValueError("out tensors must be contiguous %s tensors of shape %s on %s" % (dt, shape, self.device))
        else:
            buf = self._roll.get(K)
            if buf is None:
                buf = (torch.empty((K, n, self.obs_dim), dtype=torch.float32, device=self.device),
                       torch.empty((K, n), dtype=torch.float32, device=self.device),
                       torch.empty((K, n), dtype=torch.uint8, device=self.device))
                self._roll[K] = buf
            o, r, d = buf
            o, r, d = (o if obs else None), (r if reward else None), (d if done else None)
        with torch.cuda.device(self.device):
            a = None if actions is None else self._actions(actions, (K, n) + self.act_shape)
            oa = None if opp_actions is None else self._actions(opp_actions, (K, n) + self.act_shape)
            _lib.check(self.lib.futbol_rollout_vs(self._h, _ptr(self.state), K, _ptr(a), _ptr(oa), _ptr(o), _ptr(r), _ptr(d),
                                                  _ptr(self.stats), self._stream()))
        return o, r, d

    # ------------------------------------------------------------------ state / statistics
    STATE_DTYPE = _lib.V0_ENV_STATE

    def get_state(self):
        """Env state as a numpy structured array (``_lib.V0_ENV_STATE`` / ``V1_ENV_STATE``); synchronises."""
        aos = torch.empty(self.num_envs * self.STATE_DTYPE.itemsize, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.futbol_get_state(self._h, _ptr(self.state), _ptr(aos), self._stream()))
        return aos.cpu().numpy().view(self.STATE_DTYPE).copy()

    def set_state(self, records):
        records = np.ascontiguousarray(records, dtype=_lib.V0_ENV_STATE)
        if records.shape != (self.num_envs,):
            raise ValueError("need %d state records" % self.num_envs)
        # The kernel's fp64 division / square root are the correctly rounded fast-path sequences without the
        # range guard (csrc/ieee_fast.cuh); they are exact for pitch-scale operands.  Anything a simulation can
        # reach is many orders of magnitude inside these bounds; refuse the rest instead of computing on it.
        rows = records["rows"]
        mag = np.abs(rows)
        if not np.isfinite(rows).all() or (mag > 1e6).any() or ((mag != 0) & (mag < 1e-60)).any():
            raise ValueError("state values must be finite, at most 1e6 in magnitude and either zero or at least 1e-60")
        if (records["owner"] > 4).any() or (records["last_owner"] > 4).any() or (records["ep_step"] < 0).any():
            raise ValueError("owner / last_owner must be 0..4 and ep_step non-negative")
        aos = torch.from_numpy(records.view(np.uint8).copy()).to(self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.futbol_set_state(self._h, _ptr(self.state), _ptr(aos), self._stream()))
            torch.cuda.current_stream(self.device).synchronize()

    def read_stats(self, clear=False):
        """Accumulated rollout statistics (sums over envs and steps); synchronises."""
        rec = self.stats.cpu().numpy().view(_lib.STATS_DTYPE)[0]
        out = {k: (float(rec[k]) if k == "reward_sum" else int(rec[k])) for k in
               ("reward_sum", "env_steps", "episodes", "goals_ai", "goals_opp", "out_of_field")}
        out["contacts"], out["contacts_dropped"] = int(rec["reserved"][0]), int(rec["reserved"][1])
        if clear:
            self.stats.zero_()
        return out


class FutbolV1VecEnv(FutbolVecEnv):
    """v1 ``Futbol`` (gym_futbol/envs_v1/futbol_env.py, N-vs-N with rigid-body contacts) x ``num_envs`` on one GPU.

    ``step(actions)``: actions uint8 ``[n, 2N]`` = (arrow key, action key) per left-team player, i.e. the
    reference's ``MultiDiscrete([5, 5] * N)`` (:78-79); the right team draws uniform random actions (:429).
    Observations are the normalised ``4 + 8N`` vector (:154-180).  The physics restates the Chipmunk2D subset
    pymunk runs for the reference: the game logic is pinned to traces of the reference's own Python, the contact physics
    at the pymunk boundary is a restatement that cannot be checked against the real library here.
    """

    STATE_DTYPE = _lib.V1_ENV_STATE

    def __init__(self, num_envs, number_of_player=2, device="cuda:0", seed=0, env_id_offset=0, total_time=30,
                 auto_reset=True, dtype=torch.float32):
        if dtype not in (torch.float32, torch.float64):
            raise ValueError("dtype must be torch.float32 or torch.float64")
        if int(num_envs) <= 0:
            raise ValueError("num_envs must be positive")
        if not 1 <= int(number_of_player) <= 10:
            raise ValueError("number_of_player must be 1..10")
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise _lib.FutbolError("FutbolV1VecEnv needs a CUDA device; there is no CPU fallback")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.num_envs, self.number_of_player = int(num_envs), int(number_of_player)
        self.dtype = dtype
        self._dt = 1 if dtype == torch.float64 else 0
        self.cfg = _lib.FutbolConfig(_lib.ABI_VERSION, _lib.VARIANT_V1, self.num_envs, int(env_id_offset), int(seed),
                                     self.number_of_player, 0, 0, 0, int(bool(auto_reset)), 20, float(total_time), 12.0)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.futbol_create(C.byref(self.cfg), C.byref(h)))
        self._h = h
        self.obs_dim, self.act_shape = 4 + 8 * self.number_of_player, (2 * self.number_of_player,)
        self.STATE_DTYPE = _lib.v1_env_state_dtype(self.number_of_player)
        assert self.STATE_DTYPE.itemsize == self.lib.futbol_env_state_bytes(h)
        self._alloc()
        self.observation_space = spaces.Box(low=-1.0, high=1.0, shape=(self.obs_dim,), dtype=np.float32)
        self.action_space = spaces.MultiDiscrete([5, 5] * self.number_of_player)

    def set_state(self, records):
        """Restores every env from records of ``get_state()`` (bodies, counters and the arbiter cache); synchronises."""
        records = np.ascontiguousarray(records, dtype=self.STATE_DTYPE)
        if records.shape != (self.num_envs,):
            raise ValueError("need %d state records" % self.num_envs)
        B = 2 * self.number_of_player + 1
        for vals in (records["body"][:, :B], records["jn"]):       # same domain rule as v0 (guard-free division / square root)
            mag = np.abs(vals)
            if not np.isfinite(vals).all() or (mag > 1e6).any() or ((mag != 0) & (mag < 1e-60)).any():
                raise ValueError("state values must be finite, at most 1e6 in magnitude and either zero or at least 1e-60")
        if (records["owner_side"] > 1).any() or (records["ep_step"] < 0).any():
            raise ValueError("owner_side must be 0 / 1 and ep_step non-negative")
        if (records["stamp"] < 8).any() or (records["last"] > records["stamp"][:, None]).any():
            raise ValueError("stamp counts from 8 and no pair can have touched after it")
        aos = torch.from_numpy(records.view(np.uint8).copy()).to(self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.futbol_set_state(self._h, _ptr(self.state), _ptr(aos), self._stream()))
            torch.cuda.current_stream(self.device).synchronize()
