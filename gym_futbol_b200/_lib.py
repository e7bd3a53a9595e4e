"""ctypes binding of the C ABI (include/futbol_b200.h).

There is no CPU fallback: if the CUDA library is missing or no device is present, creating
an environment raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

ABI_VERSION = 2
VARIANT_V0, VARIANT_V1 = 0, 1


class FutbolConfig(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("variant", C.c_int32), ("n_envs", C.c_int32),
                ("env_id_offset", C.c_uint32), ("seed", C.c_uint64), ("n_players", C.c_int32),
                ("random_opp", C.c_int32), ("one_goal_end", C.c_int32), ("only_reward_goal", C.c_int32),
                ("auto_reset", C.c_int32), ("shoot_speed", C.c_int32), ("game_time", C.c_double),
                ("player_speed", C.c_double)]


class FutbolStats(C.Structure):
    _fields_ = [("reward_sum", C.c_double), ("env_steps", C.c_uint64), ("episodes", C.c_uint64),
                ("goals_ai", C.c_uint64), ("goals_opp", C.c_uint64), ("out_of_field", C.c_uint64),
                ("reserved", C.c_uint64 * 2)]


# numpy mirror of FutbolV0EnvState
V0_ENV_STATE = np.dtype([("rows", np.float64, (5, 5)), ("t_total", np.uint64), ("ep_step", np.int32),
                         ("ai_score", np.int32), ("opp_score", np.int32), ("owner", np.uint8),
                         ("last_owner", np.uint8), ("flags", np.uint8), ("pad_", np.uint8)], align=True)
V1_ENV_STATE = np.dtype([("body", np.float64, (21, 6)), ("t_total", np.uint64), ("stamp", np.uint32), ("ep_step", np.int32),
                         ("owner_side", np.uint8), ("flags", np.uint8), ("pad_", np.uint8, (6,))], align=True)   # the header


def v1_env_state_dtype(n_players):
    """numpy mirror of one whole v1 record: the FutbolV1EnvState header + the env's arbiter cache (jn[P], last[P])."""
    B = 2 * int(n_players) + 1
    P = B * (B - 1) // 2 + 12 * B
    fields = [(name, V1_ENV_STATE.fields[name][0]) for name in V1_ENV_STATE.names]
    fields += [("jn", np.float64, (P,)), ("last", np.uint32, (P,))]
    if P & 1:
        fields.append(("pad2_", np.uint32))
    return np.dtype(fields, align=True)
STATS_DTYPE = np.dtype([("reward_sum", np.float64), ("env_steps", np.uint64), ("episodes", np.uint64),
                        ("goals_ai", np.uint64), ("goals_opp", np.uint64), ("out_of_field", np.uint64),
                        ("reserved", np.uint64, (2,))])

EXPORTS = ("futbol_create", "futbol_destroy", "futbol_last_error", "futbol_abi_version", "futbol_state_bytes",
           "futbol_obs_dim", "futbol_act_dim", "futbol_draw_limit_steps", "futbol_reset", "futbol_step",
           "futbol_rollout", "futbol_env_state_bytes", "futbol_get_state", "futbol_set_state",
           "futbol_launch_count", "futbol_gae", "futbol_selftest_arith", "futbol_step_vs", "futbol_rollout_vs",
           "futbol_set_rollout_slices", "futbol_rollout_slices", "futbol_rollout_kernel",
           "futbol_set_rollout_variant", "futbol_gather_minibatch", "futbol_sample_actions")

_lib = None


class FutbolError(RuntimeError):
    pass


def lib_path():
    return _build.LIB


def load():
    """Load (building first if the in-tree .so is missing or stale and nvcc exists)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    variant = os.environ.get("FUTBOL_B200_LIB")     # tuning aid: an alternative build of the same sources (tools/)
    if variant:
        path = variant if os.path.isabs(variant) else os.path.join(os.path.dirname(path), variant)
        if not os.path.exists(path):
            raise FutbolError("FUTBOL_B200_LIB=%s does not exist" % path)
    elif _build.is_stale():
        try:
            _build.build_extension()
        except Exception as exc:  # no nvcc on this box and no prebuilt library
            if not os.path.exists(path):
                raise FutbolError("libfutbol_b200.so is not built and cannot be built here (%s); "
                                  "there is no CPU fallback" % exc) from exc
            import warnings
            warnings.warn("libfutbol_b200.so is OLDER than its sources and could not be rebuilt (%s): loading the stale "
                          "library; only its ABI version is checked" % (str(exc).splitlines()[0],), RuntimeWarning, stacklevel=2)
    L = C.CDLL(path)
    vp = C.c_void_p
    L.futbol_create.restype = C.c_int
    L.futbol_create.argtypes = [C.POINTER(FutbolConfig), C.POINTER(vp)]
    L.futbol_destroy.restype = C.c_int
    L.futbol_destroy.argtypes = [vp]
    L.futbol_last_error.restype = C.c_char_p
    L.futbol_abi_version.restype = C.c_int
    for name in ("futbol_state_bytes", "futbol_env_state_bytes"):
        getattr(L, name).restype = C.c_size_t
        getattr(L, name).argtypes = [vp]
    for name in ("futbol_obs_dim", "futbol_act_dim", "futbol_draw_limit_steps"):
        getattr(L, name).restype = C.c_int
        getattr(L, name).argtypes = [vp]
    L.futbol_set_rollout_slices.restype = C.c_int
    L.futbol_set_rollout_slices.argtypes = [vp, C.c_int]
    L.futbol_rollout_slices.restype = C.c_int
    L.futbol_rollout_slices.argtypes = [vp, C.c_int]
    L.futbol_rollout_kernel.restype = C.c_int
    L.futbol_rollout_kernel.argtypes = [vp, C.c_int]
    L.futbol_set_rollout_variant.restype = C.c_int
    L.futbol_set_rollout_variant.argtypes = [vp, C.c_int]
    L.futbol_sample_actions.restype = C.c_int
    L.futbol_sample_actions.argtypes = [vp, C.c_int, C.c_int64, C.c_int, C.c_uint64, vp, C.c_uint64, vp, vp, vp]
    L.futbol_gather_minibatch.restype = C.c_int
    L.futbol_gather_minibatch.argtypes = [vp, C.c_int64, C.c_int64, vp, C.c_int, vp] + [vp] * 10 + [vp, vp]
    L.futbol_launch_count.restype = C.c_uint64
    L.futbol_launch_count.argtypes = [vp]
    L.futbol_reset.restype = C.c_int
    L.futbol_reset.argtypes = [vp, vp, vp, vp, C.c_int, vp]
    L.futbol_step.restype = C.c_int
    L.futbol_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, C.c_int, vp]
    L.futbol_step_vs.restype = C.c_int
    L.futbol_step_vs.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, C.c_int, vp]
    L.futbol_rollout_vs.restype = C.c_int
    L.futbol_rollout_vs.argtypes = [vp, vp, C.c_int, vp, vp, vp, vp, vp, vp, vp]
    L.futbol_rollout.restype = C.c_int
    L.futbol_rollout.argtypes = [vp, vp, C.c_int, vp, vp, vp, vp, vp, vp]
    L.futbol_get_state.restype = C.c_int
    L.futbol_get_state.argtypes = [vp, vp, vp, vp]
    L.futbol_set_state.restype = C.c_int
    L.futbol_set_state.argtypes = [vp, vp, vp, vp]
    L.futbol_selftest_arith.restype = C.c_int
    L.futbol_selftest_arith.argtypes = [vp, vp, vp, C.c_size_t, vp]
    L.futbol_gae.restype = C.c_int
    L.futbol_gae.argtypes = [vp, vp, vp, C.c_float, C.c_float, vp, vp, C.c_int, C.c_int, vp]
    if L.futbol_abi_version() != ABI_VERSION:
        raise FutbolError("libfutbol_b200.so ABI version mismatch")
    assert C.sizeof(FutbolStats) == STATS_DTYPE.itemsize == 64
    _lib = L
    return L


def check(code):
    if code != 0:
        raise FutbolError("futbol_b200 error %d: %s" % (code, load().futbol_last_error().decode()))
