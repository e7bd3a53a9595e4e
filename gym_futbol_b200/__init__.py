"""gym_futbol_b200: the gym-futbol environment step as hand-written CUDA for B200 (sm_100a).

Public surface (mirrors the reference package ``gym_futbol``):
  * ``gym_futbol_b200.envs.FutbolEnv``      -- drop-in single env  (reference: gym_futbol.envs.FutbolEnv)
  * ``gym_futbol_b200.FutbolVecEnv``        -- batched front end, torch CUDA tensors in/out
  * registry ids ``Futbol-v0`` ... are registered under the same names when ``gym`` is importable
    (reference: gym_futbol/__init__.py:3-28).
"""
from .vec_env import FutbolVecEnv  # noqa: F401
from ._lib import FutbolError  # noqa: F401

__all__ = ["FutbolVecEnv", "FutbolError", "register_envs"]


def register_envs():
    """Register the reference's ids with gym, pointing at the CUDA-backed classes."""
    from gym.envs.registration import register
    register(id="Futbol-v0", entry_point="gym_futbol_b200.envs:FutbolEnv")


try:  # pragma: no cover - gym is optional
    import gym  # noqa: F401
    register_envs()
except Exception:  # noqa: BLE001
    pass
