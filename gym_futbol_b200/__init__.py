"""gym_futbol_b200: the gym-futbol environment step as hand-written CUDA for B200 (sm_100a).

Public surface (mirrors the reference package ``gym_futbol``):
  * ``gym_futbol_b200.envs.FutbolEnv``      -- drop-in single env  (reference: gym_futbol.envs.FutbolEnv)
  * ``gym_futbol_b200.envs_v1.Futbol``      -- drop-in single env  (reference: gym_futbol.envs_v1.Futbol)
  * ``gym_futbol_b200.FutbolVecEnv`` / ``FutbolV1VecEnv`` -- batched front ends, torch CUDA tensors in/out
  * registry ids ``Futbol-v0`` ... are registered under the same names when ``gym`` is importable
    (reference: gym_futbol/__init__.py:3-28).
"""
from .vec_env import FutbolVecEnv, FutbolV1VecEnv  # noqa: F401
from ._lib import FutbolError  # noqa: F401

__all__ = ["FutbolVecEnv", "FutbolV1VecEnv", "FutbolError", "register_envs"]


def register_envs():
    """Register the reference's ids with gym, pointing at the CUDA-backed classes."""
    from gym.envs.registration import register
    register(id="Futbol-v0", entry_point="gym_futbol_b200.envs:FutbolEnv")
    # gym_futbol/__init__.py:12-28
    register(id="Futbol-v1", entry_point="gym_futbol_b200.envs_v1:Futbol", kwargs={"number_of_player": 10})
    register(id="Futbol2v2-v1", entry_point="gym_futbol_b200.envs_v1:Futbol", kwargs={"number_of_player": 2})
    register(id="Futbol5v5-v1", entry_point="gym_futbol_b200.envs_v1:Futbol", kwargs={"number_of_player": 5})


try:  # pragma: no cover - gym is optional
    import gym  # noqa: F401
    register_envs()
except Exception:  # noqa: BLE001
    pass
