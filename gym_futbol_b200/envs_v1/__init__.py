from .futbol_env import Futbol  # noqa: F401  (reference: gym_futbol/envs_v1/__init__.py:1)
