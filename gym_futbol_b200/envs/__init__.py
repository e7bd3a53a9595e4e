from .futbol_env import FutbolEnv  # noqa: F401  (reference: gym_futbol/envs/__init__.py:1)
