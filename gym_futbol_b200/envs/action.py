"""Action ids of the v0 env (same names and values as the reference enum, gym_futbol/envs/action.py:3-6)."""
import enum


class Action(enum.Enum):
    run = 0
    intercept = 1
    shoot = 2
    assist = 3
