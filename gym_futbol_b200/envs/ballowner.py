"""Ball-owner ids of the v0 env (reference: gym_futbol/envs/ballowner.py:3-7); the value is the obs row."""
import enum


class BallOwner(enum.Enum):
    AI_1 = 0
    AI_2 = 1
    OPP_1 = 2
    OPP_2 = 3
    NOONE = 4
