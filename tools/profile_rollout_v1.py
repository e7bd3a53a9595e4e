"""Small fixed v1 workload for ncu captures: PROF_N players per side, PROF_ENVS envs, K = 64, a few launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_futbol_b200 import FutbolV1VecEnv

N = int(os.environ.get("PROF_N", 2))
n = int(os.environ.get("PROF_ENVS", 1 << 18))
K = int(os.environ.get("PROF_K", 64))
env = FutbolV1VecEnv(n, number_of_player=N, seed=0)
env.reset()
for _ in range(int(os.environ.get("PROF_LAUNCHES", 4))):
    env.rollout(K, actions=torch.randint(0, 5, (K, n, 2 * N), dtype=torch.uint8, device="cuda"))     # a fresh table per launch
torch.cuda.synchronize()
print("ok", env.read_stats())
