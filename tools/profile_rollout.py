"""Small fixed workload for ncu captures of the rollout kernel: 2^18 envs, K=64, a few launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_futbol_b200 import FutbolVecEnv

n = int(os.environ.get("PROF_ENVS", 1 << 18))
K = int(os.environ.get("PROF_K", 64))
random_opp = bool(int(os.environ.get("PROF_RANDOM_OPP", "0")))
env = FutbolVecEnv(n, seed=0, random_opp=random_opp)
env.reset()
acts = torch.randint(0, 16, (K, n), dtype=torch.uint8, device="cuda")
for _ in range(int(os.environ.get("PROF_LAUNCHES", 4))):
    env.rollout(K, actions=acts)
torch.cuda.synchronize()
print("ok", env.read_stats())
