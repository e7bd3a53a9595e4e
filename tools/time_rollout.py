"""Times the fused rollout kernel alone (CUDA events, warm, outputs larger than L2): env-steps/s.

    [FUTBOL_B200_LIB=libfutbol_b200_mb4.so] python tools/time_rollout.py [n_envs] [K] [reps]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_futbol_b200 import FutbolVecEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
K = int(sys.argv[2]) if len(sys.argv) > 2 else 64
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
for random_opp in (False, True):
    env = FutbolVecEnv(n, seed=0, random_opp=random_opp)
    env.reset()
    acts = torch.randint(0, 16, (K, n), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        env.rollout(K, actions=acts)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        env.rollout(K, actions=acts)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print("%s random_opp=%d n=%d K=%d: %.3f ms/rollout, %.3e env-steps/s" % (
        os.environ.get("FUTBOL_B200_LIB", "default"), random_opp, n, K, ms, n * K / ms * 1e3), flush=True)
    del env
