"""Times the rollout on the per-rank batch sizes of the 2^20-env job (N = 8, 4, 2 ranks), plain vs time-sliced.

    python tools/time_rank_batch.py [K] [reps]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_futbol_b200 import FutbolVecEnv

K = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
sizes = [int(x) for x in os.environ.get("SIZES", "131072,262144,524288").split(",")]
for n in sizes:
    for slices in [int(x) for x in os.environ.get("SLICES", "1,0,2,3,4,6,8").split(",")]:
        env = FutbolVecEnv(n, seed=0, random_opp=False)
        env.set_rollout_slices(slices)
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
        print("n=%d slices=%d: %.3f ms/rollout, %.3e env-steps/s" % (n, slices, ms, n * K / ms * 1e3), flush=True)
        del env
