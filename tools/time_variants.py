"""Rollout time per launch variant and batch size (v0, hard-coded opponents unless RANDOM_OPP=1): standard plain, standard
time-sliced (4), dense, automatic.    python tools/time_variants.py [K] [reps]   (SIZES=131072,262144,...)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_futbol_b200 import FutbolVecEnv

K = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ro = bool(int(os.environ.get("RANDOM_OPP", "0")))
sizes = [int(x) for x in os.environ.get("SIZES", "65536,131072,262144,524288,1048576").split(",")]
for n in sizes:
    acts = torch.randint(0, 16, (K, n), dtype=torch.uint8, device="cuda")
    for label, variant, slices in (("standard", 1, 1), ("sliced4", 1, 4), ("dense", 2, 0), ("auto", 0, 0)):
        env = FutbolVecEnv(n, seed=0, random_opp=ro)
        env.set_rollout_variant(variant)
        env.set_rollout_slices(slices)
        env.reset()
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
        print("n=%d %-9s %-26s %.3f ms/rollout, %.3e env-steps/s" % (n, label, env.rollout_kernel(K), ms, n * K / ms * 1e3), flush=True)
        env.close()
        del env
