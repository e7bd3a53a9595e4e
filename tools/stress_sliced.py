"""Stress of the time-sliced rollout's work queue: thousands of launches at sizes where units wait on their predecessors
(few env-blocks) and at the rank size, final state compared with the plain launch.  Run under `timeout`."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_futbol_b200 import FutbolVecEnv, FutbolV1VecEnv

for n, K, slices, launches in ((4096, 64, 16, 3000), (131072, 64, 0, 1500), (1000, 32, 32, 3000), (300000, 48, 5, 300)):
    a = FutbolVecEnv(n, seed=13, random_opp=False); a.set_rollout_slices(slices); a.reset()
    b = FutbolVecEnv(n, seed=13, random_opp=False); b.set_rollout_slices(1); b.reset()
    t0 = time.time()
    for _ in range(launches):
        a.rollout(K, obs=False, reward=False, done=False)
    torch.cuda.synchronize(); ta = time.time() - t0
    for _ in range(launches):
        b.rollout(K, obs=False, reward=False, done=False)
    torch.cuda.synchronize()
    same = a.get_state().tobytes() == b.get_state().tobytes()
    print("n=%d K=%d slices=%d launches=%d: %.2f s, final state identical to the plain launch: %s" % (n, K, slices, launches, ta, same), flush=True)
    assert same
# v1: warps of 32 envs as units (csrc/v1_kernels.cu, v1_rollout_sliced_kernel)
for N, n, K, slices, launches in ((2, 4096, 32, 8, 1500), (5, 1000, 16, 16, 1500), (5, 80000, 32, 0, 200), (10, 3000, 16, 4, 300), (1, 100000, 64, 0, 300)):
    a = FutbolV1VecEnv(n, number_of_player=N, seed=13); a.set_rollout_slices(slices); a.reset()
    b = FutbolV1VecEnv(n, number_of_player=N, seed=13); b.set_rollout_slices(1); b.reset()
    t0 = time.time()
    for _ in range(launches):
        a.rollout(K, obs=False, reward=False, done=False)
    torch.cuda.synchronize(); ta = time.time() - t0
    for _ in range(launches):
        b.rollout(K, obs=False, reward=False, done=False)
    torch.cuda.synchronize()
    same = a.get_state().tobytes() == b.get_state().tobytes()
    print("v1 %dv%d n=%d K=%d slices=%d (%s) launches=%d: %.2f s, final state identical to the plain launch: %s" % (
        N, N, n, K, slices, a.rollout_kernel(K), launches, ta, same), flush=True)
    assert same
print("stress ok")
