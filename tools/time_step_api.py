"""BASELINE.json configs[1]: v0 2v2 vs hard-coded opponents, 4096 envs on one GPU, 100 warm-up + 1000 timed steps
through the PER-STEP API (one launch per step, state round-trips HBM), and the same 1000 steps as fused rollouts.
    python tools/time_step_api.py [n_envs]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_futbol_b200 import FutbolVecEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = FutbolVecEnv(n, seed=0, random_opp=False)
env.reset()
acts = torch.randint(0, 16, (1000, n), dtype=torch.uint8, device="cuda")
for t in range(100):
    env.step(acts[t])
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
a.record()
for t in range(1000):
    env.step(acts[t])
b.record()
torch.cuda.synchronize()
wall = time.perf_counter() - t0
print("per-step API   n=%d: %.1f us/step on the device stream (%.1f us wall), %.3e env-steps/s" % (n, a.elapsed_time(b), wall * 1e3, n * 1000 / wall))
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    with torch.cuda.graph(g, stream=s):
        for t in range(100):
            env.step(acts[t])
torch.cuda.current_stream().wait_stream(s)
g.replay(); torch.cuda.synchronize()
a.record()
for _ in range(10):
    g.replay()
b.record()
torch.cuda.synchronize()
print("per-step API, 100 steps per CUDA graph   n=%d: %.2f us/step, %.3e env-steps/s" % (n, a.elapsed_time(b), n * 1000 / (a.elapsed_time(b) * 1e-3)))
for K in (1000,):
    env.rollout(K, actions=acts[:K])
    torch.cuda.synchronize()
    a.record()
    env.rollout(K, actions=acts[:K])
    b.record()
    torch.cuda.synchronize()
    print("fused rollout  n=%d K=%d: %.3f ms, %.3e env-steps/s" % (n, K, a.elapsed_time(b), n * K / (a.elapsed_time(b) * 1e-3)))
