"""Per-step API inside CUDA graphs at small batch sizes: us per step (kernel + launch gap), for launch-geometry variants.
    [FUTBOL_B200_LIB=libfutbol_b200_et32.so] python tools/time_step_graph.py [n_envs ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_futbol_b200 import FutbolVecEnv

sizes = [int(x) for x in sys.argv[1:]] or [4096]
for n in sizes:
    for random_opp in (False,):
        env = FutbolVecEnv(n, seed=0, random_opp=random_opp)
        env.reset()
        acts = torch.randint(0, 16, (100, n), dtype=torch.uint8, device="cuda")
        for t in range(20):
            env.step(acts[t])
        g, side = torch.cuda.CUDAGraph(), torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                for t in range(100):
                    env.step(acts[t])
        torch.cuda.current_stream().wait_stream(side)
        g.replay(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) / 20 / 100 * 1e3
        print("%s n=%d random_opp=%d: %.2f us per graphed step, %.3e env-steps/s" % (os.environ.get("FUTBOL_B200_LIB", "default"), n, random_opp, us, n / (us * 1e-6)), flush=True)
        del env
