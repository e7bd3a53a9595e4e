"""Per-rank timing of the bench's end-to-end loop pieces.  torchrun --nproc-per-node N tools/e2e_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from gym_futbol_b200 import FutbolVecEnv

rank = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
n = (1 << 20) // world; K = 64
env = FutbolVecEnv(n, device=dev, seed=0, env_id_offset=rank * n, random_opp=False); env.reset()
stream = torch.cuda.current_stream(dev); copy_stream = torch.cuda.Stream(dev)
t0 = time.perf_counter()
h_obs = torch.empty((K, n, 30), dtype=torch.float32).pin_memory()
h_rew = torch.empty((K, n), dtype=torch.float32).pin_memory()
h_done = torch.empty((K, n), dtype=torch.uint8).pin_memory()
h_acts = torch.randint(0, 16, (K, n), dtype=torch.uint8).pin_memory()
print("rank %d: pinned alloc %.2f s, is_pinned %s" % (rank, time.perf_counter() - t0, h_obs.is_pinned()), flush=True)
for chunks in (8, 2, 1):
    Kc = K // chunks
    bufs = [(torch.empty((Kc, n), dtype=torch.uint8, device=dev), torch.empty((Kc, n, 30), device=dev), torch.empty((Kc, n), device=dev),
             torch.empty((Kc, n), dtype=torch.uint8, device=dev)) for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]; ev_free = [torch.cuda.Event() for _ in range(2)]
    for mode in ("sim+copy", "copy only", "sim only"):
        def step():
            for c in range(chunks):
                b = c & 1; da, do, dr, dd = bufs[b]; ks = slice(c * Kc, (c + 1) * Kc)
                stream.wait_event(ev_free[b])
                if mode != "copy only":
                    da.copy_(h_acts[ks], non_blocking=True)
                    env.rollout(Kc, actions=da, out=(do, dr, dd))
                ev_done[b].record(stream)
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(ev_done[b])
                    if mode != "sim only":
                        h_obs[ks].copy_(do, non_blocking=True); h_rew[ks].copy_(dr, non_blocking=True); h_done[ks].copy_(dd, non_blocking=True)
                    ev_free[b].record(copy_stream)
            copy_stream.synchronize()
        for b in range(2): ev_free[b].record(copy_stream)
        step(); torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        for _ in range(3): step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 3
        print("rank %d chunks %d %-10s: %.1f ms/step  (%.1f GB/s D2H, %.3e env-steps/s/rank)" % (rank, chunks, mode, dt * 1e3, n * K * 125 / dt / 1e9, n * K / dt), flush=True)
        dist.barrier()
dist.destroy_process_group()
