"""A/B of two builds of the library on the same box through the bare C ABI (works across ABI versions):
    python tools/ab_rank_batch.py libA.so libB.so ... [n_envs ...]      -- v0, hard-coded opponents, K = 64, automatic slicing"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_futbol_b200._lib import FutbolConfig

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gym_futbol_b200", "csrc")
libs = [a for a in sys.argv[1:] if not a.isdigit()]
sizes = [int(a) for a in sys.argv[1:] if a.isdigit()] or [131072, 1048576]
K, reps = 64, 30
vp = C.c_void_p
for n in sizes:
    acts = torch.randint(0, 16, (K, n), dtype=torch.uint8, device="cuda")
    obs = torch.empty((K, n, 30), dtype=torch.float32, device="cuda")
    rew = torch.empty((K, n), dtype=torch.float32, device="cuda")
    done = torch.empty((K, n), dtype=torch.uint8, device="cuda")
    for rnd in range(2):
        for name in libs:
            L = C.CDLL(os.path.join(CSRC, name))
            L.futbol_state_bytes.restype = C.c_size_t
            L.futbol_state_bytes.argtypes = [vp]
            L.futbol_reset.argtypes = [vp, vp, vp, vp, C.c_int, vp]
            L.futbol_rollout.argtypes = [vp, vp, C.c_int, vp, vp, vp, vp, vp, vp]
            cfg = FutbolConfig(L.futbol_abi_version(), 0, n, 0, 0, 2, 0, 0, 0, 1, 20, 40.0, 12.0)
            h = vp()
            assert L.futbol_create(C.byref(cfg), C.byref(h)) == 0
            state = torch.zeros(L.futbol_state_bytes(h), dtype=torch.uint8, device="cuda")
            st = vp(torch.cuda.current_stream().cuda_stream)
            p = lambda t: vp(t.data_ptr())
            assert L.futbol_reset(h, p(state), None, None, 0, st) == 0
            for _ in range(3):
                assert L.futbol_rollout(h, p(state), K, p(acts), p(obs), p(rew), p(done), None, st) == 0
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                L.futbol_rollout(h, p(state), K, p(acts), p(obs), p(rew), p(done), None, st)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / reps
            print("n=%d %-26s %.3f ms/rollout, %.3e env-steps/s" % (n, name, ms, n * K / ms * 1e3), flush=True)
            L.futbol_destroy(h)
