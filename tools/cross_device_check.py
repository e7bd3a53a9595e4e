"""Two-GPU check: envs created on a device that is not the current one give the same results (step, rollout, v1)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from gym_futbol_b200 import FutbolVecEnv, FutbolV1VecEnv
assert torch.cuda.device_count() >= 2
torch.cuda.set_device(0)
a = FutbolVecEnv(4096, device="cuda:1", seed=3, random_opp=False)
b = FutbolVecEnv(4096, device="cuda:0", seed=3, random_opp=False)
a.reset(); b.reset()
acts = torch.randint(0, 16, (50, 4096), dtype=torch.uint8)
for t in range(50):
    oa, ra, da, _ = a.step(acts[t].to("cuda:1"))
    ob, rb, db, _ = b.step(acts[t].to("cuda:0"))
assert oa.device.index == 1 and torch.equal(oa.cpu(), ob.cpu()) and torch.equal(ra.cpu(), rb.cpu())
o1, r1, d1 = a.rollout(32, actions=acts[:32].to("cuda:1")); o0, r0, d0 = b.rollout(32, actions=acts[:32].to("cuda:0"))
assert torch.equal(o1.cpu(), o0.cpu()) and torch.equal(d1.cpu(), d0.cpu())
v = FutbolV1VecEnv(512, number_of_player=2, device="cuda:1", seed=1); w = FutbolV1VecEnv(512, number_of_player=2, device="cuda:0", seed=1)
v.reset(); w.reset()
x, _, _ = v.rollout(40); y, _, _ = w.rollout(40)
assert torch.equal(x.cpu(), y.cpu())
print("cross-device ok")
