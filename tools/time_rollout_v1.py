"""Times the fused v1 rollout kernel alone (CUDA events, warm): env-steps/s.
    python tools/time_rollout_v1.py [number_of_player] [n_envs] [K] [reps]        (SLICES=n: futbol_set_rollout_slices, default automatic)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_futbol_b200 import FutbolV1VecEnv

N = int(sys.argv[1]) if len(sys.argv) > 1 else 5
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 18
K = int(sys.argv[3]) if len(sys.argv) > 3 else 64
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
env = FutbolV1VecEnv(n, number_of_player=N, seed=0)
env.set_rollout_slices(int(os.environ.get("SLICES", "0")))
env.reset()
# distinct action tables, cycled (TABLES=1: one table reused every launch -- each player then repeats its K actions for ever,
# players pile up at the walls and a step has ~30 % more contacts; the numbers in profiles/ before r2i were taken that way)
T = int(os.environ.get("TABLES", "8"))
tables = [torch.randint(0, 5, (K, n, 2 * N), dtype=torch.uint8, device="cuda") for _ in range(T)]
turn = 0
for _ in range(2):
    env.rollout(K, actions=tables[turn % T]); turn += 1
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    env.rollout(K, actions=tables[turn % T]); turn += 1
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
st = env.read_stats()
print("v1 %dv%d n=%d K=%d slices=%d: %.3f ms/rollout, %.3e env-steps/s; contacts/env-step %.3f, dropped %d, goals %d/%d, outs %d" % (
    N, N, n, K, env.rollout_slices(K), ms, n * K / ms * 1e3, st["contacts"] / max(1, st["env_steps"]), st["contacts_dropped"], st["goals_ai"], st["goals_opp"],
    st["out_of_field"]), flush=True)
