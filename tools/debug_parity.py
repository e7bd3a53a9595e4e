"""Debug helper (GPU): find the first env/step/field where the CUDA per-step API leaves the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gym_futbol_b200 import FutbolVecEnv
from oracle import philox
from oracle.v0 import OracleV0

def run(n, steps, seed, off, random_opp, game_time, **flags):
    env = FutbolVecEnv(n, seed=seed, env_id_offset=off, random_opp=random_opp, game_time=game_time, dtype=torch.float64, **flags)
    orc = OracleV0(n, seed=seed, env_id0=off, random_opp=random_opp, game_time=game_time, arith=0, **flags)
    env.reset()
    acts = philox.actions_table(seed, np.arange(off, off + n), 0, steps)
    prev = None
    for t in range(steps):
        before = orc.envs.copy()
        want = orc.rollout(1, actions=acts[t:t+1], autoreset=2)
        obs, rew, done, info = env.step(acts[t])
        o = obs.cpu().numpy(); w = want["obs"][0]
        err = np.abs(o - w) / np.maximum(1, np.abs(w))
        bad = np.argwhere(err > 1e-9)
        if len(bad) or not np.array_equal(done.cpu().numpy(), want["done"][0]):
            i = bad[0][0] if len(bad) else int(np.argwhere(done.cpu().numpy() != want["done"][0])[0][0])
            print("MISMATCH step", t, "env", i, "global", off + i, "action", acts[t, i], (acts[t,i]//4, acts[t,i]%4))
            print("before (oracle): owner", before["owner"][i], "last", before["last_owner"][i], "time", before["time"][i])
            print(before["obs"][i])
            print("oracle after: owner", orc.envs["owner"][i], "last", orc.envs["last_owner"][i], "draws", orc.envs["step_draws"][i], "flags", orc.envs["flags"][i])
            print(w.reshape(6, 5))
            print("gpu after:"); print(o[i].reshape(6, 5))
            st = env.get_state()[i]
            print("gpu owner", st["owner"], "last", st["last_owner"], "flags", st["flags"], "rew", rew[i].item(), want["reward"][0, i])
            print("maxerr", err.max(), "n bad envs", len(set(bad[:, 0])))
            return False
    print("ok", n, steps, seed, off, random_opp, flags)
    return True

if __name__ == "__main__":
    run(192, 300, 11, 5000, False, 7.5)
    run(512, 600, 12, 0, False, 40.0)
    run(512, 600, 13, 0, True, 40.0)
