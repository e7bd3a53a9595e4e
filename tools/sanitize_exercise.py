"""Smoke-sized pass over EVERY kernel of libfutbol_b200.so, meant to run under compute-sanitizer
(tools/sanitize.sh): memcheck (out-of-bounds / misaligned), synccheck (barriers under divergence) and
racecheck (shared-memory hazards).  Batches are tiny and not multiples of the block size so that tail
threads, masked resets and -- for the time-sliced rollout -- units that must wait on their predecessor
are all exercised.  Prints the launch count per variant; exits non-zero on any API error.
"""
import sys

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from gym_futbol_b200 import FutbolV1VecEnv, FutbolVecEnv, rollout_buffer  # noqa: E402


def v0(n, K, random_opp, slices, dtype):
    env = FutbolVecEnv(n, seed=3, env_id_offset=17, random_opp=random_opp, game_time=1.5, dtype=dtype)
    env.set_rollout_slices(slices)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(1)
    for _ in range(3):
        a = torch.randint(0, 16, (n,), dtype=torch.uint8, device="cuda", generator=g)
        env.step(a)
        if random_opp:
            env.step(a, opp_actions=a)
    mask = (torch.arange(n, device="cuda") % 3 == 0).to(torch.uint8)
    env.reset(mask)
    env.rollout(K)                                                       # in-kernel actions
    acts = torch.randint(0, 16, (K, n), dtype=torch.uint8, device="cuda", generator=g)
    env.rollout(K, actions=acts)
    if random_opp:
        env.rollout(K, actions=acts, opp_actions=acts)
    env.rollout(K, actions=acts, obs=False, reward=False, done=False)   # statistics only
    st = env.get_state()
    env.set_state(st)
    env.rollout(5)
    torch.cuda.synchronize()
    return env.launch_count


def v1(n, K, N, dtype):
    env = FutbolV1VecEnv(n, number_of_player=N, seed=5, env_id_offset=9, total_time=1.0, dtype=dtype)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(2)
    for _ in range(3):
        a = torch.randint(0, 5, (n, 2 * N), dtype=torch.uint8, device="cuda", generator=g)
        env.step(a)
        env.step(a, opp_actions=a)
    env.reset((torch.arange(n, device="cuda") % 2).to(torch.uint8))
    env.rollout(K)
    acts = torch.randint(0, 5, (K, n, 2 * N), dtype=torch.uint8, device="cuda", generator=g)
    env.rollout(K, actions=acts)
    env.rollout(K, actions=acts, opp_actions=acts)
    st = env.get_state()
    if hasattr(env, "_set_state_supported") or "jn" in (st.dtype.names or ()):
        env.set_state(st)
        env.rollout(3)
    torch.cuda.synchronize()
    return env.launch_count


def main():
    total = 0
    for random_opp in (False, True):
        for n, K, slices in ((77, 24, 1), (300, 24, 4), (31, 12, 12)):
            total += v0(n, K, random_opp, slices, torch.float32)
    total += v0(45, 8, False, 1, torch.float64)
    for N in (1, 2, 5, 10):
        total += v1(70, 16, N, torch.float32)
    total += v1(33, 8, 3, torch.float64)
    T, n = 16, 100
    rew = torch.randn(T, n, device="cuda")
    done = (torch.rand(T, n, device="cuda") < 0.1).to(torch.uint8)
    val = torch.randn(T + 1, n, device="cuda")
    rollout_buffer.gae(rew, done, val, 0.99, 0.95)
    if hasattr(rollout_buffer, "gather_minibatch"):
        obs = torch.randn(T, n, 30, device="cuda")
        idx = torch.randperm(T * n, device="cuda")[:257]
        rollout_buffer.gather_minibatch(obs, idx)
    torch.cuda.synchronize()
    print("sanitize_exercise ok: %d env kernels launched" % total)


if __name__ == "__main__":
    main()
