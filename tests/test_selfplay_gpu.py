"""Opponent actions supplied by the caller (futbol_step_vs / futbol_rollout_vs): the CUDA
path against the oracle run with the same opponent actions."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    return torch


def test_v0_given_opponent_actions_match_oracle(torch_cuda):
    from gym_futbol_b200 import FutbolError, FutbolVecEnv
    from oracle.v0 import OracleV0
    n, steps, seed = 200, 450, 8
    rng = np.random.default_rng(3)
    acts = rng.integers(0, 16, (steps, n), dtype=np.uint8)
    opp = rng.integers(0, 16, (steps, n), dtype=np.uint8)
    orc = OracleV0(n, seed=seed, random_opp=True, arith=0)
    want = orc.rollout(steps, actions=acts, autoreset=2, n_threads=4, opp_actions=opp)
    env = FutbolVecEnv(n, seed=seed, random_opp=True, dtype=torch_cuda.float64)
    env.reset()
    for t in range(steps):
        obs, rew, done, _ = env.step(torch_cuda.from_numpy(acts[t]).cuda(), opp_actions=torch_cuda.from_numpy(opp[t]).cuda())
        assert np.array_equal(obs.cpu().numpy(), want["obs"][t]) and np.array_equal(rew.cpu().numpy(), want["reward"][t]), t
        assert np.array_equal(done.cpu().numpy(), want["done"][t])
    # fused, open-loop replay of the same opponent actions
    env2 = FutbolVecEnv(n, seed=seed, random_opp=True)
    env2.reset()
    o, r, d = env2.rollout(steps, actions=torch_cuda.from_numpy(acts).cuda(), opp_actions=torch_cuda.from_numpy(opp).cuda())
    assert np.array_equal(o.cpu().numpy(), want["obs"].astype(np.float32)) and np.array_equal(d.cpu().numpy(), want["done"])
    # a step without opponent actions differs (the random opponents act), and the hard-coded variant refuses them
    env3 = FutbolVecEnv(n, seed=seed, random_opp=False)
    env3.reset()
    with pytest.raises(FutbolError):
        env3.step(torch_cuda.from_numpy(acts[0]).cuda(), opp_actions=torch_cuda.from_numpy(opp[0]).cuda())


def test_v1_given_right_team_actions_match_oracle(torch_cuda):
    from gym_futbol_b200 import FutbolV1VecEnv
    from oracle.v1 import OracleV1
    N, n, steps, seed = 3, 150, 330, 6
    rng = np.random.default_rng(4)
    left = rng.integers(0, 5, (steps, n, 2 * N), dtype=np.uint8)
    right = rng.integers(0, 5, (steps, n, 2 * N), dtype=np.uint8)
    orc = OracleV1(n, seed=seed, number_of_player=N)
    want = orc.rollout(steps, actions=left, autoreset=2, n_threads=4, right_actions=right)
    env = FutbolV1VecEnv(n, number_of_player=N, seed=seed)
    env.reset()
    o, r, d = env.rollout(steps, actions=torch_cuda.from_numpy(left).cuda(), opp_actions=torch_cuda.from_numpy(right).cuda())
    assert np.array_equal(o.cpu().numpy(), want["obs"].astype(np.float32))
    assert np.array_equal(r.cpu().numpy(), want["reward"].astype(np.float32)) and np.array_equal(d.cpu().numpy(), want["done"])
    env2 = FutbolV1VecEnv(n, number_of_player=N, seed=seed, dtype=torch_cuda.float64)
    env2.reset()
    for t in range(40):
        obs, rew, done, _ = env2.step(torch_cuda.from_numpy(left[t]).cuda(), opp_actions=torch_cuda.from_numpy(right[t]).cuda())
        assert np.array_equal(obs.cpu().numpy(), want["obs"][t]) and np.array_equal(rew.cpu().numpy(), want["reward"][t])
