"""GPU tests of the glue around the path: device GAE against a numpy loop, CUDA-graph capture of the fused
rollout, and that the per-step outputs handed to a policy are the env's own buffers (zero copy)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available()
    return torch


def gae_numpy(r, d, v, gamma, lam):
    T, n = r.shape
    adv = np.zeros((T, n), np.float32)
    a = np.zeros(n, np.float32)
    g, l = np.float32(gamma), np.float32(lam)
    for t in range(T - 1, -1, -1):
        nd = (1 - d[t]).astype(np.float32)
        delta = (r[t] + (g * v[t + 1]) * nd) - v[t]
        a = delta + ((g * l) * nd) * a
        adv[t] = a
    return adv, adv + v[:-1]


def test_gae_matches_numpy_bit_exact(torch_cuda):
    from gym_futbol_b200.rollout_buffer import gae
    rng = np.random.default_rng(0)
    T, n = 128, 1000
    r = rng.normal(0, 5, (T, n)).astype(np.float32)
    d = (rng.random((T, n)) < 0.02).astype(np.uint8)
    v = rng.normal(0, 10, (T + 1, n)).astype(np.float32)
    adv, ret = gae(torch_cuda.from_numpy(r).cuda(), torch_cuda.from_numpy(d).cuda(), torch_cuda.from_numpy(v).cuda(), 0.99, 0.95)
    want_adv, want_ret = gae_numpy(r, d, v, 0.99, 0.95)
    assert np.array_equal(adv.cpu().numpy(), want_adv) and np.array_equal(ret.cpu().numpy(), want_ret)
    with pytest.raises(ValueError):
        gae(torch_cuda.zeros((4, 3)).cuda(), torch_cuda.zeros((4, 3), dtype=torch_cuda.uint8).cuda(), torch_cuda.zeros((4, 3)).cuda())


def test_rollout_is_cuda_graph_capturable(torch_cuda):
    """The C ABI only enqueues on the caller's stream, so a rollout can be captured once and replayed."""
    torch = torch_cuda
    from gym_futbol_b200 import FutbolVecEnv
    n, K = 4096, 16
    a = FutbolVecEnv(n, seed=5, random_opp=False)
    b = FutbolVecEnv(n, seed=5, random_opp=False)
    a.reset(); b.reset()
    acts = torch.randint(0, 16, (K, n), dtype=torch.uint8, device="cuda")
    a.rollout(K, actions=acts)                     # allocate the cached buffers outside the capture
    b.rollout(K, actions=acts)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            oa, ra, da = a.rollout(K, actions=acts)
    torch.cuda.current_stream().wait_stream(s)
    for _ in range(3):
        g.replay()
        ob, rb, db = b.rollout(K, actions=acts)
        torch.cuda.synchronize()
        assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(da, db)


def test_step_outputs_are_the_env_buffers(torch_cuda):
    torch = torch_cuda
    from gym_futbol_b200 import FutbolVecEnv
    env = FutbolVecEnv(256, seed=1)
    obs = env.reset()
    p0 = obs.data_ptr()
    for _ in range(3):
        obs, rew, done, info = env.step(torch.zeros(256, dtype=torch.uint8, device="cuda"))
        assert obs.data_ptr() == p0 == env.obs.data_ptr() and rew.data_ptr() == env.rewards.data_ptr()
        assert obs.is_cuda and rew.is_cuda and done.is_cuda and info["terminal_observation"].is_cuda


def test_replay_of_a_rollout_buffer(torch_cuda, tmp_path):
    """An env's trajectory cut out of the CUDA rollout buffer equals the per-step states of that env."""
    from gym_futbol_b200 import FutbolVecEnv
    from gym_futbol_b200.replay import save_gif, trajectory
    n, K, i = 64, 40, 17
    env = FutbolVecEnv(n, seed=3, random_opp=False, dtype=torch_cuda.float64)
    env.reset()
    obs, _, _ = env.rollout(K)
    t = trajectory(obs, env=i, variant="v0")
    o = obs[:, i].cpu().numpy().astype(np.float64)
    assert np.array_equal(t["ball"], o[:, 20:22]) and np.array_equal(t["team_b"][:, 1], o[:, 15:17])
    assert ((t["owner"] >= 0) & (t["owner"] <= 4)).all()
    assert save_gif(str(tmp_path / "r.gif"), t, scale=3) == K
