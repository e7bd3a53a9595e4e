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


@pytest.mark.parametrize("variant", ["v0", "v1"])
def test_time_sliced_rollout_is_cuda_graph_capturable(torch_cuda, variant):
    """The work-queue launch (a memset of the scheduler words + one kernel) captures and replays like the plain one."""
    torch = torch_cuda
    from gym_futbol_b200 import FutbolVecEnv, FutbolV1VecEnv
    n, K = 3000, 16

    def make():
        e = FutbolVecEnv(n, seed=5, random_opp=True) if variant == "v0" else FutbolV1VecEnv(n, number_of_player=3, seed=5)
        e.set_rollout_slices(4)
        e.reset()
        return e
    a, b = make(), make()
    assert "sliced" in a.rollout_kernel(K)
    a.rollout(K)                                   # allocate the cached buffers, query the occupancy: outside the capture
    b.rollout(K)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            oa, ra, da = a.rollout(K)
    torch.cuda.current_stream().wait_stream(s)
    for _ in range(3):
        g.replay()
        ob, rb, db = b.rollout(K)
        torch.cuda.synchronize()
        assert torch.equal(oa, ob) and torch.equal(ra, rb) and torch.equal(da, db)
    assert a.get_state().tobytes() == b.get_state().tobytes()


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


def test_step_writes_into_caller_buffers(torch_cuda):
    """step(out=...) makes the kernel write observation / reward / done straight into rows of a rollout buffer: same
    values as the env's own buffers would hold, which are left untouched; bad destinations are refused."""
    from gym_futbol_b200 import FutbolV1VecEnv, FutbolVecEnv
    torch = torch_cuda
    for make, hi, D in ((lambda: FutbolVecEnv(300, seed=4, random_opp=False), 16, 30),
                        (lambda: FutbolV1VecEnv(300, number_of_player=2, seed=4), 5, 20)):
        a_env, b_env = make(), make()
        a_env.reset(); b_env.reset()
        n, T = 300, 12
        obs_buf = torch.full((T + 1, n, D), -7.0, device="cuda")
        rew_buf, done_buf = torch.zeros((T, n), device="cuda"), torch.zeros((T, n), dtype=torch.uint8, device="cuda")
        own = b_env.obs.clone()
        g = torch.Generator(device="cuda").manual_seed(0)
        for t in range(T):
            act = torch.randint(0, hi, (n,) + tuple(a_env.act_shape), dtype=torch.uint8, device="cuda", generator=g)
            o, r, d, _ = a_env.step(act)
            o2, r2, d2, _ = b_env.step(act, out=(obs_buf[t + 1], rew_buf[t], done_buf[t]))
            assert o2.data_ptr() == obs_buf[t + 1].data_ptr()
            assert torch.equal(o, obs_buf[t + 1]) and torch.equal(r, rew_buf[t]) and torch.equal(d, done_buf[t])
        assert torch.equal(b_env.obs, own)                      # the env's own observation buffer was not written
        assert (obs_buf[0] == -7.0).all()
        with pytest.raises(ValueError):
            b_env.step(act, out=(obs_buf[:, :, 0], rew_buf[0], done_buf[0]))
        with pytest.raises(ValueError):
            b_env.step(act, out=(obs_buf[1], rew_buf[0].double(), done_buf[0]))
        o3, r3, d3, _ = b_env.step(act, out=(None, rew_buf[0], None))     # entries may be None
        assert o3 is None and d3 is None


def test_gather_minibatch_equals_torch_indexing(torch_cuda):
    from gym_futbol_b200 import rollout_buffer
    torch = torch_cuda
    g = torch.Generator(device="cuda").manual_seed(3)
    for T, n, D, m in ((16, 257, 30, 1031), (8, 64, 44, 512), (4, 33, 12, 132), (128, 4096, 30, 131072)):
        obs = torch.randn((T, n, D), device="cuda", generator=g)
        act = torch.randint(0, 16, (T, n), dtype=torch.uint8, device="cuda", generator=g)
        cols = [torch.randn((T, n), device="cuda", generator=g) for _ in range(4)]
        idx = torch.randperm(T * n, device="cuda", generator=g)[:m]
        got = rollout_buffer.gather_minibatch(obs, idx, act=act, cols=cols)
        want = [obs.view(T * n, D)[idx], act.view(-1)[idx]] + [c.view(-1)[idx] for c in cols]
        assert len(got) == 6 and all(torch.equal(a, b) for a, b in zip(got, want))
        only = rollout_buffer.gather_minibatch(obs, idx)
        assert torch.equal(only, want[0])
        dst = (torch.empty((m, D), device="cuda"), torch.empty(m, dtype=torch.uint8, device="cuda"), torch.empty(m, device="cuda"))
        out = rollout_buffer.gather_minibatch(obs, idx, act=act, cols=cols[:1], out=dst)
        assert out[0].data_ptr() == dst[0].data_ptr() and torch.equal(dst[0], want[0]) and torch.equal(dst[2], want[2])
    with pytest.raises(ValueError):
        rollout_buffer.gather_minibatch(obs, idx.int())
    with pytest.raises(ValueError):
        rollout_buffer.gather_minibatch(obs, idx, cols=cols + cols)


def test_ppo_example_fused_equals_unfused(torch_cuda):
    """configs[3] in small: the fused flow (step(out=...) + one-launch gather) computes what the copy / indexing flow
    computes -- same rollout buffers after collection, and a finite loss after an update."""
    import importlib.util
    import os
    torch = torch_cuda
    spec = importlib.util.spec_from_file_location("ppo_v0", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "ppo_v0.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    runs = []
    for fused in (True, False):
        ppo = mod.PPO(n_envs=512, n_steps=16, minibatches=2, epochs=1, seed=1, bf16=False, fused=fused, graph=False, sampler="torch")
        torch.manual_seed(5)
        ppo.collect()
        ppo.collect()                                          # the second collection continues from the first
        torch.cuda.synchronize()
        runs.append([t.clone() for t in (ppo.obs_buf[:16], ppo.act_buf, ppo.rew_buf, ppo.done_buf, ppo.adv, ppo.ret)])
        ppo.update()
        assert torch.isfinite(ppo.last_loss).item()
    assert all(torch.equal(a, b) for a, b in zip(*runs))


def test_host_rollout_delivers_the_device_rollout(torch_cuda):
    """gym_futbol_b200.host_io.HostRollout (bench.py's e2e path): pinned-host actions in, every observation / reward / done
    of the K steps in pinned host memory -- the same bytes a device-resident rollout with those actions produces."""
    from gym_futbol_b200 import FutbolV1VecEnv, FutbolVecEnv, host_io
    torch = torch_cuda
    for make, hi in ((lambda: FutbolVecEnv(1000, seed=8, random_opp=False), 16), (lambda: FutbolV1VecEnv(333, number_of_player=2, seed=8), 5)):
        a_env, b_env = make(), make()
        a_env.reset(); b_env.reset()
        K = 32
        pipe = host_io.HostRollout(b_env, K, chunks=4)
        assert pipe.h_obs.is_pinned() and pipe.h_actions.is_pinned()
        pipe.h_actions.copy_(torch.randint(0, hi, tuple(pipe.h_actions.shape), dtype=torch.uint8))
        for rep in range(2):
            want = a_env.rollout(K, actions=pipe.h_actions.cuda())
            h_obs, h_rew, h_done = pipe.run()
            assert torch.equal(h_obs, want[0].cpu()) and torch.equal(h_rew, want[1].cpu()) and torch.equal(h_done, want[2].cpu())
        assert pipe.d2h_bytes == K * b_env.num_envs * (b_env.obs_dim * 4 + 5) + 64 and pipe.h2d_bytes == pipe.h_actions.numel()
        pipe.run_resident()
        a_env.rollout(K, actions=pipe.h_actions.cuda())
        assert a_env.get_state().tobytes() == b_env.get_state().tobytes()
    with pytest.raises(ValueError):
        host_io.HostRollout(b_env, 30, chunks=4)
    assert host_io.measure_d2h_peak("cuda:0", nbytes=1 << 24, reps=1) > 0.1
    cpus = host_io.gpu_local_cpus(0)
    assert cpus is None or len(cpus) >= 1


def _sampler_uniforms(seed, t, n):
    from oracle import philox
    k0, k1 = philox.seed_key(seed)
    w = philox.philox4x32_10_np(np.full(n, t & 0xFFFFFFFF, np.uint64), np.full(n, t >> 32, np.uint64), np.arange(n, dtype=np.uint64),
                                np.full(n, 4, np.uint64), k0, k1)[0]
    return (w >> np.uint64(8)).astype(np.float64) / 16777216.0


def test_sample_actions_follows_its_rule(torch_cuda):
    """futbol_sample_actions against a float64 restatement of its rule with the same Philox uniforms (oracle/philox.py), and
    its log-probabilities against torch's log_softmax; bf16 logits; the device counter."""
    torch = torch_cuda
    from gym_futbol_b200.rollout_buffer import sample_actions
    n, A, seed, t = 20000, 16, 77, (5 << 32) + 123
    g = torch.Generator(device="cuda").manual_seed(3)
    logits = torch.randn(n, A, device="cuda", generator=g) * 3.0
    logits[:50, 3] = -200.0                                     # terms that underflow to zero are never picked
    logits[50:60] = 0.0
    act, logp = sample_actions(logits, seed=seed, t=t)
    torch.cuda.synchronize()
    l = logits.double().cpu().numpy()
    e = np.exp(l - l.max(1, keepdims=True))
    cum, tot = np.cumsum(e, 1), e.sum(1)
    target = _sampler_uniforms(seed, t, n) * tot
    a = act.cpu().numpy().astype(np.int64)
    rows = np.arange(n)
    lo = np.where(a > 0, cum[rows, np.maximum(a - 1, 0)], 0.0)
    eps = 1e-5 * tot                                            # the kernel sums in float32
    assert np.all(lo <= target + eps) and np.all(cum[rows, a] > target - eps)
    exact = np.argmax(cum > target[:, None], 1)
    assert np.mean(exact == a) > 0.999                          # away from the float32 rounding band the pick IS the rule's
    assert not np.any(a[:50] == 3)
    want = torch.log_softmax(logits, -1).gather(1, act.long().unsqueeze(1)).squeeze(1)
    assert torch.allclose(logp, want, atol=2e-5, rtol=0)
    # bf16 logits: the same picks as float32 logits holding the same values
    lb = logits.to(torch.bfloat16)
    a_b, lp_b = sample_actions(lb, seed=seed, t=t)
    a_f, lp_f = sample_actions(lb.float(), seed=seed, t=t)
    assert torch.equal(a_b, a_f) and torch.equal(lp_b, lp_f)
    # the device counter: t_base + t_off is the draw's t; out= buffers
    base = torch.tensor([t - 7], dtype=torch.int64, device="cuda")
    out = (torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, device="cuda"))
    a2, lp2 = sample_actions(logits, seed=seed, t=7, t_base=base, out=out)
    assert a2.data_ptr() == out[0].data_ptr() and torch.equal(a2, act) and torch.equal(lp2, logp)
    assert not torch.equal(sample_actions(logits, seed=seed, t=t + 1)[0], act)
    with pytest.raises(ValueError):
        sample_actions(torch.zeros(4, 33, device="cuda"))


def test_sample_actions_distribution(torch_cuda):
    """400,000 draws from one softmax: every frequency within five standard deviations."""
    torch = torch_cuda
    from gym_futbol_b200.rollout_buffer import sample_actions
    n = 400000
    row = torch.tensor([0.3, -1.0, 2.0, 0.0, -4.0, 1.5, 0.7, -0.2, 1.1, -2.5, 0.0, 0.9, -0.6, 2.2, -1.7, 0.4], device="cuda")
    act, _ = sample_actions(row.repeat(n, 1), seed=5, t=9)
    p = torch.softmax(row.double(), 0).cpu().numpy()
    freq = np.bincount(act.cpu().numpy(), minlength=16) / n
    assert np.all(np.abs(freq - p) < 5 * np.sqrt(p * (1 - p) / n))


def test_ppo_example_kernel_sampler_and_graph(torch_cuda):
    """The fused flow with futbol_sample_actions, eager and as a CUDA graph: the graph replay draws fresh numbers (device
    counter), the stored log-probabilities are those of the stored actions, and the update gives a finite loss."""
    import importlib.util
    import os
    torch = torch_cuda
    spec = importlib.util.spec_from_file_location("ppo_v0", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples", "ppo_v0.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ppo = mod.PPO(n_envs=1024, n_steps=8, minibatches=2, epochs=1, seed=2, bf16=False, fused=True, graph=True)
    assert ppo.sampler == "kernel"
    ppo.collect()
    a0 = ppo.act_buf.clone()
    with torch.no_grad():
        logits, _ = ppo.policy(ppo.obs_buf[3])
    want = torch.log_softmax(logits.float(), -1).gather(1, ppo.act_buf[3].long().unsqueeze(1)).squeeze(1)
    assert torch.allclose(ppo.logp_buf[3], want, atol=1e-4)
    ppo.prepare_graph()
    assert ppo.graph is not None
    ppo.collect()
    a1 = ppo.act_buf.clone()
    ppo.collect()
    a2 = ppo.act_buf.clone()
    torch.cuda.synchronize()
    assert int(ppo.t_base.item()) == 24 and not torch.equal(a1, a2) and not torch.equal(a0, a1)
    assert 0.02 < (a1 == a2).float().mean().item() < 0.2       # two replays agree about as often as two uniform draws of 16 would
    ppo.update()
    assert torch.isfinite(ppo.last_loss).item()
