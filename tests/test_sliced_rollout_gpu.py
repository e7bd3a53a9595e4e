"""Time-sliced rollout (futbol_set_rollout_slices, csrc/v0_kernels.cu): the work-queue launch must give the very
bytes of the plain launch -- observations, rewards, dones, statistics and the final state -- whatever the number
of slices, and agree with the oracle.  Small batches make every unit wait on its predecessor (fewer env-blocks than
resident blocks), the 131,072-env case is the shape the slicing exists for (one of eight ranks of the 2^20 job)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(n, K, reps, slices, random_opp, seed=11, off=4242, acts=None, outputs=True):
    import torch
    from gym_futbol_b200 import FutbolVecEnv
    env = FutbolVecEnv(n, seed=seed, env_id_offset=off, random_opp=random_opp, game_time=3.0)
    env.set_rollout_slices(slices)
    env.reset()
    outs = []
    for rep in range(reps):
        a = None if acts is None else acts[rep]
        if outputs:
            o, r, d = env.rollout(K, actions=a)
            outs.append((o.cpu().numpy().copy(), r.cpu().numpy().copy(), d.cpu().numpy().copy()))
        else:
            env.rollout(K, actions=a, obs=False, reward=False, done=False)
    torch.cuda.synchronize()
    return outs, env.get_state(), env.read_stats(), env.launch_count


@pytest.mark.parametrize("random_opp", [False, True])
@pytest.mark.parametrize("n,K,slices", [(4096 + 77, 64, 8), (4096 + 77, 64, 5), (513, 37, 37), (31, 20, 3)])
def test_sliced_equals_plain(n, K, slices, random_opp):
    plain = _run(n, K, 3, 1, random_opp)
    sliced = _run(n, K, 3, slices, random_opp)
    for (o0, r0, d0), (o1, r1, d1) in zip(plain[0], sliced[0]):
        assert np.array_equal(o0.view(np.uint32), o1.view(np.uint32))
        assert np.array_equal(r0.view(np.uint32), r1.view(np.uint32)) and np.array_equal(d0, d1)
    assert plain[1].tobytes() == sliced[1].tobytes()
    for key in ("env_steps", "episodes", "goals_ai", "goals_opp", "out_of_field"):
        assert plain[2][key] == sliced[2][key]
    assert abs(plain[2]["reward_sum"] - sliced[2]["reward_sum"]) <= 1e-9 * max(1.0, abs(plain[2]["reward_sum"]))
    assert plain[3] == sliced[3]                                # one kernel per rollout either way


def test_sliced_matches_oracle():
    from oracle import philox
    from oracle.v0 import OracleV0
    n, K, reps, seed, off = 300, 48, 3, 5, 77
    acts = [philox.actions_table(seed + 1, np.arange(off, off + n), rep * K, K) for rep in range(reps)]
    outs, st, stats, _ = _run(n, K, reps, 6, False, seed=seed, off=off, acts=acts)
    orc = OracleV0(n, seed=seed, env_id0=off, random_opp=False, game_time=3.0, arith=0)
    for rep in range(reps):
        want = orc.rollout(K, actions=acts[rep], autoreset=2, n_threads=4)
        o, r, d = outs[rep]
        assert np.array_equal(d, want["done"]) and np.array_equal(r, want["reward"].astype(np.float32))
        assert np.array_equal(o, want["obs"].astype(np.float32))
    assert np.array_equal(st["rows"].reshape(n, 25), orc.envs["obs"][:, :5].reshape(n, 25))
    assert stats["env_steps"] == n * K * reps


def test_rank_sized_batch_auto_slices():
    """131,072 envs x K = 64 with the default setting (sliced on a B200: 1024 env-blocks on 740 slots) against the
    plain launch: final states and statistics identical."""
    n, K = 131072, 64
    auto = _run(n, K, 2, 0, False, outputs=False)
    plain = _run(n, K, 2, 1, False, outputs=False)
    assert auto[1].tobytes() == plain[1].tobytes()
    for key in ("env_steps", "episodes", "goals_ai", "goals_opp", "out_of_field"):
        assert auto[2][key] == plain[2][key]


def test_slice_setting_is_validated():
    from gym_futbol_b200 import FutbolError, FutbolVecEnv
    env = FutbolVecEnv(64, seed=1)
    with pytest.raises(FutbolError):
        env.set_rollout_slices(-1)
    env.set_rollout_slices(1000)                        # more slices than steps: clamped to one step per slice
    env.reset()
    o, r, d = env.rollout(8)
    assert tuple(o.shape) == (8, 64, 30)


def test_sliced_rollout_is_cuda_graph_capturable():
    """The sliced launch = a memset of the work queue + one kernel, both on the caller's stream: captured once, replayed
    (the queue is re-zeroed by the captured memset), against the same rollouts launched eagerly."""
    import torch
    from gym_futbol_b200 import FutbolVecEnv
    n, K = 2048 + 5, 24
    a = FutbolVecEnv(n, seed=9, random_opp=False)
    b = FutbolVecEnv(n, seed=9, random_opp=False)
    for e in (a, b):
        e.set_rollout_slices(6)
        e.reset()
    acts = torch.randint(0, 16, (K, n), dtype=torch.uint8, device="cuda")
    a.rollout(K, actions=acts)                         # cached buffers allocated outside the capture
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



@pytest.mark.parametrize("random_opp", [False, True])
@pytest.mark.parametrize("n,K", [(4096 + 77, 64), (31, 20), (131072, 16)])
def test_dense_kernel_equals_standard(n, K, random_opp):
    """The dense rollout kernel (28 warps per SM, observation staged in three passes) gives the bytes of the standard one."""
    import torch
    from gym_futbol_b200 import FutbolVecEnv
    runs = []
    for variant in (1, 2):
        env = FutbolVecEnv(n, seed=11, env_id_offset=4242, random_opp=random_opp, game_time=3.0)
        env.set_rollout_variant(variant)
        env.set_rollout_slices(1)
        assert env.rollout_kernel(K) == ("v0_rollout_kernel", "v0_rollout_dense_kernel")[variant - 1]
        env.reset()
        outs = []
        for rep in range(2):
            o, r, d = env.rollout(K)
            outs += [o.clone(), r.clone(), d.clone()]
        torch.cuda.synchronize()
        runs.append((outs, env.get_state().tobytes(), env.read_stats()))
    for a, b in zip(runs[0][0], runs[1][0]):
        assert torch.equal(a.view(torch.uint8), b.view(torch.uint8))
    assert runs[0][1] == runs[1][1]
    for key in ("env_steps", "episodes", "goals_ai", "goals_opp", "out_of_field"):
        assert runs[0][2][key] == runs[1][2][key]


def test_automatic_kernel_choice():
    """Plain up to one wave of blocks and beyond ten, time-sliced in between (4 slices at one rank of eight of the 2^20 job, 3
    from two to six waves, 2 up to ten: profiles/r2_slices.md); the dense kernel only on request."""
    from gym_futbol_b200 import FutbolVecEnv
    assert FutbolVecEnv(4096, seed=0).rollout_kernel(64) == "v0_rollout_kernel"          # under one wave
    env = FutbolVecEnv(131072, seed=0)
    assert env.rollout_kernel(64) == "v0_rollout_sliced_kernel" and env.rollout_slices(64) == 4
    env.set_rollout_slices(1)
    assert env.rollout_kernel(64) == "v0_rollout_kernel"
    env.set_rollout_variant(2)
    assert env.rollout_kernel(64) == "v0_rollout_dense_kernel"
    env = FutbolVecEnv(262144, seed=0)                                                   # 2.8 waves: three slices
    assert env.rollout_kernel(64) == "v0_rollout_sliced_kernel" and env.rollout_slices(64) == 3
    assert FutbolVecEnv(1 << 20, seed=0).rollout_kernel(64) == "v0_rollout_kernel"       # 11 waves: plain


# ---- v1: the same queue with warps of 32 envs as units (csrc/v1_kernels.cu, v1_rollout_sliced_kernel) ----
def _run_v1(N, n, K, reps, slices, seed=13, off=900, given=True):
    import torch
    from gym_futbol_b200 import FutbolV1VecEnv
    env = FutbolV1VecEnv(n, number_of_player=N, seed=seed, env_id_offset=off, total_time=4.0)
    env.set_rollout_slices(slices)
    env.reset()
    g = torch.Generator(device="cuda").manual_seed(7)
    outs = []
    for rep in range(reps):
        a = torch.randint(0, 5, (K, n, 2 * N), dtype=torch.uint8, device="cuda", generator=g) if given else None
        o, r, d = env.rollout(K, actions=a)
        outs.append((o.cpu().numpy().copy(), r.cpu().numpy().copy(), d.cpu().numpy().copy()))
    torch.cuda.synchronize()
    return outs, env.get_state(), env.read_stats(), env.rollout_kernel(K)


@pytest.mark.parametrize("N,n,K,slices", [(2, 4096 + 77, 64, 4), (5, 2048 + 5, 48, 5), (1, 513, 37, 37), (10, 333, 20, 3),
                                          (3, 31, 16, 2), (7, 1500, 33, 6)])
def test_v1_sliced_equals_plain(N, n, K, slices):
    """Small batches make every unit wait on its predecessor; episodes end inside the rollouts (total_time 4 = 40 steps)."""
    plain = _run_v1(N, n, K, 3, 1)
    sliced = _run_v1(N, n, K, 3, slices)
    assert plain[3] == "v1_rollout_kernel" and sliced[3] == "v1_rollout_sliced_kernel"
    for (o0, r0, d0), (o1, r1, d1) in zip(plain[0], sliced[0]):
        assert np.array_equal(o0.view(np.uint32), o1.view(np.uint32))
        assert np.array_equal(r0.view(np.uint32), r1.view(np.uint32)) and np.array_equal(d0, d1)
    assert plain[1].tobytes() == sliced[1].tobytes()            # bodies, scalars AND the arbiter cache of every env
    for key in ("env_steps", "episodes", "goals_ai", "goals_opp", "out_of_field"):
        assert plain[2][key] == sliced[2][key]
    assert plain[2]["env_steps"] == 3 * n * K


def test_v1_sliced_synthetic_actions():
    """actions=None (Philox stream 1 inside the kernel) through the sliced launch."""
    plain = _run_v1(5, 1000, 40, 2, 1, given=False)
    sliced = _run_v1(5, 1000, 40, 2, 4, given=False)
    for (o0, r0, d0), (o1, r1, d1) in zip(plain[0], sliced[0]):
        assert np.array_equal(o0.view(np.uint32), o1.view(np.uint32)) and np.array_equal(r0.view(np.uint32), r1.view(np.uint32))
        assert np.array_equal(d0, d1)
    assert plain[1].tobytes() == sliced[1].tobytes()


def test_v1_automatic_slices():
    """Plain up to one wave of warps, time-sliced beyond (BASELINE configs[4], 5v5 at 2^18 envs: 2 slices), same results."""
    import torch
    from gym_futbol_b200 import FutbolV1VecEnv
    assert FutbolV1VecEnv(4096, number_of_player=5, seed=0).rollout_kernel(64) == "v1_rollout_kernel"
    env = FutbolV1VecEnv(1 << 18, number_of_player=5, seed=0)
    assert env.rollout_kernel(64) == "v1_rollout_sliced_kernel" and env.rollout_slices(64) == 2
    assert FutbolV1VecEnv(1 << 17, number_of_player=5, seed=0).rollout_slices(64) == 4
    assert FutbolV1VecEnv(1 << 20, number_of_player=2, seed=0).rollout_slices(64) == 2
    n, K = 80000, 32                                            # 2500 warps on ~1776 slots at 5v5: sliced automatically
    outs = []
    for slices in (0, 1):
        e = FutbolV1VecEnv(n, number_of_player=5, seed=3, total_time=2.0)
        e.set_rollout_slices(slices)
        e.reset()
        assert (e.rollout_kernel(K) == "v1_rollout_sliced_kernel") == (slices == 0)
        o, r, d = e.rollout(K)
        outs.append((o.clone(), r.clone(), d.clone(), e.get_state().tobytes()))
    torch.cuda.synchronize()
    assert torch.equal(outs[0][0].view(torch.int32), outs[1][0].view(torch.int32)) and torch.equal(outs[0][1], outs[1][1])
    assert torch.equal(outs[0][2], outs[1][2]) and outs[0][3] == outs[1][3]
