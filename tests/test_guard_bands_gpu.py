"""A memcheck of our own (compute-sanitizer is refused on the GPU pool, profiles/r2_sanitize_summary.txt): every buffer
the kernels touch lives between guard bands filled with a canary; after ragged-size runs of every kernel the guards must
be intact and the results must equal those of a run with ordinary allocations."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GUARD = 4096
CANARY = 0xA5


class Arena:
    """Carves tensors out of one uint8 allocation, a GUARD-byte canary band before and after each."""

    def __init__(self, torch, nbytes):
        self.torch = torch
        self.buf = torch.full((nbytes,), CANARY, dtype=torch.uint8, device="cuda")
        self.off = 0
        self.spans = []

    def take(self, shape, dtype):
        torch = self.torch
        n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
        start = (self.off + GUARD + 255) // 256 * 256
        self.off = start + n
        assert self.off + GUARD <= self.buf.numel(), "arena too small"
        self.spans.append((start, n))
        t = self.buf[start:start + n].view(dtype).view(shape)
        t.zero_()
        return t

    def intact(self):
        mask = self.torch.ones(self.buf.numel(), dtype=self.torch.bool, device="cuda")
        for start, n in self.spans:
            mask[start:start + n] = False
        return bool((self.buf[mask] == CANARY).all())


def _guarded(env, arena):
    """Re-homes every buffer of a vectorised env into the arena (before the first reset)."""
    torch = arena.torch
    n, D = env.num_envs, env.obs_dim
    env.state = arena.take((env.state.numel(),), torch.uint8)
    env.obs = arena.take((n, D), env.dtype)
    env.rewards = arena.take((n,), env.dtype)
    env.dones = arena.take((n,), torch.uint8)
    env.final_obs = arena.take((n, D), env.dtype)
    env.stats = arena.take((env.stats.numel(),), torch.uint8)
    env._step_args = (env.state.data_ptr(), env.obs.data_ptr(), env.rewards.data_ptr(), env.dones.data_ptr(), env.final_obs.data_ptr())
    env._step_info = {"terminal_observation": env.final_obs}
    return env


def _exercise(torch, env, arena, K, act_hi, vs):
    """Every entry point once or more; returns a digest of everything produced."""
    n = env.num_envs
    take = (lambda shape, dt: arena.take(shape, dt)) if arena is not None else (lambda shape, dt: torch.zeros(shape, dtype=dt, device="cuda"))
    g = torch.Generator(device="cuda").manual_seed(1)
    out = []
    env.reset()
    step_shape = (n,) + tuple(env.act_shape)
    for _ in range(3):
        a = take(step_shape, torch.uint8)
        a.copy_(torch.randint(0, act_hi, step_shape, dtype=torch.uint8, device="cuda", generator=g))
        o, r, d, _ = env.step(a)
        out += [o.clone(), r.clone(), d.clone()]
        if vs:
            o, r, d, _ = env.step(a, opp_actions=a)
            out += [o.clone(), r.clone(), d.clone()]
    mask = take((n,), torch.uint8)
    mask.copy_((torch.arange(n, device="cuda") % 3 == 0).to(torch.uint8))
    out.append(env.reset(mask).clone())
    bufs = (take((K, n, env.obs_dim), torch.float32), take((K, n), torch.float32), take((K, n), torch.uint8))
    acts = take((K,) + step_shape, torch.uint8)
    acts.copy_(torch.randint(0, act_hi, (K,) + step_shape, dtype=torch.uint8, device="cuda", generator=g))
    env.rollout(K, out=bufs)
    out += [b.clone() for b in bufs]
    env.rollout(K, actions=acts, out=bufs)
    out += [b.clone() for b in bufs]
    if vs:
        env.rollout(K, actions=acts, opp_actions=acts, out=bufs)
        out += [b.clone() for b in bufs]
    env.rollout(K, actions=acts, out=(None, None, None))
    st = env.get_state()
    env.set_state(st)
    env.rollout(3, out=(bufs[0][:3], bufs[1][:3], bufs[2][:3]))
    out += [bufs[0][:3].clone(), env.state.clone(), env.stats[8:].clone()]     # statistics without reward_sum (atomic adds: order-dependent rounding)
    torch.cuda.synchronize()
    return out


def _same(torch, a, b):
    return len(a) == len(b) and all(torch.equal(x.view(torch.uint8), y.view(torch.uint8)) for x, y in zip(a, b))


@pytest.mark.parametrize("random_opp,n,K,slices", [(False, 77, 24, 1), (True, 300, 24, 4), (False, 31, 12, 12), (True, 2049, 16, 3)])
def test_v0_kernels_stay_inside_their_buffers(random_opp, n, K, slices):
    import torch
    from gym_futbol_b200 import FutbolVecEnv
    runs = []
    for guarded in (True, False):
        for dtype in (torch.float32, torch.float64):
            env = FutbolVecEnv(n, seed=3, env_id_offset=17, random_opp=random_opp, game_time=1.5, dtype=dtype)
            env.set_rollout_slices(slices)
            arena = Arena(torch, 64 << 20) if guarded else None
            if guarded:
                _guarded(env, arena)
            runs.append(_exercise(torch, env, arena, K, 16, random_opp))
            if guarded:
                assert arena.intact(), "a v0 kernel wrote outside its buffers"
    assert _same(torch, runs[0], runs[2]) and _same(torch, runs[1], runs[3])


@pytest.mark.parametrize("N,n,K,slices", [(1, 70, 16, 1), (2, 70, 16, 4), (3, 33, 8, 1), (5, 70, 16, 3), (10, 45, 12, 12), (2, 2049, 16, 3)])
def test_v1_kernels_stay_inside_their_buffers(N, n, K, slices):
    import torch
    from gym_futbol_b200 import FutbolV1VecEnv
    runs = []
    for guarded in (True, False):
        for dtype in (torch.float32, torch.float64):
            env = FutbolV1VecEnv(n, number_of_player=N, seed=5, env_id_offset=9, total_time=1.0, dtype=dtype)
            env.set_rollout_slices(slices)                      # > 1: the work-queue kernel and its scheduler words in the state buffer
            arena = Arena(torch, 64 << 20) if guarded else None
            if guarded:
                _guarded(env, arena)
            runs.append(_exercise(torch, env, arena, K, 5, True))
            if guarded:
                assert arena.intact(), "a v1 kernel wrote outside its buffers"
    assert _same(torch, runs[0], runs[2]) and _same(torch, runs[1], runs[3])


def test_gae_and_gather_stay_inside_their_buffers():
    import torch
    from gym_futbol_b200 import rollout_buffer
    arena = Arena(torch, 16 << 20)
    T, n = 16, 101
    rew, done, val = arena.take((T, n), torch.float32), arena.take((T, n), torch.uint8), arena.take((T + 1, n), torch.float32)
    rew.copy_(torch.randn(T, n, device="cuda")); val.copy_(torch.randn(T + 1, n, device="cuda"))
    done.copy_((torch.rand(T, n, device="cuda") < 0.1).to(torch.uint8))
    adv, ret = arena.take((T, n), torch.float32), arena.take((T, n), torch.float32)
    rollout_buffer.gae(rew, done, val, 0.99, 0.95, out=(adv, ret))
    if hasattr(rollout_buffer, "gather_minibatch"):
        obs = arena.take((T, n, 30), torch.float32)
        obs.copy_(torch.randn(T, n, 30, device="cuda"))
        idx = arena.take((257,), torch.int64)
        idx.copy_(torch.randperm(T * n, device="cuda")[:257])
        dst = arena.take((257, 30), torch.float32)
        rollout_buffer.gather_minibatch(obs, idx, out=dst)
        assert torch.equal(dst, obs.view(T * n, 30)[idx])
    torch.cuda.synchronize()
    assert arena.intact()
