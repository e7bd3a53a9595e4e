"""GPU parity tests for the v1 path: the CUDA kernels (through the C ABI) against the v1 oracle on the same
seeds and actions (bar: bit-exact float64 and exact integers), and against traces of the reference's own Python
(tests/golden/v1_golden.npz).  The oracle is pinned to those traces bit for bit (tests/test_oracle_v1_golden.py); the
physics under them is the pymunk stand-in, not the real library (rows b3 / b6 stay unpinned, oracle/futbol_v1_oracle.c)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    return torch


@pytest.mark.parametrize("N", [1, 2, 5, 10])
def test_v1_reset_is_the_kickoff_formation(torch_cuda, N):
    from gym_futbol_b200 import FutbolV1VecEnv
    from oracle.v1 import OracleV1
    env = FutbolV1VecEnv(7, number_of_player=N, seed=3, env_id_offset=50, dtype=torch_cuda.float64)
    obs = env.reset().cpu().numpy()
    orc = OracleV1(7, seed=3, env_id0=50, number_of_player=N)
    want = np.stack([orc.obs(i) for i in range(7)])
    assert obs.shape == (7, 4 + 8 * N) and np.array_equal(obs, want)
    st = env.get_state()
    assert np.array_equal(st["owner_side"], orc.envs["owner_side"].astype(np.uint8))
    assert env.episode_steps == 300


@pytest.mark.parametrize("N,n,steps", [(1, 64, 320), (2, 192, 640), (5, 96, 320), (10, 48, 310)])
def test_v1_step_api_matches_oracle_every_step(torch_cuda, N, n, steps):
    """Per-step API, float64 outputs, given left actions, auto-reset across the 300-step time limit."""
    from gym_futbol_b200 import FutbolV1VecEnv
    from oracle.v1 import OracleV1
    seed, off = 9, 700
    env = FutbolV1VecEnv(n, number_of_player=N, seed=seed, env_id_offset=off, dtype=torch_cuda.float64)
    orc = OracleV1(n, seed=seed, env_id0=off, number_of_player=N)
    env.reset()
    acts = np.random.default_rng(1).integers(0, 5, (steps, n, 2 * N), dtype=np.uint8)
    want = orc.rollout(steps, actions=acts, autoreset=2, n_threads=4)
    for t in range(steps):
        obs, rew, done, info = env.step(torch_cuda.from_numpy(acts[t]).cuda())
        assert np.array_equal(done.cpu().numpy(), want["done"][t]), t
        assert np.array_equal(rew.cpu().numpy(), want["reward"][t]), t
        assert np.array_equal(obs.cpu().numpy(), want["obs"][t]), t
    st = env.get_state()
    B = 2 * N + 1
    assert np.array_equal(st["body"][:, :B, 0:2], orc.envs["p"][:, :B]) and np.array_equal(st["body"][:, :B, 2:4], orc.envs["v"][:, :B])
    assert np.array_equal(st["body"][:, :B, 4:6], orc.envs["vb"][:, :B])
    assert np.array_equal(st["t_total"], orc.envs["t_total"]) and np.array_equal(st["ep_step"], orc.envs["ep_step"])
    assert want["done"].sum() > 0 and (want["flags"] & 1).sum() > 0


@pytest.mark.parametrize("N,n", [(2, 77), (2, 256), (5, 128)])
def test_v1_rollout_matches_oracle(torch_cuda, N, n):
    """Fused rollout, in-kernel synthetic left actions (Philox stream 1), fp32 streams = the oracle's doubles rounded once."""
    from gym_futbol_b200 import FutbolV1VecEnv
    from oracle.v1 import OracleV1
    K, seed = 330, 4
    env = FutbolV1VecEnv(n, number_of_player=N, seed=seed)
    orc = OracleV1(n, seed=seed, number_of_player=N)
    env.reset()
    obs, rew, done = env.rollout(K)
    want = orc.rollout(K, actions=None, autoreset=2, n_threads=4)
    assert np.array_equal(done.cpu().numpy(), want["done"])
    assert np.array_equal(rew.cpu().numpy(), want["reward"].astype(np.float32))
    assert np.array_equal(obs.cpu().numpy(), want["obs"].astype(np.float32))
    stats = env.read_stats()
    fl = want["flags"]
    assert stats["env_steps"] == n * K and stats["episodes"] == int(want["done"].sum())
    assert stats["goals_ai"] == int(((fl & 8) > 0).sum()) and stats["goals_opp"] == int((((fl & 1) > 0) & ((fl & 8) == 0)).sum())
    assert stats["out_of_field"] == int(((fl & 2) > 0).sum()) and stats["contacts_dropped"] == 0 and stats["contacts"] > 0
    # a second rollout continues from the stored state (state + arbiter cache round-trip through HBM)
    obs2, rew2, done2 = env.rollout(64)
    want2 = orc.rollout(64, actions=None, autoreset=2, n_threads=4)
    assert np.array_equal(obs2.cpu().numpy(), want2["obs"].astype(np.float32)) and np.array_equal(rew2.cpu().numpy(), want2["reward"].astype(np.float32))


def test_v1_sharding_invariance_and_rollout_equals_steps(torch_cuda):
    from gym_futbol_b200 import FutbolV1VecEnv
    a = FutbolV1VecEnv(96, number_of_player=2, seed=2, env_id_offset=1000)
    b = FutbolV1VecEnv(32, number_of_player=2, seed=2, env_id_offset=1064)
    a.reset(); b.reset()
    oa, ra, da = a.rollout(200)
    ob, rb, db = b.rollout(200)
    assert torch_cuda.equal(oa[:, 64:], ob) and torch_cuda.equal(ra[:, 64:], rb) and torch_cuda.equal(da[:, 64:], db)


def test_v1_single_env_dropin(torch_cuda):
    from gym_futbol_b200.envs_v1 import Futbol
    from oracle.v1 import OracleV1
    env = Futbol(number_of_player=2, seed=6, env_id=9)
    orc = OracleV1(1, seed=6, env_id0=9, number_of_player=2)
    assert env.action_space.nvec.tolist() == [5, 5, 5, 5] and env.observation_space.shape == (20,)
    obs = env.reset()
    assert np.array_equal(obs, orc.obs(0)) and env.ball_owner_side in ("left", "right")
    rng = np.random.default_rng(0)
    for t in range(305):
        a = rng.integers(0, 5, 4)
        o, r, d, info = env.step(a)
        wo, wr, wd = orc.step_one(0, a)
        assert np.array_equal(o, wo) and r == wr and d == wd and info == {}
        if d:
            assert t == 299
            env.reset(); orc.reset()
    with pytest.raises(ValueError):
        env.step([0, 0, 0, 7])


def test_v1_masked_reset_matches_oracle(torch_cuda):
    """reset(mask) re-kicks-off only the selected envs (time 0, new side draw, arbiter cache kept) mid-episode."""
    from gym_futbol_b200 import FutbolV1VecEnv
    from oracle.v1 import OracleV1
    N, n, seed = 2, 64, 13
    env = FutbolV1VecEnv(n, number_of_player=N, seed=seed, dtype=torch_cuda.float64, auto_reset=False)
    orc = OracleV1(n, seed=seed, number_of_player=N)
    env.reset()
    acts = np.random.default_rng(2).integers(0, 5, (120, n, 2 * N), dtype=np.uint8)
    want = orc.rollout(60, actions=acts[:60], autoreset=0)
    for t in range(60):
        obs, _, _, _ = env.step(torch_cuda.from_numpy(acts[t]).cuda())
    assert np.array_equal(obs.cpu().numpy(), want["obs"][-1])
    mask = (np.arange(n) % 3 == 0).astype(np.uint8)
    obs = env.reset(mask=torch_cuda.from_numpy(mask).cuda()).cpu().numpy()
    orc.reset(idx=np.flatnonzero(mask))
    ref = np.stack([orc.obs(i) for i in range(n)])
    assert np.array_equal(obs[mask == 1], ref[mask == 1])
    want = orc.rollout(60, actions=acts[60:], autoreset=0)
    for t in range(60):
        obs, rew, done, _ = env.step(torch_cuda.from_numpy(acts[60 + t]).cuda())
        assert np.array_equal(obs.cpu().numpy(), want["obs"][t]) and np.array_equal(rew.cpu().numpy(), want["reward"][t])
    st = env.get_state()
    assert (st["ep_step"][mask == 1] == 60).all() and (st["ep_step"][mask == 0] == 120).all()


def test_v1_and_v0_step_api_float32_outputs_ragged_batch(torch_cuda):
    """fp32 per-step outputs take the warp-cooperative observation writer; 77 envs leave a partial last warp."""
    from gym_futbol_b200 import FutbolV1VecEnv, FutbolVecEnv
    from oracle import philox
    from oracle.v0 import OracleV0
    from oracle.v1 import OracleV1
    n, steps = 77, 60
    env = FutbolV1VecEnv(n, number_of_player=2, seed=3)
    orc = OracleV1(n, seed=3, number_of_player=2)
    env.reset()
    acts = np.random.default_rng(5).integers(0, 5, (steps, n, 4), dtype=np.uint8)
    want = orc.rollout(steps, actions=acts, autoreset=2)
    for t in range(steps):
        obs, rew, done, _ = env.step(torch_cuda.from_numpy(acts[t]).cuda())
        assert obs.dtype == torch_cuda.float32 and np.array_equal(obs.cpu().numpy(), want["obs"][t].astype(np.float32))
        assert np.array_equal(rew.cpu().numpy(), want["reward"][t].astype(np.float32))
    env0 = FutbolVecEnv(n, seed=3, random_opp=False)
    orc0 = OracleV0(n, seed=3, random_opp=False, arith=0)
    env0.reset()
    a0 = philox.actions_table(3, np.arange(n), 0, steps)
    want0 = orc0.rollout(steps, actions=a0, autoreset=2)
    for t in range(steps):
        obs, rew, done, _ = env0.step(torch_cuda.from_numpy(a0[t]).cuda())
        assert np.array_equal(obs.cpu().numpy(), want0["obs"][t].astype(np.float32))
        assert np.array_equal(rew.cpu().numpy(), want0["reward"][t].astype(np.float32)) and np.array_equal(done.cpu().numpy(), want0["done"][t])


@pytest.mark.parametrize("N", [1, 2, 5, 10])
def test_v1_set_state_from_oracle_and_continue(torch_cuda, N):
    """Checkpoint / restore incl. the arbiter cache: the oracle's mid-trajectory state (bodies, bias velocities, cached
    impulses and their ages) is written with set_state and both continue identically; get_state -> set_state on a second
    env gives the same continuation; out-of-domain records are refused."""
    from gym_futbol_b200 import FutbolV1VecEnv
    from oracle.v1 import OracleV1
    n, seed, off, B = 96, 17, 4000, 2 * N + 1
    P = B * (B - 1) // 2 + 12 * B
    orc = OracleV1(n, seed=seed, env_id0=off, number_of_player=N)
    acts = np.random.default_rng(3).integers(0, 5, (360, n, 2 * N), dtype=np.uint8)
    orc.rollout(250, actions=acts[:250], autoreset=2, n_threads=4, record=False)
    env = FutbolV1VecEnv(n, number_of_player=N, seed=seed, env_id_offset=off, dtype=torch_cuda.float64)
    rec = np.zeros(n, dtype=env.STATE_DTYPE)
    e = orc.envs
    rec["body"][:, :B, 0:2], rec["body"][:, :B, 2:4], rec["body"][:, :B, 4:6] = e["p"][:, :B], e["v"][:, :B], e["vb"][:, :B]
    rec["t_total"], rec["ep_step"], rec["owner_side"] = e["t_total"], e["ep_step"], e["owner_side"]
    stamp = 1000
    rec["stamp"] = stamp
    age = e["age"][:, :P].astype(np.int64)
    rec["last"] = np.where(age == 255, 0, np.maximum(stamp - 1 - age, 0))        # kernel: stamp - last = age + 1
    rec["jn"] = e["jn"][:, :P]
    assert (age == 0).sum() > 0 and ((age == 1) | (age == 2)).sum() > 0          # live and cached arbiters are present
    env.set_state(rec)
    back = env.get_state()
    assert back.tobytes() == rec.tobytes()
    env2 = FutbolV1VecEnv(n, number_of_player=N, seed=seed, env_id_offset=off, dtype=torch_cuda.float64)
    env2.set_state(back)
    want = orc.rollout(110, actions=acts[250:], autoreset=2, n_threads=4)
    for t in range(110):
        a = torch_cuda.from_numpy(acts[250 + t]).cuda()
        obs, rew, done, _ = env.step(a)
        assert np.array_equal(obs.cpu().numpy(), want["obs"][t]) and np.array_equal(rew.cpu().numpy(), want["reward"][t]), t
        assert np.array_equal(done.cpu().numpy(), want["done"][t])
        obs2, rew2, _, _ = env2.step(a)
        assert torch_cuda.equal(obs, obs2) and torch_cuda.equal(rew, rew2)
    for field, idx, val in (("body", (3, 0, 1), np.nan), ("jn", (5, 2), 1e-200), ("owner_side", (1,), 2), ("last", (0, 0), stamp + 1)):
        bad = rec.copy()
        bad[field][idx] = val
        with pytest.raises(ValueError):
            env.set_state(bad)


# ---- against the reference's own Python (tests/golden/v1_golden.npz, see tests/test_oracle_v1_golden.py) ----------------
def _golden_v1_cases():
    import json
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "v1_golden.npz"))
    cases = {}
    for key in z.files:
        case, field = key.split("/")
        cases.setdefault(case, {})[field] = z[key]
    cases.pop("libm_fingerprint"), cases.pop("coverage")
    for c in cases.values():
        c["meta"] = json.loads(str(c["meta"]))
    return cases


def test_v1_golden_reference_traces(torch_cuda):
    """The CUDA path against traces of the UNMODIFIED reference v1 Python (run over the pymunk stand-in): every integer
    output of every step exact (done, goal / out-of-bounds / goal-side flags, possession side), observation and reward
    within 1e-9 relative (the kernel squares with x*x, CPython's float ** calls libm pow); north_star allows 1e-4."""
    from gym_futbol_b200 import FutbolV1VecEnv
    worst, steps_checked = 0.0, 0
    for name, c in _golden_v1_cases().items():
        m = c["meta"]
        N, T = m["number_of_player"], m["steps"]
        env = FutbolV1VecEnv(1, number_of_player=N, seed=m["seed"], env_id_offset=m["env_id"], total_time=m["kwargs"].get("total_time", 30),
                             dtype=torch_cuda.float64, auto_reset=False)
        obs0 = env.reset().cpu().numpy()[0]
        assert np.abs(obs0 - c["obs0"]).max() <= 1e-12, name
        acts = torch_cuda.from_numpy(c["action"]).cuda().reshape(T, 1, 2 * N)
        obs_l, rew_l, done_l, st_l = [], [], [], []
        for t in range(T):
            obs, rew, done, _ = env.step(acts[t])
            obs_l.append(obs.clone()); rew_l.append(rew.clone()); done_l.append(done.clone())
            st = env.get_state()[0]
            st_l.append((int(st["flags"]), int(st["owner_side"])))
            if int(st["flags"]) & 4:
                env.reset()
        obs = torch_cuda.stack(obs_l).cpu().numpy()[:, 0]
        rew = torch_cuda.stack(rew_l).cpu().numpy()[:, 0]
        done = torch_cuda.stack(done_l).cpu().numpy()[:, 0]
        assert np.array_equal(done, c["done"]), name
        assert np.array_equal(np.array([s[0] for s in st_l], np.uint8), c["flags"]), name
        assert np.array_equal(np.array([s[1] for s in st_l], np.uint8), c["owner_side"]), name
        for a, b in ((obs, c["obs"]), (rew, c["reward"])):
            worst = max(worst, float((np.abs(a - b) / np.maximum(1.0, np.abs(b))).max()))
        steps_checked += T
    assert worst <= 1e-9 and steps_checked >= 15000
    print("v1 golden: %d reference steps, max relative float error %.2e" % (steps_checked, worst))


def test_v1_golden_batch_members(torch_cuda):
    """'Env #k inside a batch': the eight batch_* golden envs (global ids 2000..2007) stepped as members of one batch, fused
    rollout with given actions: fp32 streams equal the reference's doubles to fp32 resolution, dones exact."""
    from gym_futbol_b200 import FutbolV1VecEnv
    cases = _golden_v1_cases()
    for N in (2, 5):
        members = [cases["batch_n%d_s3_e%d" % (N, e)] for e in range(2000, 2008)]
        T = members[0]["meta"]["steps"]
        env = FutbolV1VecEnv(40, number_of_player=N, seed=3, env_id_offset=1990)      # ids 1990..2029: members at 10..17
        env.reset()
        acts = np.zeros((T, 40, 2 * N), np.uint8)
        for k, c in enumerate(members):
            acts[:, 10 + k] = c["action"]
        obs, rew, done = env.rollout(T, actions=torch_cuda.from_numpy(acts).cuda())
        obs, rew, done = obs.cpu().numpy(), rew.cpu().numpy(), done.cpu().numpy()
        for k, c in enumerate(members):
            assert np.array_equal(done[:, 10 + k], c["done"])
            assert (np.abs(obs[:, 10 + k] - c["obs"]) <= 2e-7 * np.maximum(1.0, np.abs(c["obs"]))).all()
            assert (np.abs(rew[:, 10 + k] - c["reward"]) <= 2e-7 * np.maximum(1.0, np.abs(c["reward"]))).all()
