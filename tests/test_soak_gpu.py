"""Large-scale bit-exactness soak: tens of thousands of environments for thousands of steps, final states compared
with the oracle (no per-step recording on the CPU side).  Catches what small cases cannot: rare draw counts, rare
branches (kicks from the touchline, simultaneous intercepts), and any operand outside the guard-free arithmetic's
domain would show up here as a mismatch."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bits(a):
    """The raw bytes of a float64 array: equal only if every value is the same BIT PATTERN (-0.0 != 0.0 here)."""
    return np.ascontiguousarray(a, dtype=np.float64).tobytes()


@pytest.mark.parametrize("random_opp", [False, True])
def test_v0_soak_final_state_bit_exact(random_opp):
    import torch
    from gym_futbol_b200 import FutbolVecEnv
    from oracle.v0 import OracleV0
    n, K, reps, seed, off = 32768, 250, 8, 77, 123456
    env = FutbolVecEnv(n, seed=seed, env_id_offset=off, random_opp=random_opp)
    env.reset()
    orc = OracleV0(n, seed=seed, env_id0=off, random_opp=random_opp, arith=0)
    for _ in range(reps):                               # 2000 steps: five 401-step episodes per env
        env.rollout(K, obs=False, reward=False, done=False)
        orc.rollout(K, actions=None, autoreset=2, n_threads=16, record=False)
    torch.cuda.synchronize()
    st = env.get_state()
    e = orc.envs
    assert _bits(st["rows"].reshape(n, 25)) == _bits(e["obs"][:, :5].reshape(n, 25))           # the sign of zero included
    assert np.array_equal(st["owner"], e["owner"].astype(np.uint8)) and np.array_equal(st["last_owner"], e["last_owner"].astype(np.uint8))
    assert np.array_equal(st["ai_score"], e["ai_score"]) and np.array_equal(st["opp_score"], e["opp_score"])
    assert np.array_equal(st["t_total"], e["t_total"]) and (st["t_total"] == K * reps).all()
    stats = env.read_stats()
    assert stats["env_steps"] == n * K * reps and stats["episodes"] == n * (K * reps // 401)


@pytest.mark.parametrize("N,n", [(2, 16384), (5, 4096)])
def test_v1_soak_final_state_bit_exact(N, n):
    import torch
    from gym_futbol_b200 import FutbolV1VecEnv
    from oracle.v1 import OracleV1
    K, reps, seed, off = 250, 6, 31, 5000
    env = FutbolV1VecEnv(n, number_of_player=N, seed=seed, env_id_offset=off)
    env.reset()
    orc = OracleV1(n, seed=seed, env_id0=off, number_of_player=N)
    for _ in range(reps):                               # 1500 steps: five 300-step episodes per env
        env.rollout(K, obs=False, reward=False, done=False)
        orc.rollout(K, actions=None, autoreset=2, n_threads=16, record=False)
    torch.cuda.synchronize()
    st = env.get_state()
    B = 2 * N + 1
    assert _bits(st["body"][:, :B, 0:2]) == _bits(orc.envs["p"][:, :B]) and _bits(st["body"][:, :B, 2:4]) == _bits(orc.envs["v"][:, :B])
    assert _bits(st["body"][:, :B, 4:6]) == _bits(orc.envs["vb"][:, :B])
    assert np.array_equal(st["owner_side"], orc.envs["owner_side"].astype(np.uint8)) and np.array_equal(st["ep_step"], orc.envs["ep_step"])
    stats = env.read_stats()
    assert stats["contacts_dropped"] == 0 and orc.envs["overflow"].sum() == 0 and stats["episodes"] == n * (K * reps // 300)


@pytest.mark.parametrize("cfg", [
    dict(seed=1, random_opp=False, one_goal_end=True, game_time=25.0),
    dict(seed=2, random_opp=True, only_reward_goal=True, game_time=12.5),
    dict(seed=3, random_opp=False, player_speed=9.0, shoot_speed=24, game_time=40.0),
    dict(seed=4, random_opp=True, player_speed=15.5, shoot_speed=17, one_goal_end=True, game_time=7.3),
    dict(seed=2 ** 40 + 5, random_opp=False, player_speed=12.0, shoot_speed=20, game_time=0.0),
    dict(seed=6, random_opp=False, only_reward_goal=True, one_goal_end=True, game_time=60.0),
])
def test_v0_constructor_options_at_scale(cfg):
    """Every constructor option of FutbolEnv (futbol_env.py:134-138) away from its default, 4096 envs x 600 fused steps
    with recorded rewards and dones, and a 64-bit seed: rewards and dones every step, final state bit pattern."""
    import torch
    from gym_futbol_b200 import FutbolVecEnv
    from oracle.v0 import OracleV0
    n, K, reps, off = 4096, 150, 4, 31337
    env = FutbolVecEnv(n, env_id_offset=off, **cfg)
    env.reset()
    orc = OracleV0(n, env_id0=off, arith=0, **{k: (float(v) if k in ("player_speed", "shoot_speed") else v) for k, v in cfg.items()})
    for _ in range(reps):
        _, rew, done = env.rollout(K, obs=False)
        want = orc.rollout(K, actions=None, autoreset=2, n_threads=16)
        assert np.array_equal(done.cpu().numpy(), want["done"])
        assert np.array_equal(rew.cpu().numpy(), want["reward"].astype(np.float32))
    torch.cuda.synchronize()
    st = env.get_state()
    e = orc.envs
    assert _bits(st["rows"].reshape(n, 25)) == _bits(e["obs"][:, :5].reshape(n, 25))
    assert np.array_equal(st["ai_score"], e["ai_score"]) and np.array_equal(st["opp_score"], e["opp_score"])
    assert np.array_equal(st["owner"], e["owner"].astype(np.uint8)) and np.array_equal(st["last_owner"], e["last_owner"].astype(np.uint8))


@pytest.mark.parametrize("N,n,total_time", [(3, 2048, 30.0), (4, 1024, 12.5), (6, 1024, 30.0), (7, 512, 5.0), (8, 512, 30.0), (9, 256, 30.0),
                                            (10, 512, 30.0), (1, 4096, 3.1)])
def test_v1_every_team_size_and_time_limit(N, n, total_time):
    """Team sizes the other GPU tests skip (3, 4, 6-9: the three formation families of team.py:52-112) and time limits away
    from the default: observations, rewards and dones of a fused rollout against the oracle, final bodies bit pattern."""
    import torch
    from gym_futbol_b200 import FutbolV1VecEnv
    from oracle.v1 import OracleV1
    K, seed, off = 340, 21, 808
    env = FutbolV1VecEnv(n, number_of_player=N, seed=seed, env_id_offset=off, total_time=total_time)
    orc = OracleV1(n, seed=seed, env_id0=off, number_of_player=N, total_time=total_time)
    env.reset()
    obs, rew, done = env.rollout(K)
    want = orc.rollout(K, actions=None, autoreset=2, n_threads=16)
    assert np.array_equal(done.cpu().numpy(), want["done"]) and int(want["done"].sum()) > 0
    assert np.array_equal(rew.cpu().numpy(), want["reward"].astype(np.float32))
    assert np.array_equal(obs.cpu().numpy(), want["obs"].astype(np.float32))
    torch.cuda.synchronize()
    st = env.get_state()
    B = 2 * N + 1
    assert _bits(st["body"][:, :B, 0:2]) == _bits(orc.envs["p"][:, :B]) and _bits(st["body"][:, :B, 2:4]) == _bits(orc.envs["v"][:, :B])
    assert env.read_stats()["contacts_dropped"] == 0
