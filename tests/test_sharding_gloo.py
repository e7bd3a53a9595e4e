"""N > 1 host logic on CPU: two gloo ranks own contiguous shards of the global env ids, step them with the
ORACLE (test infrastructure; there is no GPU here) and exchange only the statistics record.  Checks that
(i) the shards tile the id range, (ii) trajectories do not depend on the sharding (Philox is keyed by the
global env id), (iii) allreduce_stats sums the per-rank records."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_tiles_the_range():
    from gym_futbol_b200.sharding import shard
    for total in (0, 1, 7, 1 << 20, 1000003):
        for world in (1, 2, 3, 4, 8):
            spans = [shard(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        shard(10, 2, 2)


def _worker(rank, world, port, total, steps, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from gym_futbol_b200.sharding import allreduce_stats, shard
    from oracle import philox
    from oracle.v0 import OracleV0
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    first, count = shard(total, rank, world)
    orc = OracleV0(count, seed=7, env_id0=first, random_opp=False, game_time=6.0, arith=0)
    acts = philox.actions_table(7, np.arange(first, first + count), 0, steps)
    out = orc.rollout(steps, actions=acts, autoreset=2, n_threads=2)
    local = {"reward_sum": float(out["reward"].sum()), "env_steps": count * steps, "episodes": int(out["done"].sum()),
             "goals_ai": int(((out["flags"] & 1) > 0).sum()), "goals_opp": 0, "out_of_field": int(((out["flags"] & 2) > 0).sum())}
    total_stats = allreduce_stats(local)
    q.put((rank, first, count, out["obs"][-1].copy(), out["done"].sum(0), local, total_stats))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gloo_ranks_match_one_process():
    import multiprocessing as mp
    from oracle import philox
    from oracle.v0 import OracleV0
    total, steps, world = 257, 150, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=180) for _ in procs), key=lambda g: g[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    orc = OracleV0(total, seed=7, env_id0=0, random_opp=False, game_time=6.0, arith=0)
    acts = philox.actions_table(7, np.arange(total), 0, steps)
    want = orc.rollout(steps, actions=acts, autoreset=2, n_threads=2)
    obs = np.concatenate([g[3] for g in got])
    assert got[0][1] == 0 and got[0][2] + got[1][2] == total and got[1][1] == got[0][2]
    assert np.array_equal(obs, want["obs"][-1])                       # sharding-invariant trajectories
    assert np.array_equal(np.concatenate([g[4] for g in got]), want["done"].sum(0))
    for g in got:                                                     # every rank holds the global sums
        assert g[6]["env_steps"] == total * steps
        assert g[6]["episodes"] == int(want["done"].sum())
        assert abs(g[6]["reward_sum"] - float(want["reward"].sum())) <= 1e-6 * max(1.0, abs(float(want["reward"].sum())))
        assert g[6]["out_of_field"] == got[0][5]["out_of_field"] + got[1][5]["out_of_field"]
