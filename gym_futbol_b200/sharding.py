"""Env sharding over ranks (one process per GPU).  Environments share nothing (reference: every FutbolEnv
is an independent object), so the step path has NO collective: a rank owns a contiguous block of global
env ids and the Philox streams are keyed by the GLOBAL id, which makes trajectories independent of the
number of ranks.  The only exchange is the optional sum of the rollout statistics (a 64-byte record).
"""
from __future__ import annotations

import numpy as np

STAT_KEYS = ("reward_sum", "env_steps", "episodes", "goals_ai", "goals_opp", "out_of_field")


def shard(total_envs: int, rank: int, world: int):
    """(first global env id, number of envs) of `rank`: contiguous blocks, remainder on the first ranks."""
    if world <= 0 or not 0 <= rank < world or total_envs < 0:
        raise ValueError("bad shard request: total=%r rank=%r world=%r" % (total_envs, rank, world))
    base, rem = divmod(int(total_envs), int(world))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def allreduce_stats(stats: dict, group=None, device=None):
    """Sum a rollout-statistics dict (FutbolVecEnv.read_stats) over all ranks.  Works with any backend
    (NCCL on CUDA tensors, gloo on CPU tensors); returns the local dict unchanged when not distributed."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dict(stats)
    # the integer counters are summed as integers (exact beyond 2**53); reward_sum in float64 (its last bits depend on the
    # order of the device-side atomic adds and of this reduction)
    counts = torch.tensor([int(stats[k]) for k in STAT_KEYS[1:]], dtype=torch.int64, device=device)
    reward = torch.tensor([float(stats["reward_sum"])], dtype=torch.float64, device=device)
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(reward, op=dist.ReduceOp.SUM, group=group)
    out = {"reward_sum": float(reward.item())}
    out.update({k: int(v) for k, v in zip(STAT_KEYS[1:], counts.tolist())})
    return out
