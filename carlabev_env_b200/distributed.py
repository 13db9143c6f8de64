"""Multi-GPU plumbing: one process per GPU, environments sharded by contiguous index ranges.

Environments are independent (the reference's SyncVectorEnv is a serial loop, envs/__init__.py:116-119),
so the data path has NO collective.  The only exchange is an all-reduce (sum) of the small
episode-statistics vector (fields of stats.py:127-148; CBEV_S_* in include/cbev.h) per logging
interval -- NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations


def shard_range(num_envs_total: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous [lo, hi) env index range of `rank`."""
    if num_envs_total % world_size != 0:
        raise ValueError(f"num_envs={num_envs_total} must be divisible by world_size={world_size}")
    per = num_envs_total // world_size
    return rank * per, (rank + 1) * per


def allreduce_stats(stats, group=None):
    """Sum the CBEV_STATS_FIELDS vector over all ranks in place and return it."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def summarize_stats(stats) -> dict:
    """Global means from the reduced vector (what Stats.get_episode_info reports per env)."""
    s = [float(v) for v in stats.tolist()]
    n = max(s[0], 1.0)
    causes = ("none", "ckpt", "collision", "success", "out_of_bounds", "off_road", "max_actions", "unknown")
    out = {"episodes": s[0], "mean_return": s[1] / n, "mean_length": s[2] / n, "mean_speed": s[3] / n,
           "env_steps": s[-1]}
    for i, c in enumerate(causes):
        out[f"rate_{c}"] = s[4 + i] / n
    for i, k in enumerate(("accel_long", "accel_lat", "jerk_long", "jerk_lat", "yaw_rate", "yaw_acc")):
        out[f"mean_abs_{k}"] = s[12 + i] / n
    out["comfort_violation_rate"] = s[18] / n
    out["harsh_brake_rate"] = s[19] / n
    return out
