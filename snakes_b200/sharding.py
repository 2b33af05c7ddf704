"""Multi-GPU host logic: contiguous sharding of the global env range and the one collective.

The reference's analogue is `ncpu` worker processes (train_snake.py:31,41).  Here GPU g of G owns
global env ids [base, base + count); stepping needs no exchange; episode statistics (8 doubles)
are summed with a single all-reduce (NCCL on GPUs; any torch.distributed backend works).
"""
import torch


def shard_range(total_envs, rank, world):
    """(base, count) of rank's contiguous slice; the first `total_envs % world` ranks get one more."""
    if not 0 <= rank < world:
        raise ValueError("rank %d outside world %d" % (rank, world))
    q, r = divmod(int(total_envs), int(world))
    count = q + (1 if rank < r else 0)
    base = rank * q + min(rank, r)
    return base, count


def all_reduce_stats(stats):
    """Sum the per-rank statistics vector over the default process group (no-op when single process)."""
    if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
        torch.distributed.all_reduce(stats)
    return stats


def make_sharded_env(total_envs, rank=None, world=None, **kwargs):
    """This rank's SnakeVecEnv over its slice of `total_envs` global envs (device = local rank)."""
    import os
    from .vec_env import SnakeVecEnv
    if world is None:
        world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank is None:
        rank = int(os.environ.get("RANK", "0"))
    base, count = shard_range(total_envs, rank, world)
    kwargs.setdefault("device", int(os.environ.get("LOCAL_RANK", "0")))
    return SnakeVecEnv(count, env_id_base=base, **kwargs)
