"""Multi-GPU plumbing: environments shard by contiguous global index, one process per GPU; nothing on the step
path is collective.  The only exchange is the int64[8] episode-statistics vector."""
from __future__ import annotations

import os
from typing import Tuple


def rank_world() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (1-process defaults)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard(total_envs: int, rank: int, world: int) -> Tuple[int, int]:
    """(env_offset, num_envs) of `rank` when `total_envs` are split into `world` contiguous ranges."""
    base, rem = divmod(int(total_envs), int(world))
    n = base + (1 if rank < rem else 0)
    off = rank * base + min(rank, rem)
    return off, n


def all_reduce_stats(stats):
    """Sum the per-rank episode statistics (works for NCCL on CUDA tensors and gloo on CPU tensors)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats
