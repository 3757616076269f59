"""Environment sharding across the GPUs of one box.

Games are independent: rank r of W owns the contiguous slice of *global* environment ids returned by
``shard_range``; the Philox counter uses the global id, so per-environment trajectories do not depend on W.
The only collective is a SUM all-reduce of the int64[32] episode-statistics vector (NCCL on GPUs, gloo in the
CPU unit tests), once per measurement window.
"""
from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(total_envs: int, rank: int, world: int) -> Tuple[int, int]:
    """(first global env id, number of envs) of `rank`; slices are contiguous, ordered and cover [0, total)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(int(total_envs), int(world))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def all_reduce_stats(stats: torch.Tensor, async_op: bool = False):
    """SUM over ranks of the statistics vector, in place on `stats` (a fresh reduction of the device rows when it comes
    from `env.stats`); identity when not distributed.  async_op=True: returns (tensor, work) -- `work.wait()` orders
    the current stream after the collective without blocking the host (None when there is nothing to wait for)."""
    work = None
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        work = dist.all_reduce(stats, op=dist.ReduceOp.SUM, async_op=async_op)
    return (stats, work) if async_op else stats
