"""Multi-GPU plumbing: environments shard trivially, one process per GPU, no collective on the step path.

`shard_range` gives every rank a contiguous slice of the global env index range; the slice start is passed to
the env as `env_id_offset`, so the Philox piece streams are keyed by GLOBAL env id and results do not depend
on the number of GPUs.  The only collective is the optional all-reduce (sum) of the 4-double episode
statistics vector -- NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of global env ids owned by `rank` (first `n_total % world` ranks get one extra)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(int(n_total), int(world_size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def allreduce_episode_stats(stats: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Sum the [episodes, sum_return, sum_length, sum_lines] vector over all ranks (in place)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def summarize(stats: torch.Tensor) -> dict:
    s = stats.detach().cpu().tolist()
    ep = max(s[0], 1.0)
    return {"episodes": s[0], "mean_return": s[1] / ep, "mean_length": s[2] / ep, "mean_lines": s[3] / ep}


def make_sharded_env(n_total: int, rank: Optional[int] = None, world_size: Optional[int] = None, **env_kwargs):
    """Build this rank's shard of a global `n_total`-env job (uses torch.distributed's rank/world when initialised)."""
    from .envs.tetris import Tetris

    if rank is None:
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    start, stop = shard_range(n_total, rank, world_size)
    return Tetris(num_envs=stop - start, env_id_offset=start, **env_kwargs)
