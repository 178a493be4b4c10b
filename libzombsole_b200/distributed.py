"""Multi-GPU plumbing: environments are independent, so they shard by contiguous global index
range — one batch per GPU, one process per GPU — and the step path has NO collective.  Draw
counters use the GLOBAL env index (philox.py), so the trajectory of env g does not depend on how
many ranks there are or which rank owns it.  The only collective is an optional all-reduce of the
per-rank episode statistics (a 32-byte tensor) at reporting time."""
import torch
import torch.distributed as dist


def shard_envs(total_envs, rank, world_size):
    """Contiguous shard of `total_envs` for `rank`: (env_index_base, num_envs).  The first
    `total_envs % world_size` ranks get one extra env."""
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    q, r = divmod(int(total_envs), int(world_size))
    n = q + (1 if rank < r else 0)
    base = rank * q + min(rank, r)
    return base, n


def all_reduce_stats(stats):
    """Sum the int64 [4] episode statistics (episodes, wins, length sum, zombie-death sum) over ranks.
    Works with any initialised backend (NCCL on GPUs, gloo in the CPU tests); a no-op without one."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats
