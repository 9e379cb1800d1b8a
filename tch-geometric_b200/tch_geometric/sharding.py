"""Seed / walker sharding for multi-GPU runs (one process per GPU, replicated CSC, no collective on
the data path).  The RNG counters carry the *global* batch / walker index (batch_base, walker_base),
so the union of all ranks' outputs is bit-identical to a single-GPU run over the whole job."""
from typing import Tuple


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) share of n units for `rank`; sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def reduce_job(elapsed_ms: float, units: float, device=None):
    """Whole-job figures over all ranks: (max elapsed over ranks, sum of units).  Single process: identity."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return elapsed_ms, units
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
    u = torch.tensor([units], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), float(u.item())
