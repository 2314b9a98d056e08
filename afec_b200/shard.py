"""File-batch sharding across ranks: files are independent units, every rank owns a
contiguous-by-cost slice of the file list; the only cross-rank traffic is the barrier and the reduction
of timings / totals (no data-path collective)."""
from __future__ import annotations


def shard_by_cost(costs, world: int) -> list:
    """Greedy longest-processing-time partition of file indices into `world` shards with balanced total
    cost (cost = decoded duration or byte size).  Deterministic; returns a list of sorted index lists."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0.0] * world
    shards = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += costs[i]
    return [sorted(s) for s in shards]


def my_shard(costs, rank: int, world: int) -> list:
    return shard_by_cost(costs, world)[rank]


def reduce_scalar(x: float, op: str, dist=None, device=None) -> float:
    """max / sum of a Python float over ranks (identity when torch.distributed is not initialised)."""
    if dist is None or not dist.is_initialized():
        return float(x)
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(units_local: float, seconds_local: float, dist=None, device=None) -> float:
    """Whole-job throughput = units of all ranks / max-over-ranks time (bench.py contract)."""
    units = reduce_scalar(units_local, "sum", dist, device)
    secs = reduce_scalar(seconds_local, "max", dist, device)
    return units / secs if secs > 0 else 0.0
