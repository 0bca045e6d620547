"""Multi-GPU host logic: the path shards by query with no exchange step; torch.distributed is plumbing only."""
import bisect


def shard_queries(lengths, rank, world):
    """contiguous slice [lo, hi) of the batch for `rank`, balanced by cumulative query length"""
    total = sum(lengths)
    cum, acc = [], 0
    for x in lengths:
        acc += x
        cum.append(acc)
    lo = 0 if rank == 0 else bisect.bisect_left(cum, total * rank / world)
    hi = len(lengths) if rank == world - 1 else bisect.bisect_left(cum, total * (rank + 1) / world)
    return lo, max(lo, hi)


def reduce_counters(rows, bases, step_ms, device=None):
    """the one collective of the path: SUM of rows / bases, MAX of the step time (NCCL on GPUs, gloo in the CPU tests)"""
    import torch
    import torch.distributed as dist
    t = torch.tensor([rows, bases], dtype=torch.float64, device=device)
    m = torch.tensor([step_ms], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return float(t[0]), float(t[1]), float(m[0])
