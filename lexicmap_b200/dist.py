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


def shard_hit_counts(rows, n_queries):
    """per-query number of genomes in ONE shard's rows (the `hits` column is constant over a query's rows): int32[n_queries]"""
    import numpy as np
    h = np.zeros(n_queries, np.int32)
    if len(rows):
        first = np.r_[True, rows["query"][1:] != rows["query"][:-1]]
        h[rows["query"][first]] = rows["hits"][first]
    return h


def allreduce_hits(counts, device=None):
    """the exchange step of a genome-sharded search: SUM over the shards of the per-query genome counts (what `lexicmap utils
    merge-search-results` computes offline, merge-search-results.go:143-153). NCCL all-reduce of int32[n_queries] when `device` is a CUDA
    device, gloo in the CPU tests; identity when no process group is initialised. Returns a torch tensor on `device`."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(counts)
    if device is not None:
        t = t.to(device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def apply_global_hits(rows, hits):
    """write the all-reduced genome counts into a shard's rows (TSV column 3)"""
    import numpy as np
    h = hits.cpu().numpy() if hasattr(hits, "cpu") else np.asarray(hits)
    if len(rows):
        rows["hits"] = h[rows["query"]]
    return rows


def merge_genome_shards(parts):
    """Merge the results of the same query batch searched against the genome shards of one index (`lmg_index_open(..., shard, n_shards)`:
    every shard holds all masks but only the seed values / genomes with dense_id % n_shards == shard; SURVEY.md §8e option 2, the offline
    equivalent is `lexicmap utils merge-search-results`, merge-search-results.go:143-153,318-320).
    parts: [(rows, seqids, cigars), ...] one per shard. Per query: `hits` becomes the number of genomes over all shards and the genomes are
    re-ordered exactly as the unsharded search orders them (stable by genome index in its batch, then stable by the best
    bitscore x pident of the genome, lib-index-search.go:1848-1853,2919-2921); rows inside a genome keep their order.
    Returns (rows, seqids, cigars)."""
    import numpy as np
    groups = {}   # query -> list of (bgi, sim, part, first_row, last_row)
    for pi, (rows, _, _) in enumerate(parts):
        n = len(rows)
        i = 0
        while i < n:
            j = i
            q, g = rows["query"][i], rows["genome"][i]
            while j < n and rows["query"][j] == q and rows["genome"][j] == g:
                j += 1
            sim = float(np.max(rows["bitscore"][i:j].astype(np.float64) * rows["pident"][i:j]))
            groups.setdefault(int(q), []).append((int(g), sim, pi, i, j))
            i = j
    out_rows, out_ids, out_cig = [], [], []
    for q in sorted(groups):
        gs = sorted(groups[q], key=lambda t: t[0])                 # dense order (batch-major, index in batch): the unsharded segment order
        gs = sorted(gs, key=lambda t: t[0] & 131071)               # stable, :1848-1853
        gs = sorted(gs, key=lambda t: -t[1])                       # stable, best similarity first, :2919-2921
        for g, _, pi, i, j in gs:
            rows, ids, cig = parts[pi]
            r = rows[i:j].copy()
            r["hits"] = len(gs)
            out_rows.append(r)
            out_ids += ids[i:j]
            out_cig += cig[i:j]
    dtype = parts[0][0].dtype
    return (np.concatenate(out_rows) if out_rows else np.zeros(0, dtype)), out_ids, out_cig
