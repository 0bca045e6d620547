"""world_size-2 gloo test of the multi-GPU host logic (query sharding + the one counter reduction), CPU only."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lexicmap_b200.dist import shard_queries, reduce_counters
    lens = [100 + (i * 37) % 900 for i in range(1001)]
    lo, hi = shard_queries(lens, rank, world)
    rows, bp, tmax = reduce_counters(float(hi - lo), float(sum(lens[lo:hi])), 10.0 + rank)
    if rank == 0:
        np.save(out, np.array([rows, bp, tmax, lo, hi]))
    dist.destroy_process_group()


def test_query_sharding_and_counter_reduction_gloo(tmp_path):
    out = str(tmp_path / "r.npy")
    mp.spawn(_worker, args=(2, 29517, out), nprocs=2, join=True)
    rows, bp, tmax, lo, hi = np.load(out)
    lens = [100 + (i * 37) % 900 for i in range(1001)]
    assert rows == 1001 and bp == sum(lens) and tmax == 11.0 and lo == 0


def test_shards_are_contiguous_and_balanced():
    from lexicmap_b200.dist import shard_queries
    lens = [1000] * 10 + [50000] + [1000] * 10
    cuts = [shard_queries(lens, r, 4) for r in range(4)]
    assert cuts[0][0] == 0 and cuts[-1][1] == len(lens)
    for a, b in zip(cuts, cuts[1:]):
        assert a[1] == b[0]
