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


def test_merge_genome_shards_restores_unsharded_rows(oracle_small, small_queries):
    """host merge of genome-sharded results (SURVEY.md §8e option 2): split the oracle's rows by genome into 3 pseudo-shards
    (per-shard `hits`), merge, compare with the original ordering and hits column"""
    import numpy as np
    from lexicmap_b200.dist import merge_genome_shards
    ids, seqs = small_queries
    rows, sid, cig = oracle_small.search(seqs, oracle_small.default_params(output_seq=1))
    assert len(rows) > 50
    parts = []
    for sh in range(3):
        keep = [i for i in range(len(rows)) if ((int(rows["genome"][i]) >> 17) * 7 + (int(rows["genome"][i]) & 131071)) % 3 == sh]
        r = rows[keep].copy()
        for q in np.unique(r["query"]):
            m = r["query"] == q
            r["hits"][m] = len(np.unique(r["genome"][m]))
        parts.append((r, [sid[i] for i in keep], [cig[i] for i in keep]))
    mr, ms, mc = merge_genome_shards(parts)
    assert ms == sid and mc == cig
    for f in rows.dtype.names:
        if f not in ("cigar_off", "pad", "pad0"):
            assert np.array_equal(mr[f], rows[f]), f


def _shard_worker(rank, world, port, idx_dir, fa, out):
    """one rank = one genome shard: its share of the rows, NCCL-shaped exchange (gloo here): all-reduce of the per-query genome counts,
    then rank 0 gathers the shard rows and merges them"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_binding import Oracle, read_fasta
    from lexicmap_b200.dist import shard_hit_counts, allreduce_hits, apply_global_hits, merge_genome_shards
    ids, seqs = read_fasta(fa)
    o = Oracle(idx_dir)
    rows, sid, cig = o.search(seqs, o.default_params(output_seq=1))
    keep = [i for i in range(len(rows)) if (int(rows["genome"][i]) & 131071) % world == rank]      # this shard's genomes
    r = rows[keep].copy()
    for q in np.unique(r["query"]):
        m = r["query"] == q
        r["hits"][m] = len(np.unique(r["genome"][m]))
    hits = allreduce_hits(shard_hit_counts(r, len(seqs)))
    r = apply_global_hits(r, hits)
    gathered = [None] * world
    dist.gather_object((r, [sid[i] for i in keep], [cig[i] for i in keep]), gathered if rank == 0 else None, dst=0)
    if rank == 0:
        mr, ms, mc = merge_genome_shards(gathered)
        ok = ms == sid and mc == cig and all(np.array_equal(mr[f], rows[f]) for f in rows.dtype.names if f not in ("cigar_off", "pad", "pad0"))
        ok = ok and np.array_equal(hits.numpy()[rows["query"]], rows["hits"].astype(np.int32))     # the all-reduced counts alone already give the right column
        np.save(out, np.array([int(ok), len(mr)]))
    dist.destroy_process_group()


def test_genome_sharded_ranks_allreduce_hits_and_merge_gloo(small_index, workdir, tmp_path):
    """world_size 2 (gloo): the N>1 path of `bench.py --config c3` — shard rows, all-reduce(sum) of the per-query genome counts, gather + merge on rank 0"""
    from conftest import make_queries
    fa = make_queries(workdir, small_index, "small_q", 24, 800)
    out = str(tmp_path / "m.npy")
    mp.spawn(_shard_worker, args=(2, 29531, small_index, fa, out), nprocs=2, join=True)
    ok, n = np.load(out)
    assert ok == 1 and n > 50
