"""Full-size checks at BASELINE.json configs[1] (1,000 genomes x 1 Mbp, 10,000 x 1-kb queries): size-independent properties of the CUDA
path plus bit-exact agreement with the oracle on a random sample of the batch. The index / query files are shared with bench.py's cache."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

WORK = os.environ.get("LMG_BENCH_DIR", "/tmp/lmg_bench")


@pytest.fixture(scope="module")
def full_case():
    from conftest import make_index, make_queries
    from oracle_binding import read_fasta
    import lexicmap_b200
    os.makedirs(WORK, exist_ok=True)
    idx = make_index(WORK, "c2d_50x20x1000000", "50,20,1000000,20260924,20", chunks=16)   # same directory and arguments as bench.py::ensure_c2 (desert-filled, the writer's default)
    ids, seqs = read_fasta(make_queries(WORK, idx, "c2_fullsize_q", 10000, 1000, seed=20260925))
    return idx, lexicmap_b200.Index(idx, device=0), seqs


def _same(a, b):
    (ra, sa, ca), (rb, sb, cb) = a, b
    assert len(ra) == len(rb)
    for f in ra.dtype.names:
        if f not in ("cigar_off", "pad", "pad0"):
            assert np.array_equal(ra[f], rb[f]), f
    assert sa == sb and ca == cb


def test_full_size_lanes_idempotence_and_bounds(full_case):
    _, g, seqs = full_case
    one = g.search(seqs, g.default_params(output_seq=1, lanes=1))
    six = g.search(seqs, g.default_params(output_seq=1, lanes=6))
    again = g.search(seqs, g.default_params(output_seq=1))          # automatic lane count, second pass over warm arenas
    _same(one, six)
    _same(one, again)
    r = one[0]
    assert len(r) > 100000
    assert np.all(np.diff(r["query"].astype(np.int64)) >= 0), "rows are grouped by query in input order"
    qlen = np.array([len(s) for s in seqs])[r["query"]]
    assert np.all((r["qb"] >= 0) & (r["qb"] <= r["qe"]) & (r["qe"] < qlen))          # 0-based inclusive in the row struct, +1 in the TSV
    assert np.all((r["tb"] >= 0) & (r["tb"] <= r["te"]) & (r["te"] < r["seq_len"]))
    assert np.all((r["matches"] <= r["alen"]) & (r["gaps"] <= r["alen"]) & (r["pident"] <= 100.0) & (r["pident"] >= 70.0))
    assert np.all(r["qe"] - r["qb"] + 1 <= r["alen"])
    # every synthetic query was cut from an indexed genome with <= 10 % substitutions: nearly all must be found
    assert len(np.unique(r["query"])) >= 0.99 * len(seqs)
    # the first HSP of the first cluster of every (query, genome) group is numbered 1
    first = np.r_[True, (r["query"][1:] != r["query"][:-1]) | (r["genome"][1:] != r["genome"][:-1])]
    assert np.all(r["cls"][first] == 1) and np.all(r["hsp"][first] == 1)


def test_full_size_batch_matches_oracle(full_case):
    """all 10,000 queries of the benchmark batch, every column of every row, CIGARs and sequence ids included, against the CPU oracle"""
    idx, g, seqs = full_case
    from oracle_binding import Oracle
    o = Oracle(idx)
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 8)
    orr, os_, oc = o.search(seqs, o.default_params(output_seq=1), threads=threads)
    gr, gs, gc = g.search(seqs, g.default_params(output_seq=1))
    assert len(orr) > 100000
    _same((gr, gs, gc), (orr, os_, oc))
    rng = np.random.default_rng(7)
    pick = sorted(rng.choice(len(seqs), size=96, replace=False).tolist())
    sub = [seqs[i] for i in pick]
    _same(g.search(sub, g.default_params(output_seq=1, lanes=3)), o.search(sub, o.default_params(output_seq=1), threads=threads))   # forced lanes on a small batch
