"""GPU parity, stage by stage: CUDA path (through the C ABI) vs the CPU oracle on the same seeded inputs. Bit-exact."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_mask_capture_matches_oracle(gpu_small, oracle_small, small_queries):
    ids, seqs = small_queries
    m = gpu_small.info.masks
    gk, gn, gl, gs = gpu_small.mask(seqs)
    ok, on, ol, os_ = oracle_small.mask(seqs, m)
    assert np.array_equal(gk, ok), "captured k-mers differ"
    assert np.array_equal(gn, on), "number of query locations differ"
    assert np.array_equal(gl, ol), "first query location differs"
    assert gs.shape == os_.shape and np.array_equal(gs, os_), "suffix (reversed k-mer) probes differ"


def test_mask_fast_equals_bruteforce_definition(oracle_small, small_queries):
    ids, seqs = small_queries
    m = 20000
    a = oracle_small.mask(seqs[:2], m)
    b = oracle_small.mask(seqs[:2], m, bruteforce=True)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_anchors_match_oracle(gpu_small, oracle_small, small_queries):
    ids, seqs = small_queries
    ga = gpu_small.anchors(seqs)
    oa = oracle_small.anchors(seqs)
    assert len(ga) == len(oa) and len(ga) > 100
    assert ga.tobytes() == oa.tobytes(), "anchor multiset differs"


def test_chains_match_oracle(gpu_small, oracle_small, small_queries):
    ids, seqs = small_queries
    gc = gpu_small.chains(seqs)
    oc = oracle_small.chains(seqs)
    assert len(gc) == len(oc) and len(gc) > 20
    for f in gc.dtype.names:
        assert np.array_equal(gc[f], oc[f]), "chain field %s differs" % f


def test_chains_top_n(gpu_small, oracle_small, small_queries):
    ids, seqs = small_queries
    p = dict(top_n_genomes=2, top_n_chains=1)
    gc = gpu_small.chains(seqs, gpu_small.default_params(**p))
    oc = oracle_small.chains(seqs, oracle_small.default_params(**p))
    assert gc.tobytes() == oc.tobytes()


def test_pseudoalignment_matches_oracle(gpu_small, oracle_small, small_queries):
    """a9-a12 stage-wise: target windows of the chains, SeqComparator.Compare anchors, ClearSubstrPairs / TrimSubStrPairs, Chainer2 regions"""
    ids, seqs = small_queries
    for kw in (dict(), dict(top_n_chains=1, align_min_len=80, align_max_gap=10)):
        ga = np.sort(gpu_small.pseudoalign(seqs, gpu_small.default_params(**kw)), order=["query", "genome", "t_begin", "t_end", "rc", "qb", "qe", "tb", "te", "aligned_q", "aligned_t", "matched", "n_anchors"])
        oa = np.sort(oracle_small.pseudoalign(seqs, oracle_small.default_params(**kw)), order=["query", "genome", "t_begin", "t_end", "rc", "qb", "qe", "tb", "te", "aligned_q", "aligned_t", "matched", "n_anchors"])
        assert len(oa) > 50 and ga.tobytes() == oa.tobytes()


def _rows_equal(gr, orr, gs, os_, gc=None, oc=None):
    assert len(gr) == len(orr), "row count differs: gpu %d oracle %d" % (len(gr), len(orr))
    for f in gr.dtype.names:
        if f in ("cigar_off", "pad", "pad0"):
            continue
        if gr.dtype[f].kind == "f":
            assert np.array_equal(gr[f], orr[f]), "float column %s differs (must be bit-identical: same double ops on host)" % f
        else:
            assert np.array_equal(gr[f], orr[f]), "column %s differs" % f
    assert gs == os_
    if gc is not None:
        assert gc == oc


def test_search_rows_match_oracle(gpu_small, oracle_small, small_queries):
    ids, seqs = small_queries
    gr, gs, gc = gpu_small.search(seqs, gpu_small.default_params(output_seq=1))
    orr, os_, oc = oracle_small.search(seqs, oracle_small.default_params(output_seq=1))
    assert len(orr) > 50
    _rows_equal(gr, orr, gs, os_, gc, oc)
    assert gpu_small.last_align_text == oracle_small.last_align_text   # qseq / sseq / align columns of the -a output


@pytest.mark.parametrize("lanes", [2, 3, 5])
def test_search_lanes_match_oracle(gpu_small, oracle_small, small_queries, lanes):
    """the batch split into concurrent sub-batches (own stream, arena and host thread each) gives the same rows in the same order;
    both the host-buffer and the staged entry point"""
    ids, seqs = small_queries
    orr, os_, oc = oracle_small.search(seqs, oracle_small.default_params(output_seq=1))
    gr, gs, gc = gpu_small.search(seqs, gpu_small.default_params(output_seq=1, lanes=lanes))
    _rows_equal(gr, orr, gs, os_, gc, oc)
    assert int(gpu_small.timing()[0][12]) == min(lanes, len(seqs))
    os.environ["LMG_LANES"] = str(lanes)
    try:
        q = gpu_small.stage(seqs)
    finally:
        del os.environ["LMG_LANES"]
    gr, gs, gc = gpu_small.search_staged(q, gpu_small.default_params(output_seq=1))
    gpu_small.free_staged(q)
    _rows_equal(gr, orr, gs, os_, gc, oc)


def test_oversized_batch_is_halved(gpu_small, oracle_small, small_queries):
    """a sub-batch whose pseudo-alignment slot space would overflow is split in two and retried (recursively); rows are unchanged.
    LMG_SLOT_LIMIT lowers the 2^31 limit so that the 29-query batch has to be halved several times."""
    ids, seqs = small_queries
    orr, os_, oc = oracle_small.search(seqs, oracle_small.default_params(output_seq=1))
    os.environ["LMG_SLOT_LIMIT"] = "60000"
    try:
        gr, gs, gc = gpu_small.search(seqs, gpu_small.default_params(output_seq=1))
        q = gpu_small.stage(seqs)
        gr2, gs2, gc2 = gpu_small.search_staged(q, gpu_small.default_params(output_seq=1, lanes=2))
        gpu_small.free_staged(q)
    finally:
        del os.environ["LMG_SLOT_LIMIT"]
    _rows_equal(gr, orr, gs, os_, gc, oc)
    _rows_equal(gr2, orr, gs2, os_, gc2, oc)


def test_search_filters_match_oracle(gpu_small, oracle_small, small_queries):
    ids, seqs = small_queries
    kw = dict(min_qcov_hsp=50.0, min_pident=80.0, min_qcov_genome=60.0, top_n_genomes=3, max_evalue=1e-20)
    gr, gs, _ = gpu_small.search(seqs, gpu_small.default_params(**kw))
    orr, os_, _ = oracle_small.search(seqs, oracle_small.default_params(**kw))
    _rows_equal(gr, orr, gs, os_)


def test_wfa_batch_matches_oracle(oracle_small):
    import random
    from lexicmap_b200.api import wfa_batch
    rnd = random.Random(7)
    pairs = []
    for n in (1, 5, 40, 300, 1500):
        for div in (0.0, 0.02, 0.1, 0.25):
            q = "".join(rnd.choice("ACGT") for _ in range(n))
            t = []
            for c in q:
                r = rnd.random()
                if r < div / 3:
                    continue
                if r < 2 * div / 3:
                    t.append(rnd.choice("ACGT"))
                if r < div:
                    c = rnd.choice("ACGT")
                t.append(c)
            pairs.append((q, "".join(t) or "A"))
    pairs += [("ACGT", "TTTT"), ("A", "ACGTACGT"), ("ACGTACGTAA", "A"), ("AAAAAAAAAA", "AAAAA"), ("ACGTTTGACA" * 30, "ACGTTGACA" * 30)]
    for adaptive in (1, 0):   # WFA-adaptive reduction (reference default) and the exact algorithm
        assert wfa_batch(pairs, adaptive=adaptive) == oracle_small.wfa(pairs, adaptive), "adaptive=%d" % adaptive


def test_search_exact_wfa_matches_oracle(gpu_small, oracle_small, small_queries):
    ids, seqs = small_queries
    gr, gs, gc = gpu_small.search(seqs, gpu_small.default_params(output_seq=1, wfa_adaptive=0))
    orr, os_, oc = oracle_small.search(seqs, oracle_small.default_params(output_seq=1, wfa_adaptive=0))
    _rows_equal(gr, orr, gs, os_, gc, oc)


@pytest.fixture(scope="module")
def long_case(workdir):
    """longer queries: 3-kb and 12-kb reads (12 kb: k-mer table > 200 KB -> L2 pseudo-alignment kernel, windows >= 10 kb -> min prefix 13)
    against a 2 x 3 x 150-kb index; divergent enough that some alignments leave the fast WFA kernel."""
    from conftest import make_index, make_queries
    from oracle_binding import Oracle, read_fasta
    import lexicmap_b200
    idx = make_index(workdir, "longcase", "2,3,150000,11,2")
    _, s3 = read_fasta(make_queries(workdir, idx, "long_q3k", 6, 3000, seed=5))
    _, s12 = read_fasta(make_queries(workdir, idx, "long_q12k", 3, 12000, seed=6, max_sub=0.06, max_indel=0.02))
    _, s20 = read_fasta(make_queries(workdir, idx, "long_q20k", 1, 20000, seed=7, max_sub=0.02, max_indel=0.005))
    return lexicmap_b200.Index(idx, device=0), Oracle(idx), s3 + s12 + s20


def test_long_queries_match_oracle(long_case):
    g, o, seqs = long_case
    gr, gs, gc = g.search(seqs, g.default_params(output_seq=1))
    orr, os_, oc = o.search(seqs, o.default_params(output_seq=1), threads=8)
    assert len(orr) > 10
    _rows_equal(gr, orr, gs, os_, gc, oc)
    ga, oa = g.anchors(seqs), o.anchors(seqs)
    assert ga.tobytes() == oa.tobytes()
    # the 3-kb reads alone: tables of ~5,900 rows still take the fused capture kernel (shared-memory table + owner array)
    s3 = seqs[:6]
    gr, gs, gc = g.search(s3, g.default_params(output_seq=1))
    orr, os_, oc = o.search(s3, o.default_params(output_seq=1), threads=8)
    _rows_equal(gr, orr, gs, os_, gc, oc)


@pytest.fixture(scope="module")
def c3_case(workdir):
    """BASELINE.json configs[2] in miniature: 5-kb plasmid-scale queries (k-mer tables of ~10,000 rows: fused capture kernel with 1,024-thread
    CTAs, one pseudo-alignment CTA per SM, windows of ~7 kb with thousands of anchors) against a 400-genome index."""
    from conftest import make_index, make_queries
    from oracle_binding import Oracle, read_fasta
    import lexicmap_b200
    idx = make_index(workdir, "c3mini", "20,20,200000,31,20", chunks=16)
    _, seqs = read_fasta(make_queries(workdir, idx, "c3mini_q", 160, 5000, seed=41))
    return lexicmap_b200.Index(idx, device=0), Oracle(idx), seqs


def test_5kb_queries_match_oracle(c3_case):
    g, o, seqs = c3_case
    orr, os_, oc = o.search(seqs, o.default_params(output_seq=1), threads=os.cpu_count() or 8)
    assert len(orr) > 2000
    for lanes in (1, 2):
        gr, gs, gc = g.search(seqs, g.default_params(output_seq=1, lanes=lanes))
        _rows_equal(gr, orr, gs, os_, gc, oc)
    again = g.search(seqs, g.default_params(output_seq=1, lanes=2))
    _rows_equal(again[0], orr, again[1], os_, again[2], oc)


def test_genome_sharded_image_merges_to_the_unsharded_result(small_index, gpu_small, small_queries):
    """SURVEY.md §8e option 2: the image split by genome over 2 and 3 shards; every shard searches the whole batch, the host merge
    (hits summed, genomes re-ordered) reproduces the unsharded rows exactly."""
    import lexicmap_b200
    from lexicmap_b200.dist import merge_genome_shards
    ids, seqs = small_queries
    ref = gpu_small.search(seqs, gpu_small.default_params(output_seq=1))
    for n in (2, 3):
        parts = []
        for sh in range(n):
            g = lexicmap_b200.Index(small_index, device=0, shard=sh, n_shards=n)
            parts.append(g.search(seqs, g.default_params(output_seq=1)))
            g.close()
        assert sum(len(p[0]) for p in parts) == len(ref[0])
        mr, ms, mc = merge_genome_shards(parts)
        _rows_equal(mr, ref[0], ms, ref[1], mc, ref[2])
