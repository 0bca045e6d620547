"""GPU parity, stage by stage: CUDA path (through the C ABI) vs the CPU oracle on the same seeded inputs. Bit-exact."""
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
