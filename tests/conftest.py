import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _tools():
    from lexicmap_b200 import build
    return build.build_tools()


def make_index(tmp, name, synth, masks=20000, chunks=4, extra=()):
    out = os.path.join(str(tmp), name + ".lmi")
    if not os.path.exists(os.path.join(out, "info.toml")):
        subprocess.check_call([_tools(), "index", "--synth", synth, "--out", out, "--masks", str(masks), "--chunks", str(chunks), *extra], stderr=subprocess.DEVNULL)
    return out


def make_queries(tmp, index, name, n, length, seed=20260925, max_sub=0.10, max_indel=0.01):
    out = os.path.join(str(tmp), name + ".fa")
    subprocess.check_call([_tools(), "synth-queries", "--index", index, "--n", str(n), "--len", str(length), "--seed", str(seed), "--out", out,
                           "--max-sub", str(max_sub), "--max-indel", str(max_indel)])
    return out


@pytest.fixture(scope="session")
def workdir(tmp_path_factory):
    return tmp_path_factory.mktemp("lmi")


@pytest.fixture(scope="session")
def small_index(workdir):
    """4 families x 4 members x 40 kb, multi-contig; 20,000 masks (the reference default)."""
    return make_index(workdir, "small", "4,4,40000,42,3")


@pytest.fixture(scope="session")
def small_queries(workdir, small_index):
    from oracle_binding import read_fasta
    ids, seqs = read_fasta(make_queries(workdir, small_index, "small_q", 24, 800))
    # edge cases: shorter than k, exactly k, poly-A, lower case, with N
    seqs += ["ACGTACGTAC", seqs[0][:31], "A" * 200, seqs[1].lower(), seqs[2][:300] + "NNNNNNNNNN" + seqs[2][310:]]
    ids += ["short", "exactk", "polyA", "lower", "withN"]
    return ids, seqs


@pytest.fixture(scope="session")
def oracle_small(small_index):
    from oracle_binding import Oracle
    return Oracle(small_index)


@pytest.fixture(scope="session")
def gpu_small(small_index):
    import lexicmap_b200
    return lexicmap_b200.Index(small_index, device=0)
