"""CPU checks of the measurement host logic in bench.py: the SURVEY.md §8(d) byte model of the seed lookup and the usable-core count
(the benchmark's `roofline.achieved` and `cpu_baseline.cores` come from these two functions)."""
import os
import sys

from conftest import ROOT

sys.path.insert(0, ROOT)


def _bench():
    saved = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "NCCL_DEBUG")}
    try:
        import bench
    finally:   # importing bench.py adjusts thread-count variables for its own process; the test process keeps its own
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return bench


def test_probe_byte_model_reproduces_the_survey_figures():
    b = _bench()
    # C5 nominal, n_a ~ 122 -> 7 binary-search sectors, one matched entry, one anchor written: 12 + 32 + 224 + 32 + 16 = 316 B per probe
    assert b.probe_model_bytes(1, 7, 1, 1) == 316
    assert b.probe_model_bytes(2 * 10**7, 7 * 2 * 10**7, 2 * 10**7, 2 * 10**7) == 316 * 2 * 10**7   # "6.3 GB for 2x10^7 probes"
    # a probe that finds nothing still pays its record, the anchor sector and the search
    assert b.probe_model_bytes(1, 1, 0, 0) == 12 + 32 + 32
    # three matched 16-byte entries span two sectors; their five values are five anchors written
    assert b.probe_model_bytes(1, 2, 2, 5) == 12 + 32 + 64 + 64 + 80


def test_usable_cpus_respects_affinity_and_quota():
    b = _bench()
    n, visible, quota = b.usable_cpus()
    assert 1 <= n <= visible
    assert n <= len(os.sched_getaffinity(0))
    if quota:
        assert n <= int(quota + 0.5) or n == 1
