"""CPU checks of the measurement host logic in bench.py: the SURVEY.md §8(d) byte model of the seed lookup and the usable-core count
(the benchmark's `roofline.achieved` and `cpu_baseline.cores` come from these two functions). bench.py adjusts thread-count variables
for its own process when imported, so it is imported in a child process."""
import subprocess
import sys

from conftest import ROOT


def _in_child(code):
    r = subprocess.run([sys.executable, "-c", "import sys; sys.path.insert(0, %r); import bench as b\n%s" % (ROOT, code)], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]


def test_probe_byte_model_reproduces_the_survey_figures():
    _in_child("""
# SURVEY.md 8(d): C5 nominal, n_a ~ 122 -> 7 binary-search sectors, one matched entry, one anchor written: 12 + 32 + 224 + 32 + 16 = 316 B per probe
assert b.probe_model_bytes(1, 7, 1, 1) == 316
assert b.probe_model_bytes(2 * 10**7, 7 * 2 * 10**7, 2 * 10**7, 2 * 10**7) == 316 * 2 * 10**7   # "6.3 GB for 2x10^7 probes"
assert b.probe_model_bytes(1, 1, 0, 0) == 12 + 32 + 32            # a probe that finds nothing still pays its record, the anchor sector and the search
assert b.probe_model_bytes(1, 2, 2, 5) == 12 + 32 + 64 + 64 + 80   # three matched 16-byte entries span two sectors; their five values are five anchors written
""")


def test_usable_cpus_respects_affinity_and_quota():
    _in_child("""
import os
n, visible, quota = b.usable_cpus()
assert 1 <= n <= visible and n <= len(os.sched_getaffinity(0))
if quota:
    assert n <= max(1, int(quota + 0.5))
""")
