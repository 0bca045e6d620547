#!/usr/bin/env python
"""Calibrate the CPU port (oracle/) against the reference's own published throughput: the reference reports 8,859.9 queries/min = 2.17 Mbp/s
on 16 threads for demo/q.long-reads.fasta.gz vs the 15 demo genomes with `--min-qcov-per-hsp 70 --top-n-genomes 5 --top-n-chains 1`
(/root/reference/demo/README.md:365-398). Same reads, same genomes (index built by this repo's writer with --fill-deserts), same flags, here.
Needs /root/reference (build container only). usage: calibrate_cpu_port.py [n_reads] [threads]"""
import gzip, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_binding import Oracle
from lexicmap_b200 import build
demo = "/root/reference/demo"
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 600
threads = int(sys.argv[2]) if len(sys.argv) > 2 else (os.cpu_count() or 8)
work = "/tmp/lmg_calib"; os.makedirs(work, exist_ok=True)
idx = os.path.join(work, "demo_desert.lmi")
if not os.path.exists(os.path.join(idx, "info.toml")):
    lst = os.path.join(work, "refs.list")
    open(lst, "w").write("\n".join(os.path.join(demo, "refs", f) for f in sorted(os.listdir(os.path.join(demo, "refs")))) + "\n")
    subprocess.check_call([build.build_tools(), "index", "--in-list", lst, "--out", idx, "--fill-deserts"])
seqs, cur = [], []
with gzip.open(os.path.join(demo, "q.long-reads.fasta.gz"), "rt") as f:
    for line in f:
        if line.startswith(">"):
            if cur:
                seqs.append("".join(cur)); cur = []
            if len(seqs) >= n_reads:
                break
        else:
            cur.append(line.strip())
o = Oracle(idx)
tot = sum(len(s) for s in seqs)
t = time.time()
rows, _, _ = o.search(seqs, o.default_params(top_n_genomes=5, top_n_chains=1, min_qcov_hsp=70.0), threads=threads)
dt = time.time() - t
print("%d reads, %d bp, %d threads (%d cores): %.2f s -> %.3f Mbp/s, %d rows   [reference, its own log: 2.17 Mbp/s on 16 threads]" % (len(seqs), tot, threads, os.cpu_count() or 0, dt, tot / dt / 1e6, len(rows)))
