"""Builds tests/golden/demo_long_reads_sample.fasta.gz and demo_long_reads_readme_rows.tsv from the reference's demo data
(/root/reference/demo/q.long-reads.fasta.gz, /root/reference/demo/README.md:410-419). Run in the build container only; the fixtures are committed.

Sample (deterministic): every read named in the README's result overview (their rows are the reference's own output: golden), the 24 longest
reads (52-90 kb), 36 reads of 32-50 kb (beyond the fast WFA kernel's 32,000-base limit), and 170 reads drawn with numpy's PCG64(20260924)."""
import gzip
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_binding import read_fasta

DEMO = "/root/reference/demo"
ids, seqs = read_fasta(os.path.join(DEMO, "q.long-reads.fasta.gz"))
rows = []
for line in open(os.path.join(DEMO, "README.md")).read().splitlines()[407:420]:
    f = line.split()
    if len(f) >= 20 and re.match(r"GCF_\d+\.\d_r\d+$", f[0]):
        rows.append(f[:20])
assert len(rows) == 10, len(rows)
L = np.array([len(s) for s in seqs])
order = np.argsort(-L, kind="stable")
pick = [ids.index(r[0]) for r in rows]
pick += order[:24].tolist()
mid = [i for i in order if 32000 < L[i] <= 50000]
pick += mid[:: max(1, len(mid) // 36)][:36]
rng = np.random.Generator(np.random.PCG64(20260924))
pick += rng.choice(len(seqs), size=170, replace=False).tolist()
seen, out = set(), []
for i in pick:
    if i not in seen:
        seen.add(i)
        out.append(i)
with gzip.GzipFile(os.path.join(ROOT, "tests", "golden", "demo_long_reads_sample.fasta.gz"), "wb", mtime=0) as f:
    for i in out:
        f.write((">%s\n%s\n" % (ids[i], seqs[i])).encode())
hdr = "query qlen hits sgenome sseqid qcovGnm cls hsp qcovHSP alenHSP pident gaps qstart qend sstart send sstr slen evalue bitscore".split()
with open(os.path.join(ROOT, "tests", "golden", "demo_long_reads_readme_rows.tsv"), "w") as f:
    f.write("\t".join(hdr) + "\n")
    for r in rows:
        f.write("\t".join(r) + "\n")
print(len(out), "reads,", int(L[out].sum()), "bases; >32kb:", int((L[out] > 32000).sum()))
