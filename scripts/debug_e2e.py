"""manual experiment: why is the host-buffer (e2e) call slower than the staged call?"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("LMG_BENCH_FAMILIES", "10"); os.environ.setdefault("LMG_BENCH_MEMBERS", "10"); os.environ.setdefault("LMG_BENCH_NQ", "2000")
import bench, torch, lexicmap_b200
from lexicmap_b200.api import pack_queries
from oracle_binding import read_fasta
idx_dir, qf = bench.ensure_workload(0)
ids, seqs = read_fasta(qf)
idx = lexicmap_b200.Index(idx_dir)
packed = pack_queries(seqs); prm = idx.default_params(); st = idx.stage(packed=packed)
def run(name, fn, n=6, sync=False):
    out = []
    for i in range(n):
        t = time.perf_counter(); fn(); 
        if sync: torch.cuda.synchronize()
        w = (time.perf_counter() - t) * 1e3; ms, _ = idx.timing(); out.append((round(w, 1), [round(x, 1) for x in ms[:8]]))
    print(name); [print("   ", o) for o in out]
run("staged", lambda: idx.search_staged(st, prm, collect=False))
run("staged+devsync", lambda: idx.search_staged(st, prm, collect=False), sync=True)
run("e2e", lambda: idx.search_count(packed, prm))
run("e2e+devsync", lambda: idx.search_count(packed, prm), sync=True)
run("staged again", lambda: idx.search_staged(st, prm, collect=False))
