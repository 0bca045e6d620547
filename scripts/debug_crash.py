import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("LMG_BENCH_FAMILIES", "10"); os.environ.setdefault("LMG_BENCH_MEMBERS", "10"); os.environ.setdefault("LMG_BENCH_NQ", "2000")
import bench, lexicmap_b200
from lexicmap_b200.api import pack_queries
from oracle_binding import read_fasta
idx_dir, qf = bench.ensure_workload(0)
ids, seqs = read_fasta(qf)
idx = lexicmap_b200.Index(idx_dir)
packed = pack_queries(seqs); prm = idx.default_params()
mode = sys.argv[1] if len(sys.argv) > 1 else "e2e"
if mode == "e2e":
    print("rows", idx.search_count(packed, prm)); print("rows", idx.search_count(packed, prm))
else:
    st = idx.stage(packed=packed); print("rows", idx.search_staged(st, prm, collect=False)); print("rows", idx.search_staged(st, prm, collect=False))
print(idx.timing()[0][:8])
