import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, lexicmap_b200
from lexicmap_b200.api import pack_queries
from oracle_binding import read_fasta
idx_dir, qf = bench.ensure_workload(0)
ids, seqs = read_fasta(qf)
idx = lexicmap_b200.Index(idx_dir)
packed = pack_queries(seqs); prm = idx.default_params(); st = idx.stage(packed=packed)
for i in range(3): idx.search_staged(st, prm, collect=False)
os.environ["X"]="1"
print("---- timed call", file=sys.stderr)
idx.search_staged(st, prm, collect=False)
ms, cnt = idx.timing(); print([round(x,1) for x in ms])
