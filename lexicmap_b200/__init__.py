"""lexicmap_b200 — B200-native (sm_100a CUDA) implementation of LexicMap's query-side search path.

Public host API mirrors the reference's `Index.Search` seam (lexicmap/cmd/lib-index-search.go:237, :1191):

    idx = lexicmap_b200.Index("db.lmi", device=0)          # NewIndexSearcher
    rows = idx.search(["ACGT...", ...])                    # Index.Search, batched
    for line in idx.format_tsv(rows, query_ids): ...        # printResult (search.go:437-533)

The compute path is the CUDA library liblexicmap_gpu.so (C ABI: include/lexicmap_gpu.h). There is no CPU fallback:
importing works without a GPU, but opening an index raises if the library or a CUDA device is missing.
"""
from .api import Index, Params, HSP_DTYPE, ANCHOR_DTYPE, CHAIN_DTYPE, load_library, TSV_HEADER  # noqa: F401
