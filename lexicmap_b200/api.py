"""ctypes binding of the C ABI (include/lexicmap_gpu.h) + the host-side mirror of the reference's search interface."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liblexicmap_gpu.so")
TSV_HEADER = "query\tqlen\thits\tsgenome\tsseqid\tqcovGnm\tcls\thsp\tqcovHSP\talenHSP\tpident\tgaps\tqstart\tqend\tsstart\tsend\tsstr\tslen\tevalue\tbitscore"  # search.go:426


class Params(C.Structure):  # lmg_params
    _fields_ = [("min_prefix", C.c_int32), ("min_single_prefix", C.c_int32), ("top_n_genomes", C.c_int32), ("top_n_chains", C.c_int32),
                ("max_gap", C.c_float), ("max_distance", C.c_float), ("ext_len", C.c_int32), ("ext_len2", C.c_int32),
                ("min_qcov_genome", C.c_double), ("max_evalue", C.c_double),
                ("align_max_gap", C.c_int32), ("align_min_len", C.c_int32), ("align_band", C.c_int32), ("output_seq", C.c_int32),
                ("min_pident", C.c_double), ("min_qcov_hsp", C.c_double), ("wfa_adaptive", C.c_int32), ("lanes", C.c_int32)]


class Info(C.Structure):  # lmg_info
    _fields_ = [("k", C.c_int32), ("masks", C.c_int32), ("chunks", C.c_int32), ("partitions", C.c_int32), ("genomes", C.c_int32), ("genome_batches", C.c_int32),
                ("contig_interval", C.c_int32), ("mask_prefix", C.c_int32), ("anchor_prefix", C.c_int32), ("input_bases", C.c_int64),
                ("seed_keys", C.c_uint64), ("seed_values", C.c_uint64), ("image_bytes", C.c_uint64)]


HSP_DTYPE = np.dtype([("query", "<u4"), ("hits", "<u4"), ("genome", "<u8"), ("seq_idx", "<u4"), ("n_seqs", "<u4"), ("chunk_idx", "<u4"), ("n_chunks", "<u4"),
                      ("seq_len", "<i4"), ("cls", "<i4"), ("hsp", "<i4"), ("qb", "<i4"), ("qe", "<i4"), ("tb", "<i4"), ("te", "<i4"), ("rc", "<i4"),
                      ("alen", "<i4"), ("matches", "<i4"), ("gaps", "<i4"), ("score", "<i4"), ("bitscore", "<i4"), ("pad0", "<i4"),
                      ("evalue", "<f8"), ("qcov_hsp", "<f8"), ("pident", "<f8"), ("qcov_gnm", "<f8"), ("cigar_off", "<u8"), ("cigar_len", "<u4"), ("pad", "<u4")])
ANCHOR_DTYPE = np.dtype([("genome", "<u8"), ("query", "<u4"), ("qbegin", "<i4"), ("tbegin", "<i4"), ("len", "u1"), ("qrc", "u1"), ("trc", "u1"), ("pad", "u1")])
CHAIN_DTYPE = np.dtype([("genome", "<u8"), ("query", "<u4"), ("score", "<f4"), ("n_seeds", "<i4"), ("q0", "<i4"), ("t0", "<i4"), ("len0", "<i4"),
                        ("q1", "<i4"), ("t1", "<i4"), ("len1", "<i4"), ("rc", "<i4")])

PA_DTYPE = np.dtype([("genome", "<u8"), ("query", "<u4"), ("t_begin", "<i4"), ("t_end", "<i4"), ("rc", "<i4"), ("qb", "<i4"), ("qe", "<i4"), ("tb", "<i4"), ("te", "<i4"),
                     ("aligned_q", "<i4"), ("aligned_t", "<i4"), ("matched", "<i4"), ("n_anchors", "<i4")])


def split_align_text(rows, pool):
    """with output_seq the pool entry of a row is cigar | qseq | sseq | align (alen bytes each); returns [(qseq, sseq, align)] or None"""
    if not len(rows) or len(pool) < int(rows["cigar_off"][-1]) + int(rows["cigar_len"][-1]) + 3 * int(rows["alen"][-1]) or not int(rows["cigar_len"].max()):
        return None
    out = []
    for o, l, a in zip(rows["cigar_off"], rows["cigar_len"], rows["alen"]):
        b = int(o) + int(l)
        a = int(a)
        out.append((pool[b:b + a].decode(), pool[b + a:b + 2 * a].decode(), pool[b + 2 * a:b + 3 * a].decode()))
    return out


_lib = None


def load_library():
    """Load liblexicmap_gpu.so. Fails loudly if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("liblexicmap_gpu.so is missing: run `python -m lexicmap_b200.build` (needs nvcc); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, u64p = C.c_void_p, C.POINTER(C.c_uint64)
    L.lmg_default_params.argtypes = [C.POINTER(Params)]
    L.lmg_last_error.restype = C.c_char_p
    L.lmg_index_open.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.lmg_index_info.argtypes = [vp, C.POINTER(Info)]
    L.lmg_genome_name.argtypes = [vp, C.c_uint64, C.POINTER(C.c_char_p)]
    L.lmg_index_close.argtypes = [vp]
    L.lmg_search_batch.argtypes = [vp, C.POINTER(Params), vp, vp, C.c_int32, C.POINTER(vp)]
    L.lmg_results_rows.argtypes = [vp, C.POINTER(vp), u64p, C.POINTER(vp), u64p]
    L.lmg_results_seq_id.argtypes = [vp, C.c_uint64, C.POINTER(C.c_char_p)]
    L.lmg_results_free.argtypes = [vp]
    L.lmg_last_timing.argtypes = [vp, vp, vp]
    L.lmg_queries_upload.argtypes = [vp, vp, vp, C.c_int32, C.POINTER(vp)]
    L.lmg_search_staged.argtypes = [vp, C.POINTER(Params), vp, C.POINTER(vp)]
    L.lmg_queries_free.argtypes = [vp, vp]
    L.lmg_mask_batch.argtypes = [vp, vp, vp, C.c_int32, vp, vp, vp, vp, C.c_uint64, u64p]
    L.lmg_anchor_batch.argtypes = [vp, C.POINTER(Params), vp, vp, C.c_int32, C.POINTER(vp), u64p]
    L.lmg_chain_batch.argtypes = [vp, C.POINTER(Params), vp, vp, C.c_int32, C.POINTER(vp), u64p]
    L.lmg_pseudoalign_batch.argtypes = [vp, C.POINTER(Params), vp, vp, C.c_int32, C.POINTER(vp), u64p]
    L.lmg_wfa_batch.argtypes = [C.c_int, vp, vp, C.c_int32, C.c_int32, C.POINTER(vp), u64p]
    L.lmg_free.argtypes = [vp]
    L.lmg_index_set_total_bases.argtypes = [vp, C.c_int64]
    L.lmg_index_load_times.argtypes = [vp, vp]
    L.lmg_probe_model.argtypes = [vp, vp]
    L.lmg_index_synth.argtypes = [C.c_int, C.c_int32, C.c_uint64, C.c_uint64, C.c_int32, C.c_int32, C.c_int32, C.POINTER(vp)]
    L.lmg_gather_bench.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_int32, C.c_int32, vp]
    L.lmg_probe_bench.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_int32, C.c_int32, vp]
    _lib = L
    return L


def pack_queries(seqs):
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    off = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs])
    buf = np.frombuffer(b"".join(bs) + b"\0" * 16, dtype=np.uint8).copy()
    return buf, off


class Index:
    """GPU-resident LexicMap index (NewIndexSearcher, lib-index-search.go:237)."""

    def __init__(self, lmi_dir, device=0, shard=0, n_shards=1):
        self.lib = load_library()
        h = C.c_void_p()
        if self.lib.lmg_index_open(os.fsencode(lmi_dir), device, shard, n_shards, C.byref(h)) != 0:
            raise RuntimeError(self.lib.lmg_last_error().decode())
        self.h = h
        self.info = Info()
        self.lib.lmg_index_info(self.h, C.byref(self.info))

    @classmethod
    def synthetic(cls, masks=20000, per_mask=100000, seed=1, mask_lo=0, mask_hi=None, device=0, with_values=False):
        """seeds-only synthetic image for the seed-lookup microbenchmark (BASELINE.json configs[4]); only probe_bench() runs on it"""
        self = cls.__new__(cls)
        self.lib = load_library()
        h = C.c_void_p()
        if self.lib.lmg_index_synth(device, masks, per_mask, seed, mask_lo, masks if mask_hi is None else mask_hi, int(with_values), C.byref(h)) != 0:
            raise RuntimeError(self.lib.lmg_last_error().decode())
        self.h = h
        self.info = Info()
        self.lib.lmg_index_info(self.h, C.byref(self.info))
        return self

    def probe_bench(self, n_queries, seed=20260926, min_prefix=15, iters=5):
        out = np.zeros(16, np.float64)
        if self.lib.lmg_probe_bench(self.h, n_queries, seed, min_prefix, iters, out.ctypes.data) != 0:
            self._err()
        keys = ["issued", "survivors", "kernel_ms", "hits", "sum_log2", "sum_hit_sectors", "sum_values", "steps", "entries", "gen_ms", "kernel_ms_best", "regroup_ms"]
        return dict(zip(keys, out.tolist()))

    def set_total_bases(self, n):
        if self.lib.lmg_index_set_total_bases(self.h, int(n)) != 0:
            self._err()

    def load_times(self):
        ms = np.zeros(4, np.float64)
        self.lib.lmg_index_load_times(self.h, ms.ctypes.data)
        return dict(zip(["genomes_ms", "seed_count_ms", "seed_fill_ms", "total_ms"], ms.tolist()))

    def probe_model(self):
        s = np.zeros(4, np.uint64)
        self.lib.lmg_probe_model(self.h, s.ctypes.data)
        return [int(x) for x in s]

    def close(self):
        if getattr(self, "h", None):
            self.lib.lmg_index_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def default_params(self, **kw):
        p = Params()
        self.lib.lmg_default_params(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    def _err(self):
        raise RuntimeError(self.lib.lmg_last_error().decode())

    def genome_name(self, g):
        s = C.c_char_p()
        self.lib.lmg_genome_name(self.h, int(g), C.byref(s))
        return (s.value or b"").decode()

    # ---- Index.Search, batched (lib-index-search.go:1191)
    def search(self, seqs, params=None, packed=None, rows_only=False):
        """returns (rows: HSP_DTYPE array, seqids, cigars). `packed` = (uint8 buf, uint64 off) to skip re-packing.
        rows_only: (rows, None, None) without the per-row Python work (benchmarks)."""
        p = params or self.default_params()
        buf, off = packed if packed is not None else pack_queries(seqs)
        n = len(off) - 1
        r = C.c_void_p()
        if self.lib.lmg_search_batch(self.h, C.byref(p), buf.ctypes.data, off.ctypes.data, n, C.byref(r)) != 0:
            self._err()
        return self._collect(r, rows_only)

    def stage(self, seqs=None, packed=None):
        """copy a query batch to HBM ahead of time; returns an opaque handle for search_staged()"""
        buf, off = packed if packed is not None else pack_queries(seqs)
        q = C.c_void_p()
        if self.lib.lmg_queries_upload(self.h, buf.ctypes.data, off.ctypes.data, len(off) - 1, C.byref(q)) != 0:
            self._err()
        return q

    def search_staged(self, q, params=None, collect=True):
        p = params or self.default_params()
        r = C.c_void_p()
        if self.lib.lmg_search_staged(self.h, C.byref(p), q, C.byref(r)) != 0:
            self._err()
        if collect:
            return self._collect(r, rows_only=(collect == "rows"))
        nr = self._nrows(r)
        self.lib.lmg_results_free(r)
        return nr

    def free_staged(self, q):
        self.lib.lmg_queries_free(self.h, q)

    def _nrows(self, r):
        rows_p, pool_p, nr, npool = C.c_void_p(), C.c_void_p(), C.c_uint64(), C.c_uint64()
        self.lib.lmg_results_rows(r, C.byref(rows_p), C.byref(nr), C.byref(pool_p), C.byref(npool))
        return nr.value

    def search_count(self, packed, params=None):
        """search from host buffers, return only the number of rows (bench e2e leg: no Python-side row copies)"""
        p = params or self.default_params()
        buf, off = packed
        r = C.c_void_p()
        if self.lib.lmg_search_batch(self.h, C.byref(p), buf.ctypes.data, off.ctypes.data, len(off) - 1, C.byref(r)) != 0:
            self._err()
        nr = self._nrows(r)
        self.lib.lmg_results_free(r)
        return nr

    def _collect(self, r, rows_only=False):
        rows_p, pool_p, nr, npool = C.c_void_p(), C.c_void_p(), C.c_uint64(), C.c_uint64()
        self.lib.lmg_results_rows(r, C.byref(rows_p), C.byref(nr), C.byref(pool_p), C.byref(npool))
        rows = np.frombuffer(C.string_at(rows_p, nr.value * HSP_DTYPE.itemsize), dtype=HSP_DTYPE).copy() if nr.value else np.zeros(0, HSP_DTYPE)
        if rows_only:
            self.lib.lmg_results_free(r)
            return rows, None, None
        pool = C.string_at(pool_p, npool.value) if npool.value else b""
        seqids = []
        s = C.c_char_p()
        for i in range(nr.value):
            self.lib.lmg_results_seq_id(r, i, C.byref(s))
            seqids.append(s.value.decode())
        cig = [pool[int(o):int(o) + int(l)].decode() for o, l in zip(rows["cigar_off"], rows["cigar_len"])]
        self.last_align_text = split_align_text(rows, pool)
        self.lib.lmg_results_free(r)
        return rows, seqids, cig

    def timing(self):
        ms = np.zeros(16, np.float64)
        cnt = np.zeros(16, np.uint64)
        self.lib.lmg_last_timing(self.h, ms.ctypes.data, cnt.ctypes.data)
        return ms, cnt

    def format_tsv(self, rows, seqids, qids, qlens, cigars=None, texts=None):
        """printResult (search.go:437-533): 20 columns; with -a also cigar, qseq, sseq, align (`texts` = self.last_align_text)."""
        out = []
        for i, r in enumerate(rows):
            q = int(r["query"])
            line = "%s\t%d\t%d\t%s\t%s\t%.3f\t%d\t%d\t%.3f\t%d\t%.3f\t%d\t%d\t%d\t%d\t%d\t%s\t%d\t%.2e\t%d" % (
                qids[q], qlens[q], r["hits"], self.genome_name(r["genome"]), seqids[i], r["qcov_gnm"], r["cls"], r["hsp"], r["qcov_hsp"], r["alen"], r["pident"], r["gaps"],
                r["qb"] + 1, r["qe"] + 1, r["tb"] + 1, r["te"] + 1, "-" if r["rc"] else "+", r["seq_len"], r["evalue"], r["bitscore"])
            if cigars is not None:
                line += "\t" + cigars[i]
                if texts is not None:
                    line += "\t%s\t%s\t%s" % texts[i]
            out.append(line)
        return out

    # ---- stage-wise entry points (parity tests)
    def mask(self, seqs):
        buf, off = pack_queries(seqs)
        n, m = len(seqs), self.info.masks
        kmers, nlocs, minloc = np.zeros(n * m, np.uint64), np.zeros(n * m, np.uint32), np.zeros(n * m, np.uint32)
        suf = np.zeros(4 * n * m, np.uint64)
        ns = C.c_uint64()
        if self.lib.lmg_mask_batch(self.h, buf.ctypes.data, off.ctypes.data, n, kmers.ctypes.data, nlocs.ctypes.data, minloc.ctypes.data, suf.ctypes.data, n * m, C.byref(ns)) != 0:
            self._err()
        return kmers, nlocs, minloc, suf[:4 * ns.value].reshape(-1, 4)

    def _stage(self, fn, dtype, seqs, params):
        p = params or self.default_params()
        buf, off = pack_queries(seqs)
        ptr, n = C.c_void_p(), C.c_uint64()
        if fn(self.h, C.byref(p), buf.ctypes.data, off.ctypes.data, len(seqs), C.byref(ptr), C.byref(n)) != 0:
            self._err()
        a = np.frombuffer(C.string_at(ptr, n.value * dtype.itemsize), dtype=dtype).copy() if n.value else np.zeros(0, dtype)
        self.lib.lmg_free(ptr)
        return a

    def anchors(self, seqs, params=None):
        return self._stage(self.lib.lmg_anchor_batch, ANCHOR_DTYPE, seqs, params)

    def chains(self, seqs, params=None):
        return self._stage(self.lib.lmg_chain_batch, CHAIN_DTYPE, seqs, params)

    def pseudoalign(self, seqs, params=None):
        """window geometry + pseudo-alignment + Chainer2: one record per Chain2Result, window coordinates (a9-a12)"""
        return self._stage(self.lib.lmg_pseudoalign_batch, PA_DTYPE, seqs, params)


def gather_bench(device=0, gbytes=8.0, n_threads=1 << 24, per_thread=8, iters=5):
    """random 32-byte-sector read rate of the device: {"sectors_per_s", "gbs_at_32B", "best_ms", "mean_ms"}"""
    L = load_library()
    out = np.zeros(4, np.float64)
    if L.lmg_gather_bench(device, int(gbytes * (1 << 30)), n_threads, per_thread, iters, out.ctypes.data) != 0:
        raise RuntimeError(L.lmg_last_error().decode())
    return dict(zip(["sectors_per_s", "gbs_at_32B", "best_ms", "mean_ms"], out.tolist()))


def wfa_batch(pairs, device=0, adaptive=1):
    L = load_library()
    flat = []
    for q, t in pairs:
        flat += [q, t]
    buf, off = pack_queries(flat)
    ptr, n = C.c_void_p(), C.c_uint64()
    if L.lmg_wfa_batch(device, buf.ctypes.data, off.ctypes.data, len(pairs), adaptive, C.byref(ptr), C.byref(n)) != 0:
        raise RuntimeError(L.lmg_last_error().decode())
    s = C.string_at(ptr, n.value).decode()
    L.lmg_free(ptr)
    return s.split("\n")[:-1]
