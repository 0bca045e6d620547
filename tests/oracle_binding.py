"""ctypes binding of the CPU oracle (oracle/liblexicmap_oracle.so). Test infrastructure only.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "liblexicmap_oracle.so")


class Params(C.Structure):
    _fields_ = [("min_prefix", C.c_int32), ("min_single_prefix", C.c_int32), ("top_n_genomes", C.c_int32), ("top_n_chains", C.c_int32),
                ("max_gap", C.c_float), ("max_distance", C.c_float), ("ext_len", C.c_int32), ("ext_len2", C.c_int32),
                ("min_qcov_genome", C.c_double), ("max_evalue", C.c_double),
                ("align_max_gap", C.c_int32), ("align_min_len", C.c_int32), ("align_band", C.c_int32), ("output_seq", C.c_int32),
                ("min_pident", C.c_double), ("min_qcov_hsp", C.c_double), ("wfa_adaptive", C.c_int32), ("lanes", C.c_int32)]


HSP_DTYPE = np.dtype([("query", "<u4"), ("hits", "<u4"), ("genome", "<u8"), ("seq_idx", "<u4"), ("n_seqs", "<u4"), ("chunk_idx", "<u4"), ("n_chunks", "<u4"),
                      ("seq_len", "<i4"), ("cls", "<i4"), ("hsp", "<i4"), ("qb", "<i4"), ("qe", "<i4"), ("tb", "<i4"), ("te", "<i4"), ("rc", "<i4"),
                      ("alen", "<i4"), ("matches", "<i4"), ("gaps", "<i4"), ("score", "<i4"), ("bitscore", "<i4"), ("pad0", "<i4"),
                      ("evalue", "<f8"), ("qcov_hsp", "<f8"), ("pident", "<f8"), ("qcov_gnm", "<f8"), ("cigar_off", "<u8"), ("cigar_len", "<u4"), ("pad", "<u4")])
ANCHOR_DTYPE = np.dtype([("genome", "<u8"), ("query", "<u4"), ("qbegin", "<i4"), ("tbegin", "<i4"), ("len", "u1"), ("qrc", "u1"), ("trc", "u1"), ("pad", "u1")])
CHAIN_DTYPE = np.dtype([("genome", "<u8"), ("query", "<u4"), ("score", "<f4"), ("n_seeds", "<i4"), ("q0", "<i4"), ("t0", "<i4"), ("len0", "<i4"),
                        ("q1", "<i4"), ("t1", "<i4"), ("len1", "<i4"), ("rc", "<i4")])
PA_DTYPE = np.dtype([("genome", "<u8"), ("query", "<u4"), ("t_begin", "<i4"), ("t_end", "<i4"), ("rc", "<i4"), ("qb", "<i4"), ("qe", "<i4"), ("tb", "<i4"), ("te", "<i4"),
                     ("aligned_q", "<i4"), ("aligned_t", "<i4"), ("matched", "<i4"), ("n_anchors", "<i4")])
assert PA_DTYPE.itemsize == 56
assert HSP_DTYPE.itemsize == 136 and ANCHOR_DTYPE.itemsize == 24 and CHAIN_DTYPE.itemsize == 48


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


def pack_queries(seqs):
    """list of str/bytes -> (uint8 buffer, uint64 offsets[n+1])"""
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    off = np.zeros(len(bs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(b) for b in bs])
    buf = np.frombuffer(b"".join(bs) + b"\0", dtype=np.uint8).copy()
    return buf, off


def read_fasta(path):
    ids, seqs = [], []
    import gzip
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as f:
        for line in f:
            line = line.rstrip("\n")
            if line.startswith(">"):
                ids.append(line[1:].split()[0])
                seqs.append([])
            elif seqs:
                seqs[-1].append(line.strip())
    return ids, ["".join(s) for s in seqs]


class Oracle:
    def __init__(self, lmi_dir):
        if not os.path.exists(LIB):
            build()
        self.lib = L = C.CDLL(LIB)
        L.lmo_open.restype = C.c_void_p
        L.lmo_open.argtypes = [C.c_char_p]
        L.lmo_last_error.restype = C.c_char_p
        L.lmo_close.argtypes = [C.c_void_p]
        L.lmo_genome_name.restype = C.c_char_p
        L.lmo_genome_name.argtypes = [C.c_void_p, C.c_uint64]
        L.lmo_total_bases.restype = C.c_int64
        L.lmo_total_bases.argtypes = [C.c_void_p]
        L.lmo_search_batch.restype = C.c_void_p
        L.lmo_search_batch.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_int32, C.c_int]
        L.lmo_rows.restype = C.c_uint64
        L.lmo_rows.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        L.lmo_row_seqid.restype = C.c_char_p
        L.lmo_row_seqid.argtypes = [C.c_void_p, C.c_uint64]
        L.lmo_rows_free.argtypes = [C.c_void_p]
        L.lmo_mask_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_int]
        for f in (L.lmo_anchor_batch, L.lmo_chain_batch, L.lmo_pseudoalign_batch):
            f.restype = C.c_void_p
            f.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_uint64)]
        L.lmo_wfa_batch.restype = C.c_void_p
        L.lmo_wfa_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_uint64)]
        L.lmo_kv_search.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.lmo_open_kv.restype = C.c_void_p
        L.lmo_open_kv.argtypes = [C.c_char_p]
        L.lmo_tree_search.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.lmo_subseq.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_char_p, C.c_int]
        L.lmo_dust.argtypes = [C.c_uint64, C.c_int]
        L.lmo_free.argtypes = [C.c_void_p]
        self.h = L.lmo_open(lmi_dir.encode()) if lmi_dir else None
        if lmi_dir and not self.h:
            raise RuntimeError(L.lmo_last_error().decode())

    def close(self):
        if self.h:
            self.lib.lmo_close(self.h)
            self.h = None

    def default_params(self, **kw):
        p = Params()
        self.lib.lmo_default_params(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    def search(self, seqs, params=None, threads=1):
        """returns (rows structured array, seqids list, cigars list)"""
        p = params or self.default_params()
        buf, off = pack_queries(seqs)
        r = self.lib.lmo_search_batch(self.h, C.byref(p), buf.ctypes.data, off.ctypes.data, len(seqs), threads)
        if not r:
            raise RuntimeError(self.lib.lmo_last_error().decode())
        rows_p, pool_p = C.c_void_p(), C.c_void_p()
        n = self.lib.lmo_rows(r, C.byref(rows_p), C.byref(pool_p))
        rows = np.frombuffer(C.string_at(rows_p, n * HSP_DTYPE.itemsize), dtype=HSP_DTYPE).copy() if n else np.zeros(0, HSP_DTYPE)
        seqids = [self.lib.lmo_row_seqid(r, i).decode() for i in range(n)]
        cig = []
        if n:
            has_text = int(rows["cigar_len"].max()) > 0
            end = int(rows["cigar_off"][-1] + rows["cigar_len"][-1]) + (3 * int(rows["alen"][-1]) if has_text else 0)
            pool = C.string_at(pool_p, end) if end else b""
            cig = [pool[int(o):int(o) + int(l)].decode() for o, l in zip(rows["cigar_off"], rows["cigar_len"])]
            self.last_align_text = []
            for o, l, a in zip(rows["cigar_off"], rows["cigar_len"], rows["alen"]):
                b, a = int(o) + int(l), int(a)
                self.last_align_text.append((pool[b:b + a].decode(), pool[b + a:b + 2 * a].decode(), pool[b + 2 * a:b + 3 * a].decode()) if has_text else ("", "", ""))
        self.lib.lmo_rows_free(r)
        return rows, seqids, cig

    def genome_name(self, g):
        return self.lib.lmo_genome_name(self.h, int(g)).decode()

    def mask(self, seqs, n_masks, bruteforce=False, suf_cap=None):
        buf, off = pack_queries(seqs)
        n = len(seqs)
        kmers = np.zeros(n * n_masks, np.uint64)
        nlocs = np.zeros(n * n_masks, np.uint32)
        minloc = np.zeros(n * n_masks, np.uint32)
        cap = suf_cap or n * n_masks
        suf = np.zeros(4 * cap, np.uint64)
        ns = C.c_uint64()
        self.lib.lmo_mask_batch(self.h, buf.ctypes.data, off.ctypes.data, n, kmers.ctypes.data, nlocs.ctypes.data, minloc.ctypes.data, suf.ctypes.data, cap, C.byref(ns), int(bruteforce))
        return kmers, nlocs, minloc, suf[:4 * ns.value].reshape(-1, 4)

    def _stage(self, fn, dtype, seqs, params):
        p = params or self.default_params()
        buf, off = pack_queries(seqs)
        n = C.c_uint64()
        ptr = fn(self.h, C.byref(p), buf.ctypes.data, off.ctypes.data, len(seqs), C.byref(n))
        if not ptr:
            raise RuntimeError(self.lib.lmo_last_error().decode())
        a = np.frombuffer(C.string_at(ptr, n.value * dtype.itemsize), dtype=dtype).copy()
        self.lib.lmo_free(ptr)
        return a

    def anchors(self, seqs, params=None):
        return self._stage(self.lib.lmo_anchor_batch, ANCHOR_DTYPE, seqs, params)

    def chains(self, seqs, params=None):
        return self._stage(self.lib.lmo_chain_batch, CHAIN_DTYPE, seqs, params)

    def pseudoalign(self, seqs, params=None):
        return self._stage(self.lib.lmo_pseudoalign_batch, PA_DTYPE, seqs, params)

    def wfa(self, pairs, adaptive=1):
        flat = []
        for q, t in pairs:
            flat += [q, t]
        buf, off = pack_queries(flat)
        n = C.c_uint64()
        ptr = self.lib.lmo_wfa_batch(buf.ctypes.data, off.ctypes.data, len(pairs), adaptive, C.byref(n))
        s = C.string_at(ptr, n.value).decode()
        self.lib.lmo_free(ptr)
        return s.split("\n")[:-1]

    def open_kv(self, file):
        self.h = self.lib.lmo_open_kv(file.encode())
        if not self.h:
            raise RuntimeError(self.lib.lmo_last_error().decode())

    def tree_search(self, keys, k, key, p):
        a = np.asarray(keys, dtype=np.uint64)
        lo, hi = C.c_int(), C.c_int()
        ok = self.lib.lmo_tree_search(a.ctypes.data, len(a), k, int(key), p, C.byref(lo), C.byref(hi))
        return bool(ok), lo.value, hi.value

    def subseq(self, bgi, start, end):
        buf = C.create_string_buffer(end - start + 2)
        n = self.lib.lmo_subseq(self.h, int(bgi), start, end, buf, end - start + 1)
        return buf.raw[:n].decode()

    def kv_search(self, mask, kmer, p, reversed_=False, check_flag=True, cap=4096):
        lens = np.zeros(cap, np.uint8)
        vals = np.zeros(cap, np.uint64)
        n = self.lib.lmo_kv_search(self.h, mask, int(kmer), p, int(reversed_), int(check_flag), lens.ctypes.data, vals.ctypes.data, cap)
        return n, lens[:min(n, cap)], vals[:min(n, cap)]


def format_tsv(rows, seqids, qids, qlens, genome_name, cigars=None, texts=None):
    """Reference TSV rows (search.go:426-518)."""
    out = []
    for i, r in enumerate(rows):
        q = int(r["query"])
        line = "%s\t%d\t%d\t%s\t%s\t%.3f\t%d\t%d\t%.3f\t%d\t%.3f\t%d\t%d\t%d\t%d\t%d\t%s\t%d\t%.2e\t%d" % (
            qids[q], qlens[q], r["hits"], genome_name(r["genome"]), seqids[i], r["qcov_gnm"], r["cls"], r["hsp"], r["qcov_hsp"], r["alen"], r["pident"], r["gaps"],
            r["qb"] + 1, r["qe"] + 1, r["tb"] + 1, r["te"] + 1, "-" if r["rc"] else "+", r["seq_len"], r["evalue"], r["bitscore"])
        if cigars is not None:
            line += "\t" + cigars[i]
            if texts is not None:
                line += "\t%s\t%s\t%s" % texts[i]
        out.append(line)
    return out
