/* lexicmap_gpu.h — C ABI of liblexicmap_gpu.so: the drop-in boundary for LexicMap's query-side search path.
 *
 * The reference (shenwei356/LexicMap v0.10.0) is a statically linked pure-Go binary (lexicmap/build.sh:7,
 * CGO_ENABLED=0) with no plugin/FFI seam. The narrowest seam containing the whole named path is
 *
 *     func (idx *Index) Search(query *Query, genomeIds *map[uint64]*[]uint64, debug bool) (*[]*SearchResult, error)
 *         lexicmap/cmd/lib-index-search.go:1191        (caller: lexicmap/cmd/search.go:589-602)
 *     func NewIndexSearcher(outDir string, opt *IndexSearchingOptions) (*Index, error)
 *         lexicmap/cmd/lib-index-search.go:237
 *
 * A GPU is batch oriented, so `Search` is replaced by a batched equivalent. INTEGRATION.md shows the cgo binding a
 * maintainer would add to search.go. All functions return 0 on success, <0 on error; lmg_last_error() gives text.
 * Plain pointers and sizes only; no C++/torch types cross this boundary.
 */
#ifndef LEXICMAP_GPU_H
#define LEXICMAP_GPU_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* IndexSearchingOptions (lib-index-search.go:57-106) + SeqComparatorOptions (lib-seq_compare.go:34-46) +
 * Chaining2Options (lib-chaining2.go:29-39). Defaults = `lexicmap search` flag defaults (search.go:631-731). */
typedef struct lmg_params {
  int32_t min_prefix;        /* -p/--seed-min-prefix 15          lib-index-search.go:68 */
  int32_t min_single_prefix; /* -P/--seed-min-single-prefix 17   :70 */
  int32_t top_n_genomes;     /* -n 0                              :72 */
  int32_t top_n_chains;      /* -N 0                              :73 */
  float   max_gap;           /* --seed-max-gap 50                 :76 */
  float   max_distance;      /* --seed-max-dist 1000              :77 */
  int32_t ext_len;           /* --align-ext-len 1000              :80 */
  int32_t ext_len2;          /* 50 (search.go:328)                :81 */
  double  min_qcov_genome;   /* -Q 0                              :84 */
  double  max_evalue;        /* -e 10                             :85 */
  int32_t align_max_gap;     /* --align-max-gap 20                search.go:366 */
  int32_t align_min_len;     /* -l/--align-min-match-len 50       search.go:369 */
  int32_t align_band;        /* --align-band 100                  search.go:374 */
  int32_t output_seq;        /* -a/--all: fill CIGAR/qseq/sseq/align */
  double  min_pident;        /* -i/--align-min-match-pident 70 */
  double  min_qcov_hsp;      /* -q/--min-qcov-per-hsp 0 */
  int32_t wfa_adaptive;      /* 1 = WFA adaptive wavefront reduction (MinWFLen 10, MaxDistDiff 50) as the reference enables at lib-index-search.go:1911; 0 = exact */
  int32_t lanes;              /* concurrent sub-batches per call (own stream + host thread each); 0 = automatic (up to 6 for large batches, else 1) */
} lmg_params;

typedef struct lmg_info {          /* IndexInfo, lib-index-build.go:1914-1932 */
  int32_t k, masks, chunks, partitions, genomes, genome_batches, contig_interval, mask_prefix, anchor_prefix;
  int64_t input_bases;
  uint64_t seed_keys, seed_values; /* entries of the GPU-resident image */
  uint64_t image_bytes;            /* HBM bytes held by the image */
} lmg_info;

/* One output row = one HSP, i.e. one line of the reference TSV (search.go:492-518). Coordinates 0-based inclusive;
 * the TSV writer adds 1. Rows of a query are contiguous and in reference order (lib-index-search.go:2745-2932). */
typedef struct lmg_hsp {
  uint32_t query;       /* index in the batch */
  uint32_t hits;        /* genomes in this query's result (TSV `hits`) */
  uint64_t genome;      /* batch<<17 | index (SearchResult.BatchGenomeIndex) */
  uint32_t seq_idx, n_seqs, chunk_idx, n_chunks;
  int32_t  seq_len;     /* slen */
  int32_t  cls, hsp;
  int32_t  qb, qe, tb, te;
  int32_t  rc;          /* sstr: 0 '+', 1 '-' */
  int32_t  alen, matches, gaps, score, bitscore, pad0;
  double   evalue, qcov_hsp, pident, qcov_gnm;
  uint64_t cigar_off;   /* into the string pool; SAM convention (I/D already swapped, lib-index-search.go:2331-2338). With output_seq the
                           pool entry of a row is cigar (cigar_len bytes) | qseq | sseq | align (alen bytes each; cigar.AlignmentText, :2342-2346) */
  uint32_t cigar_len;
  uint32_t pad;
} lmg_hsp;

/* stage records (test / profiling entry points) */
typedef struct lmg_anchor {       /* SubstrPair + genome, lib-index-search.go:805-817, :1526-1556 */
  uint64_t genome; uint32_t query; int32_t qbegin, tbegin; uint8_t len, qrc, trc, pad;
} lmg_anchor;
typedef struct lmg_chain {        /* one lexichash chain: first/last anchor (all the later stages read, :1987-2006) */
  uint64_t genome; uint32_t query; float score; int32_t n_seeds;
  int32_t q0, t0, len0, q1, t1, len1; int32_t rc;
} lmg_chain;

typedef struct lmg_pa {           /* one Chain2Result of SeqComparator.Compare (lib-seq_compare.go:335-522, lib-chaining2.go:106-135) in window coordinates */
  uint64_t genome; uint32_t query; int32_t t_begin, t_end, rc;   /* the chain's target window [t_begin, t_end] on the concatenated genome, rc = minus strand (lib-index-search.go:1987-2051) */
  int32_t qb, qe, tb, te, aligned_q, aligned_t, matched, n_anchors;
} lmg_pa;

typedef struct lmg_index lmg_index;
typedef struct lmg_results lmg_results;

void        lmg_default_params(lmg_params* p);
const char* lmg_last_error(void);
/* NewIndexSearcher (lib-index-search.go:237): parse the .lmi directory, decode all seed chunks and genomes into a
 * GPU-resident image on `device`. shard/n_shards: genome sharding for multi-GPU (keep values with genome % n == shard). */
int  lmg_index_open(const char* lmi_dir, int device, int shard, int n_shards, lmg_index** out);
int  lmg_index_info(const lmg_index* idx, lmg_info* out);
int  lmg_genome_name(const lmg_index* idx, uint64_t genome, const char** name);
void lmg_index_close(lmg_index* idx);

/* Index.Search for a batch (lib-index-search.go:1191). seqs: concatenated query sequences (ASCII, any case);
 * seq_off[n+1]. Host buffers in, host rows out; blocking. */
int  lmg_search_batch(lmg_index* idx, const lmg_params* p, const uint8_t* seqs, const uint64_t* seq_off, int32_t n_queries, lmg_results** out);
/* same work with the query batch staged to HBM beforehand (bench "value" leg: inputs resident when the timed region starts) */
typedef struct lmg_queries lmg_queries;
int  lmg_queries_upload(lmg_index* idx, const uint8_t* seqs, const uint64_t* seq_off, int32_t n_queries, lmg_queries** out);
int  lmg_search_staged(lmg_index* idx, const lmg_params* p, lmg_queries* q, lmg_results** out);
void lmg_queries_free(lmg_index* idx, lmg_queries* q);
int  lmg_results_rows(const lmg_results* r, const lmg_hsp** rows, uint64_t* n_rows, const char** strpool, uint64_t* strpool_len);
int  lmg_results_seq_id(const lmg_results* r, uint64_t row, const char** seqid);
void lmg_results_free(lmg_results* r);
/* CUDA-event times of the last search, ms[16]: [0]=h2d [1]=sketch [2]=seed probe [3]=anchor sort/chain [4]=pseudo-align
 * [5]=extend+wfa [6]=host finish [7]=total [8]=k_probe_find kernel alone.  counters[16]: [0..5] probe statistics of the last
 * lmg_anchor_batch (issued, with anchor, search steps, entries scanned, hit records, anchors) [6]=query bases [7]=queries
 * [8]=probe slots [12]=microseconds of the probe-regrouping pass (+ the filter kernel of long queries) [13]=microseconds of the lookup kernel
 * [15]=kernels launched by this library so far */
int  lmg_last_timing(const lmg_index* idx, double* ms16, uint64_t* counters16);
/* byte-model sums of the last lmg_anchor_batch: [0] sum over probes of ceil(log2(n_a+1)), n_a = entries of the probe's anchor run
 * [1] sum of 32-byte sectors of matched entries [2] sum of matched values (anchors before the query-location product) [3] reserved */
int  lmg_probe_model(const lmg_index* idx, uint64_t* sums4);

/* e-values use the index's total bases (lib-index-search.go:1918). A search over genome shards (one index per shard, results merged as
 * `lexicmap utils merge-search-results` does, merge-search-results.go:143-153) sets the total of ALL shards here so that every shard reports
 * the e-values of the whole collection. */
int  lmg_index_set_total_bases(lmg_index* idx, int64_t total_bases);
/* wall-clock milliseconds of lmg_index_open: [0] genomes [1] seed chunks: upload + counting pass [2] seed chunks: fill pass + anchors [3] total */
int  lmg_index_load_times(const lmg_index* idx, double* ms4);

/* ---- seed-lookup microbenchmark (BASELINE.json configs[4]): a synthetic seeds-only image of `masks` buckets x `per_mask` sorted k-mers held for the
 * mask range [mask_lo, mask_hi) (range partitioning across GPUs), and a run of the index-lookup kernel over n_queries synthetic 31-mers (one prefix and one
 * suffix probe each; only probes of the held mask range are issued). out16: [0] probes issued [1] probes with an anchor (kernel work items) [2] mean kernel ms
 * [3] hit records [4] sum ceil(log2(n_a+1)) [5] sum of 32-B sectors of matched entries [6] sum of matched values [7] search steps taken [8] entries scanned
 * [9] generator ms [10] best kernel ms [11] mean ms of the pass that regroups the probes by bucket before the kernel. Search entry points refuse such an index. */
int  lmg_index_synth(int device, int32_t masks, uint64_t per_mask, uint64_t seed, int32_t mask_lo, int32_t mask_hi, int32_t with_values, lmg_index** out);
int  lmg_probe_bench(lmg_index* idx, uint64_t n_queries, uint64_t seed, int32_t min_prefix, int32_t iters, double* out16);

/* random 32-byte-sector read rate of the device (the physical ceiling of the seed lookup, whose accesses are dependent random sectors):
 * n_threads x per_thread independent random reads of a `bytes`-sized buffer. out4: [0] sectors/s [1] GB/s at 32 B per access [2] best ms [3] mean ms */
int  lmg_gather_bench(int device, uint64_t bytes, uint64_t n_threads, int32_t per_thread, int32_t iters, double* out4);

/* ---- stage-wise entry points (parity tests; mirror a1-a7 of SURVEY.md §8a) ---- */
/* lexichash mask + DUST filter + suffix re-masking (lib-index-search.go:1212-1350): kmers[n*m], nlocs[n*m], minloc[n*m];
 * suffix triples (query,new_mask,old_mask,kmer) flattened into suf[4*cap], *n_suf written. */
int  lmg_mask_batch(lmg_index* idx, const uint8_t* seqs, const uint64_t* seq_off, int32_t n, uint64_t* kmers, uint32_t* nlocs, uint32_t* minloc,
                    uint64_t* suf, uint64_t suf_cap, uint64_t* n_suf);
/* seed lookup + anchor materialisation (kv-searcher.go:190-1088, lib-index-search.go:1357-1569), canonical order */
int  lmg_anchor_batch(lmg_index* idx, const lmg_params* p, const uint8_t* seqs, const uint64_t* seq_off, int32_t n, lmg_anchor** out, uint64_t* n_out);
/* ClearSubstrPairs + Chainer.Chain (lib-index-search.go:864-990, lib-chaining.go:122-633) */
int  lmg_chain_batch(lmg_index* idx, const lmg_params* p, const uint8_t* seqs, const uint64_t* seq_off, int32_t n, lmg_chain** out, uint64_t* n_out);
/* window geometry + SeqComparator.Index/Compare + Chainer2 (a9-a12: lib-index-search.go:1987-2051, lib-seq_compare.go:115-159,:335-522, lib-chaining2.go) */
int  lmg_pseudoalign_batch(lmg_index* idx, const lmg_params* p, const uint8_t* seqs, const uint64_t* seq_off, int32_t n, lmg_pa** out, uint64_t* n_out);
/* WFA batch (wfa.Aligner.Align): pairs of ASCII sequences -> CIGAR strings in wfa convention, '\n' separated */
int  lmg_wfa_batch(int device, const uint8_t* seqs, const uint64_t* off /*2n+1*/, int32_t n, int32_t adaptive, char** cigars, uint64_t* cigars_len);
void lmg_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
