// oracle_core.hpp — CPU restatement of LexicMap's query-side search path.  TEST INFRASTRUCTURE ONLY.
//
// This is the parity oracle (task §③): only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may build, link or call it. The product (lexicmap_b200/) never includes this file.
//
// PARITY PINNING STATUS
//   * formats / seed lookup: pinned by the reference's own known-answer test kv/kv-data_test.go:30-361
//     (restated in tests/test_oracle_cpu.py::test_kv_known_answer) and util/varint-GB_test.go (round trip).
//   * chaining, pseudo-alignment: the reference tests carry inputs but no expected outputs
//     (lib-chaining_test.go, lib-seq_compare_test.go) -> "parity unpinned" at unit level; pinned end-to-end
//     against demo/q.gene.fasta.lexicmap.tsv rows (tests/test_oracle_cpu.py::test_oracle_reproduces_reference_demo_rows, runs where the demo genomes of /root/reference exist).
//   * LexicHash masking (github.com/shenwei356/lexichash v0.5.5) and WFA (github.com/shenwei356/wfa v0.5.0)
//     are third-party Go modules whose source is NOT under /root/reference -> "parity unpinned": restated from
//     the call sites, the published algorithm (LexicHash: argmin(kmer XOR mask); WFA: Marco-Sola et al. 2021 with
//     WFA2-lib backtrace priorities) and checked against the demo TSV golden rows.
//
// Every function cites the reference file:line it follows (paths relative to /root/reference/lexicmap/cmd).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <cfloat>
#include <string>
#include <vector>
#include <array>
#include <map>
#include <unordered_map>
#include <set>
#include <algorithm>
#include <stdexcept>

namespace lmo {

// =====================================================================================================
// small helpers
// =====================================================================================================
static inline uint64_t rd_be(const uint8_t* b, int n) { uint64_t v = 0; for (int i = 0; i < n; i++) v = (v << 8) | b[i]; return v; }
static inline std::vector<uint8_t> slurp(const std::string& p) {
  FILE* f = fopen(p.c_str(), "rb"); if (!f) throw std::runtime_error("oracle: cannot open " + p);
  fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET); std::vector<uint8_t> b((size_t)n);
  if (n && fread(b.data(), 1, (size_t)n, f) != (size_t)n) { fclose(f); throw std::runtime_error("oracle: short read " + p); }
  fclose(f); return b;
}
// genome.base2bit (genome/genome.go:1427-1444); also assumed for the lexichash k-mer iterator (kmers package:
// degenerate bases map to their alphabetically first base)
static inline uint8_t b2b(uint8_t c) {
  switch (c) { case 'C': case 'c': case 'B': case 'b': case 'S': case 's': case 'Y': case 'y': return 1;
    case 'G': case 'g': case 'K': case 'k': return 2; case 'T': case 't': case 'U': case 'u': return 3; default: return 0; }
}
static inline uint64_t ns(uint64_t b, int k) { uint64_t c = b; for (int i = 1; i < k; i++) c = (c << 2) + b; return c; }  // util.Ns util/kmers.go:434
static inline uint64_t kmer_reverse(uint64_t c, int k) { uint64_t r = 0; for (int i = 0; i < k; i++) { r = (r << 2) | (c & 3); c >>= 2; } return r; }  // kmers.MustReverse
// util.IsLowComplexityDust util/kmers.go:162-328
static inline bool dust(uint64_t code, int k) {
  uint8_t cnt[64] = {0}; for (int i = 0; i <= k - 2; i++) cnt[(code >> (i << 1)) & 63]++;
  uint16_t score = 0; for (int i = 0; i < 64; i++) { uint16_t c = cnt[i]; score += (uint16_t)((uint16_t)(c - 1) * c) >> 1; }
  return score > 50;
}
static inline int lz64(uint64_t x) { return x ? __builtin_clzll(x) : 64; }

// =====================================================================================================
// parameters (IndexSearchingOptions lib-index-search.go:57-106, SeqComparatorOptions lib-seq_compare.go:34-46,
// Chaining2Options lib-chaining2.go:29-39; defaults search.go:631-731, :305-382)
// =====================================================================================================
struct Params {
  int min_prefix = 15, min_single_prefix = 17, top_n_genomes = 0, top_n_chains = 0;
  float max_gap = 50, max_distance = 1000;
  int ext_len = 1000, ext_len2 = 50;
  double min_qcov_genome = 0, max_evalue = 10;
  int align_max_gap = 20, align_min_len = 50, align_band = 100;
  double min_pident = 70, min_qcov_hsp = 0;
  int output_seq = 0;
  int wfa_adaptive = 1;
};

struct Sub { int32_t q, t; uint8_t len; bool trc, qrc; };  // SubstrPair lib-index-search.go:805-817

// =====================================================================================================
// index (open: lib-index-search.go:237-757)
// =====================================================================================================
struct KvChunkFile {  // kv.Searcher kv/kv-searcher.go:42-61 — file bytes + dense anchor tables
  int k = 0, chunk_index = 0, chunk_size = 0, mask_prefix = 0, anchor_prefix = 0; bool use7 = false;
  std::vector<uint8_t> data; std::vector<std::vector<uint64_t>> indexes;  // per mask: [2 + 2*4^anchor_prefix]
};
struct GenomeBatchFile { std::vector<uint8_t> data; std::vector<uint64_t> rec_off; std::vector<uint32_t> nbases; };
struct GenomeMeta { int genome_size = 0, num_seqs = 0; std::vector<int> seq_sizes; std::vector<std::string> seq_ids; uint64_t seq_offset = 0; };

struct Index {
  int k = 31, n_masks = 0, contig_interval = 1000, mask_prefix = 7, anchor_prefix = 6; int64_t total_bases = 0;
  std::vector<uint64_t> masks; std::vector<KvChunkFile> chunks; std::vector<GenomeBatchFile> batches;
  std::map<uint64_t, std::string> id2name;
  // split genomes (genomes.chunks.bin, lib-index-build.go:1787-1812 / readGenomeChunksLists :2193-2246): batch+genome index -> {#chunks, chunk index, group}
  struct ChunkInfo { uint32_t n, i, group; }; std::map<uint64_t, ChunkInfo> genome_chunks;

  static std::map<std::string, std::string> toml(const std::string& file) {
    std::map<std::string, std::string> kv; FILE* f = fopen(file.c_str(), "r"); if (!f) throw std::runtime_error("oracle: cannot open " + file);
    char line[512], key[128], val[256]; while (fgets(line, sizeof line, f)) { if (line[0] == '#') continue; if (sscanf(line, " %127[^ =] = %255[^\n]", key, val) == 2) kv[key] = val; }
    fclose(f); return kv;
  }
  // kv.readKVIndex kv/kv-data.go:631-769
  static void read_kv_index(const std::string& file, KvChunkFile& c) {
    std::vector<uint8_t> x = slurp(file); if (x.size() < 32 || memcmp(x.data(), ".kvindex", 8)) throw std::runtime_error("oracle: bad kv index");
    c.k = x[10]; c.mask_prefix = x[11]; c.anchor_prefix = x[12]; c.use7 = x[13] & 1; c.chunk_index = (int)rd_be(&x[16], 8); c.chunk_size = (int)rd_be(&x[24], 8);
    size_t index_size = 2 + (((size_t)1 << (2 * c.anchor_prefix)) << 1); int shift = (c.k - c.mask_prefix - c.anchor_prefix) << 1; uint64_t am = ((uint64_t)1 << (c.anchor_prefix << 1)) - 1;
    size_t p = 32; c.indexes.resize(c.chunk_size);
    for (int i = 0; i < c.chunk_size; i++) {
      uint64_t n = rd_be(&x[p], 8); p += 8; if (n == 0) continue;
      std::vector<uint64_t>& idx = c.indexes[i]; idx.assign(index_size, 0);
      for (uint64_t j = 0; j < n; j++, p += 16) { uint64_t kmer = rd_be(&x[p], 8), off = rd_be(&x[p + 8], 8); size_t _j = (j == 0) ? 0 : ((((kmer >> shift) & am) << 1) + 2); idx[_j] = kmer; idx[_j + 1] = off; }
    }
  }
  void open(const std::string& dir) {
    auto info = toml(dir + "/info.toml"); if (atoi(info["main-version"].c_str()) != 3) throw std::runtime_error("oracle: main-version must be 3");
    k = atoi(info["max-K"].c_str()); n_masks = atoi(info["masks"].c_str()); total_bases = atoll(info["input-bases"].c_str()); contig_interval = atoi(info["contig-interval"].c_str());
    int nchunks = atoi(info["chunks"].c_str()), nbatches = atoi(info["genome-batches"].c_str()), partitions = atoi(info["index-partitions"].c_str());
    mask_prefix = std::max((int)(std::log2((double)n_masks) / 2), 1); anchor_prefix = std::max((int)(std::log2((double)partitions) / 2), 1);  // :467-469
    { std::vector<uint8_t> m = slurp(dir + "/masks.bin"); size_t n = (m.size() - 32) / 8; masks.resize(n); for (size_t i = 0; i < n; i++) masks[i] = rd_be(&m[32 + 8 * i], 8); }
    if ((int)masks.size() != n_masks) throw std::runtime_error("oracle: masks.bin size mismatch");
    chunks.resize(nchunks);
    for (int c = 0; c < nchunks; c++) { char b[64]; snprintf(b, sizeof b, "/seeds/chunk_%03d.bin", c); chunks[c].data = slurp(dir + b); read_kv_index(dir + b + ".idx", chunks[c]); }
    batches.resize(nbatches);
    for (int b = 0; b < nbatches; b++) { char s[64]; snprintf(s, sizeof s, "/genomes/batch_%04d/genomes.bin", b); batches[b].data = slurp(dir + s); std::vector<uint8_t> x = slurp(dir + s + ".idx");
      uint32_t n = (uint32_t)rd_be(&x[20], 4); for (uint32_t i = 0; i < n; i++) { batches[b].rec_off.push_back(rd_be(&x[24 + 12 * i], 8)); batches[b].nbases.push_back((uint32_t)rd_be(&x[32 + 12 * i], 4)); } }
    { FILE* f = fopen((dir + "/genomes.chunks.bin").c_str(), "rb"); if (f) { fclose(f); std::vector<uint8_t> d = slurp(dir + "/genomes.chunks.bin"); size_t p = 0; uint32_t grp = 0;
        while (p + 8 <= d.size()) { uint64_t n = rd_be(&d[p], 8); p += 8; if (p + 8 * n > d.size()) throw std::runtime_error("oracle: broken genome chunk file"); for (uint64_t i = 0; i < n; i++, p += 8) genome_chunks[rd_be(&d[p], 8)] = {(uint32_t)n, (uint32_t)i, grp}; grp++; } } }
    { std::vector<uint8_t> d = slurp(dir + "/genomes.map.bin"); size_t p = 0; while (p + 2 <= d.size()) { size_t l = rd_be(&d[p], 2); p += 2; std::string id((const char*)&d[p], l); p += l; id2name[rd_be(&d[p], 8)] = id; p += 8; } }
  }
};

// =====================================================================================================
// (1) masking — lexichash.MaskKnownDistinctPrefixes(s, nil, true) (EXTERNAL, call lib-index-search.go:1212-1220)
// Definition restated: for each mask the k-mer of either strand minimising (kmer XOR mask); all positions where it
// occurs; loc = pos<<1 | strand. Two implementations: brute force (the definition) and a sorted-array descent.
// =====================================================================================================
struct MaskResult { std::vector<uint64_t> kmers; std::vector<std::vector<int>> locs; };

static inline void all_kmers(const uint8_t* s, int n, int k, std::vector<uint64_t>& fw, std::vector<uint64_t>& rc) {
  fw.clear(); rc.clear(); if (n < k) return; uint64_t km = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1), f = 0, r = 0;
  for (int i = 0; i < n; i++) { uint64_t c = b2b(s[i]); f = ((f << 2) | c) & km; r = (r >> 2) | ((3 - c) << (2 * (k - 1))); if (i >= k - 1) { fw.push_back(f); rc.push_back(r); } }
}
static inline MaskResult mask_bruteforce(const Index& ix, const uint8_t* s, int n) {
  MaskResult R; R.kmers.assign(ix.n_masks, 0); R.locs.assign(ix.n_masks, {}); std::vector<uint64_t> fw, rc; all_kmers(s, n, ix.k, fw, rc);
  for (int j = 0; j < ix.n_masks; j++) { uint64_t best = ~0ull, m = ix.masks[j];
    for (size_t p = 0; p < fw.size(); p++) for (int st = 0; st < 2; st++) { uint64_t km = st ? rc[p] : fw[p], h = km ^ m;
      if (h < best) { best = h; R.kmers[j] = km; R.locs[j].clear(); } if (h == best) R.locs[j].push_back((int)(p << 1 | st)); } }
  return R;
}
// argmin over a sorted array of (a[i] XOR x): bitwise trie descent by binary search
static inline size_t xor_argmin_sorted(const uint64_t* a, size_t lo, size_t hi, uint64_t x) {
  for (int bit = 63; bit >= 0 && hi - lo > 1; bit--) { uint64_t b = 1ull << bit; if (((a[lo] ^ a[hi - 1]) & b) == 0) continue;
    size_t l = lo, h = hi; while (l < h) { size_t m = (l + h) >> 1; if (a[m] & b) h = m; else l = m + 1; } if (x & b) lo = l; else hi = l; }
  return lo;
}
static inline MaskResult mask_fast(const Index& ix, const uint8_t* s, int n) {
  MaskResult R; R.kmers.assign(ix.n_masks, 0); R.locs.assign(ix.n_masks, {}); std::vector<uint64_t> fw, rc; all_kmers(s, n, ix.k, fw, rc); if (fw.empty()) return R;
  std::vector<std::pair<uint64_t, int>> e; e.reserve(2 * fw.size()); for (size_t p = 0; p < fw.size(); p++) { e.push_back({fw[p], (int)(p << 1)}); e.push_back({rc[p], (int)(p << 1 | 1)}); }
  std::sort(e.begin(), e.end()); std::vector<uint64_t> keys; std::vector<size_t> first; for (size_t i = 0; i < e.size(); i++) if (i == 0 || e[i].first != e[i - 1].first) { keys.push_back(e[i].first); first.push_back(i); } first.push_back(e.size());
  for (int j = 0; j < ix.n_masks; j++) { size_t u = xor_argmin_sorted(keys.data(), 0, keys.size(), ix.masks[j]); R.kmers[j] = keys[u]; for (size_t i = first[u]; i < first[u + 1]; i++) R.locs[j].push_back(e[i].second); }
  return R;
}
static inline bool low_complexity(uint64_t kmer, int k) {  // lib-index-search.go:1222-1238
  uint64_t ttt = (1ull << (k << 1)) - 1; return kmer == ns(1, k) || kmer == ns(2, k) || kmer == ttt || dust(kmer, k);
}

// =====================================================================================================
// (2) seed lookup — kv.Searcher.Search / Search2, on-disk semantics (kv/kv-searcher.go:190-635, :638-1088)
// =====================================================================================================
struct KvHit { int iquery, iquery2; uint8_t len; bool is_suffix; std::vector<uint64_t> values; };

// one probe: scan the mask's records from the anchor offset, exactly as the Go code walks the file
static inline void kv_probe(const KvChunkFile& c, int iq /*mask index within chunk*/, uint64_t kmer, int p, bool reversed, int iquery2, std::vector<KvHit>& out, bool check_flag = true) {
  const std::vector<uint64_t>& index = c.indexes[iq]; if (index.empty() || kmer == 0) return;   // :285-296
  const int k = c.k; const uint8_t shift = (uint8_t)(k - 32); const uint8_t rvflag = reversed ? 1 : 0; const int vb = c.use7 ? 7 : 8;
  uint64_t left, right; if (p < k) { int s2 = (k - p) << 1; uint64_t mask = (1ull << s2) - 1; left = kmer & ~mask; right = ((kmer >> s2) << s2) | mask; } else { left = right = kmer; }  // :298-304
  uint64_t anchor = (left >> ((k - c.mask_prefix - c.anchor_prefix) << 1)) & ((1ull << (c.anchor_prefix << 1)) - 1);  // AnchorExtracter kv-data.go:319-325
  size_t i = (size_t)(anchor << 1) + 2; uint64_t offset = index[i + 1]; bool is2nd = offset & 1; offset >>= 1; if (offset == 0) return;  // :349-355
  const uint8_t* d = c.data.data(); size_t pos = (size_t)offset; bool first = true, found = false; uint64_t _offset = 0, kmer1, kmer2;
  auto take = [&](uint64_t km, uint64_t nval, bool save) {  // value block of one k-mer (:460-529 / :553-622)
    if (save && check_flag) { if ((d[pos + vb - 1] & 1) != rvflag) save = false; }   // first value's reverse flag decides (:466-488)
    if (save) { KvHit h; h.iquery = iq + c.chunk_index; h.iquery2 = iquery2; h.len = (uint8_t)((uint8_t)(lz64(kmer ^ km) >> 1) + shift); h.is_suffix = reversed;
      for (uint64_t j = 0; j < nval; j++) h.values.push_back(rd_be(d + pos + j * vb, vb)); out.push_back(std::move(h)); }
    pos += (size_t)nval * vb;
  };
  for (;;) {
    uint8_t ctrl = d[pos++]; bool lastPair = ctrl & 128, hasKmer2 = (ctrl & 64) == 0; ctrl &= 63;
    int b1 = ((ctrl >> 3) & 7) + 1, b2 = (ctrl & 7) + 1; uint64_t v1 = rd_be(d + pos, b1), v2 = rd_be(d + pos + b1, b2); pos += b1 + b2;
    if (first) { first = false; if (!is2nd) { kmer1 = index[i]; kmer2 = kmer1 + v2; } else { kmer1 = 0; kmer2 = index[i]; } } else { kmer1 = v1 + _offset; kmer2 = kmer1 + v2; }  // :400-413
    _offset = kmer2;
    if (kmer1 > right) break;                                       // :421
    if (kmer1 >= left || kmer2 >= left) found = true;              // :428
    ctrl = d[pos++]; b1 = ((ctrl >> 3) & 7) + 1; b2 = (ctrl & 7) + 1; uint64_t n1 = rd_be(d + pos, b1), n2 = rd_be(d + pos + b1, b2); pos += b1 + b2;
    take(kmer1, n1, found && kmer1 >= left);                        // :460
    if (kmer2 > right) break;                                       // :531
    if (lastPair && !hasKmer2) break;                               // :536
    take(kmer2, n2, found);                                         // :553
    if (lastPair) break;                                            // :624
  }
}

// =====================================================================================================
// ClearSubstrPairs lib-index-search.go:864-990 (ties beyond the reference comparator broken by (qrc,trc) so the
// result is deterministic; the reference's slices.SortFunc is unstable there)
// =====================================================================================================
static inline void clear_subs(std::vector<Sub>& subs, int k) {
  std::sort(subs.begin(), subs.end(), [](const Sub& a, const Sub& b) {
    if (a.q != b.q) return a.q < b.q; if (a.len != b.len) return a.len > b.len; if (a.t != b.t) return a.t < b.t; if (a.qrc != b.qrc) return a.qrc < b.qrc; return a.trc < b.trc; });
  size_t n = subs.size(); std::vector<char> mark(n, 0);
  for (size_t i = 0; i + 1 < n; i++) { const Sub& v = subs[i + 1]; int32_t vQEnd = v.q + v.len, up = std::max(vQEnd - k, 0), vTEnd = v.t + v.len;
    size_t start = std::lower_bound(subs.begin(), subs.begin() + i + 1, up, [](const Sub& s, int32_t t) { return s.q < t; }) - subs.begin();
    for (size_t j = start; j <= i; j++) { const Sub& p = subs[j]; if (vQEnd <= p.q + p.len && v.t >= p.t && vTEnd <= p.t + p.len) { mark[i + 1] = 1; break; } } }
  size_t j = 0; for (size_t i = 0; i < n; i++) if (!mark[i]) subs[j++] = subs[i]; if (j > 0) subs.resize(j);
}

// =====================================================================================================
// (3) Chainer.Chain lib-chaining.go:122-633 (live branch :339-477); float32 arithmetic, no FMA contraction
// =====================================================================================================
#pragma GCC push_options
#pragma GCC optimize("fp-contract=off")
static inline float seed_weight(float l) { return 0.1f * l * l; }                                   // :635
static inline float gap_score(float g) { if (g == 0) return 0; return 0.1f * g + 0.5f * (float)std::log2((double)g); }  // :662
static inline float gap_f(const Sub& a, const Sub& b) {                                               // :655
  if (a.t >= b.t) return (float)std::fabs(std::fabs((double)(a.q - b.q)) - std::fabs((double)(a.t - b.t)));
  return (float)std::fabs(std::fabs((double)(a.q - b.q)) - std::fabs((double)(a.t + (int32_t)a.len - b.t - (int32_t)b.len)));
}
struct ChainOpts { float max_gap, min_score, max_distance; int top_chains; };
static inline std::vector<std::vector<int32_t>> chain1(const std::vector<Sub>& subs, const ChainOpts& o, float* score_out) {
  std::vector<std::vector<int32_t>> paths; int n = (int)subs.size();
  if (n == 1) { float w = seed_weight((float)subs[0].len); if (w >= o.min_score) paths.push_back({0}); *score_out = w; return paths; }  // :125-138
  std::vector<uint64_t> msi(n), s2i(n); std::vector<int8_t> dirs(n, 0);
  { float s = seed_weight((float)subs[0].len); uint32_t b; memcpy(&b, &s, 4); msi[0] = (uint64_t)b << 32; s2i[0] = (uint64_t)b << 32; }
  const int32_t maxDist = (int32_t)o.max_distance;
  std::vector<uint64_t> ri(n); for (int i = 0; i < n; i++) ri[i] = ((uint64_t)(uint32_t)subs[i].t << 32) | (uint32_t)i; std::sort(ri.begin(), ri.end());  // rangeindex
  std::vector<int32_t> js;
  for (int i = 1; i < n; i++) {
    const Sub& a = subs[i]; float m = seed_weight((float)a.len); int mj = i; int8_t mdir = 0;
    uint32_t start = (a.t < maxDist) ? 0u : (uint32_t)(a.t - maxDist), end = (uint32_t)(a.t + maxDist);
    auto lo = std::lower_bound(ri.begin(), ri.end(), (uint64_t)start << 32), hi = std::upper_bound(ri.begin(), ri.end(), ((uint64_t)end << 32) | 0xffffffffu);
    js.clear(); for (auto it = lo; it != hi; ++it) js.push_back((int32_t)(*it & 0xffffffffu)); std::sort(js.begin(), js.end());
    for (int _j = (int)js.size() - 1; _j >= 0; _j--) { int j = js[_j]; if (j >= i) continue; const Sub& b = subs[j];
      if (a.q == b.q || a.t == b.t) continue; if (a.q - b.q > maxDist) break;
      float g = gap_f(a, b); if (g > o.max_gap) continue;
      float w; int32_t length;
      if (a.q > b.q + (int32_t)b.len) { length = a.len; w = seed_weight((float)length); }
      else if (g == 0) { length = a.q + a.len - b.q; w = -seed_weight((float)b.len) + seed_weight((float)length); }
      else { length = a.q + a.len - (b.q + b.len); w = seed_weight((float)length); }
      int8_t dir = (a.t >= b.t) ? 1 : -1; float s;
      if (dirs[j] == 0 || dirs[j] == dir) { uint32_t bits = (uint32_t)(msi[j] >> 32); float pm; memcpy(&pm, &bits, 4); s = pm + w - gap_score(g); }
      else s = seed_weight((float)b.len) + w - gap_score(g);
      if (s >= o.min_score && s > m) { m = s; mj = j; mdir = dir; } }
    uint32_t mb; memcpy(&mb, &m, 4); msi[i] = ((uint64_t)mb << 32) | (uint32_t)mj; dirs[i] = mdir; s2i[i] = ((uint64_t)mb << 32) | (uint32_t)i;
  }
  std::vector<char> visited(n, 0); std::sort(s2i.begin(), s2i.end()); int iMax = n - 1; float maxScore = 0; bool first = true; int nChecked = 0;
  for (;;) {   // backtrack :520-629
    nChecked++; if (o.top_chains > 0 && nChecked > o.top_chains) break;
    float M = 0; uint32_t Mi = 0;
    while (iMax >= 0) { uint32_t bits = (uint32_t)(s2i[iMax] >> 32); memcpy(&M, &bits, 4); Mi = (uint32_t)s2i[iMax]; if (!visited[Mi]) { iMax--; break; } iMax--; }
    if (M < o.min_score) break;
    std::vector<int32_t> path; int i = (int)Mi; if (first) { maxScore = M; first = false; }
    for (;;) { int j = (int)(msi[i] & 0xffffffffu); bool change = (i != j && dirs[j] != 0 && dirs[i] != dirs[j]);
      if (visited[j] && !change) { path.clear(); visited[i] = 1; break; }
      path.push_back(i); visited[i] = 1;
      if (i == j || change) { if (change) path.push_back(j); std::reverse(path.begin(), path.end()); paths.push_back(path); path.clear(); break; } else i = j; }
  }
  *score_out = maxScore; return paths;
}
#pragma GCC pop_options

// =====================================================================================================
// (4) pseudo-alignment: SeqComparator.Index / Compare (lib-seq_compare.go:115-159, :335-522),
//     tree.Search semantics (tree/tree.go:441-527) on a sorted array incl. the uint8-wrap quirk (:498-501),
//     TrimSubStrPairs (:553-621), Chainer2 (lib-chaining2.go:152-658)
// =====================================================================================================
struct QueryTable { int k; std::vector<uint64_t> keys; std::vector<uint32_t> voff; std::vector<uint32_t> vals; };
static inline QueryTable build_query_table(const uint8_t* s, int n, int k) {
  QueryTable T; T.k = k; std::vector<uint64_t> fw, rc; all_kmers(s, n, k, fw, rc); std::vector<std::pair<uint64_t, uint32_t>> e;
  uint64_t ccc = ns(1, k), ggg = ns(2, k), ttt = ns(3, k);
  for (size_t p = 0; p < fw.size(); p++) { uint64_t km = fw[p]; if (km == 0 || km == ccc || km == ggg || km == ttt || dust(km, k)) continue; e.push_back({km, (uint32_t)(p << 1)}); e.push_back({rc[p], (uint32_t)(p << 1 | 1)}); }
  std::stable_sort(e.begin(), e.end(), [](const std::pair<uint64_t, uint32_t>& a, const std::pair<uint64_t, uint32_t>& b) { return a.first < b.first; });
  for (size_t i = 0; i < e.size(); i++) { if (i == 0 || e[i].first != e[i - 1].first) { T.keys.push_back(e[i].first); T.voff.push_back((uint32_t)T.vals.size()); } T.vals.push_back(e[i].second); }
  T.voff.push_back((uint32_t)T.vals.size()); return T;
}
static inline int lcp_bases(uint64_t a, uint64_t b, int k) { return std::min(k, (lz64(a ^ b) >> 1) + k - 32); }
// tree.Search: emulate the radix-tree descent on the sorted key array. Returns [lo,hi) of keys reported.
static inline bool tree_search(const QueryTable& T, uint64_t key, int p, size_t* rlo, size_t* rhi) {
  const int K = T.k; if (p < 1) p = 1; if (p > K) p = K; size_t lo = 0, hi = T.keys.size(); int depth = 0; const uint64_t* a = T.keys.data();
  while (depth < K) {
    // child whose base at `depth` equals the query's
    int sh = 2 * (K - depth - 1); uint64_t qb = (key >> sh) & 3;
    size_t l = lo, h = hi; { size_t x = lo, y = hi; while (x < y) { size_t m = (x + y) >> 1; if (((a[m] >> sh) & 3) < qb) x = m + 1; else y = m; } l = x; y = hi; while (x < y) { size_t m = (x + y) >> 1; if (((a[m] >> sh) & 3) <= qb) x = m + 1; else y = m; } h = x; }
    if (l == h) return false;
    int nodeEnd = (h - l == 1) ? K : lcp_bases(a[l], a[h - 1], K);   // the child's compressed edge covers bases [depth, nodeEnd)
    int nk = nodeEnd - depth; int m = lcp_bases(key, a[l], K);
    if (m >= nodeEnd) { depth = nodeEnd; lo = l; hi = h; if (depth >= p) { *rlo = lo; *rhi = hi; return true; } continue; }
    int atleast = p - depth;  // uint8 in Go; if atleast > nk the shift wraps and n.prefix>>huge == 0
    bool hit; if (atleast <= nk) hit = (m >= depth + atleast); else { uint64_t nxt = (key >> (2 * (K - depth - atleast))) & ((1ull << (2 * atleast)) - 1); hit = (nxt == 0); }
    if (hit) { *rlo = l; *rhi = h; return true; } return false;
  }
  return false;
}
static inline float distance_f(const Sub& a, const Sub& b) { return (float)std::max(std::fabs((double)(a.q - b.q)), std::fabs((double)(a.t - b.t))); }  // lib-chaining.go:639
static inline double gap2(const Sub& a, const Sub& b) { return std::fabs(std::fabs((double)(a.q - b.q)) - std::fabs((double)(a.t - b.t))); }            // lib-chaining2.go:664
static inline int32_t overlap(const Sub& a, const Sub& b) { int32_t qo = 0, to = 0; if (b.q >= a.q && b.q <= a.q + a.len) qo = a.q + a.len - b.q + 1; if (b.t >= a.t && b.t <= a.t + a.len) to = a.t + a.len - b.t + 1; return std::max(qo, to); }
static inline void trim_subs(std::vector<Sub>& subs, float minDist) {  // TrimSubStrPairs lib-seq_compare.go:553-621
  if (subs.size() < 2) return; int last = (int)subs.size() - 1; Sub _p = subs[0]; int start = 0;
  for (int i = 0; i < last; i++) { const Sub& p = subs[i + 1];   // `for i, p = range (*subs)[1:]` — i is the index in the sub-slice
    if (distance_f(p, _p) < minDist && ((p.q == _p.q || p.t == _p.t) || (gap2(_p, p) > 11 && (double)overlap(_p, p) / (double)_p.len > 0.8))) { start = i; _p = p; continue; } break; }
  _p = subs[last]; int end = last;
  for (int i = (int)subs.size() - 2; i >= 0; i--) { const Sub& p = subs[i];
    if (distance_f(p, _p) < minDist && ((p.q == _p.q || p.t == _p.t) || (gap2(p, _p) > 11 && (double)overlap(p, _p) / (double)_p.len > 0.8))) { end = i; _p = p; continue; } break; }
  if (start >= end) { subs.clear(); return; }
  std::vector<Sub> t(subs.begin() + start, subs.begin() + end + 1); subs.swap(t);
}
struct Chain2 { int n_anchors = 0, matched = 0, aligned_q = 0, aligned_t = 0; double pident = 0; int qb = 0, qe = 0, tb = 0, te = 0;
  // filled later (lib-index-search.go:2083-2626)
  int t_pos_offset_begin = 0, max_ext_len = 0; int aligned_len = 0, gaps = 0, score = 0, bitscore = 0; double evalue = 0, af = 0; bool dead = false;
  std::string cigar, qseq, tseq, align; };
struct Chain2Opts { int max_gap, min_score, min_align_len, band_count, band_base; double kmer_pident_threshold; };
static void chain_a_region(const Sub* subs, const uint64_t* msi, int len, int offset, const Chain2Opts& o, std::vector<Chain2>& paths, int Mi0) {  // lib-chaining2.go:360-658
  int Mi = 0; if (Mi0 < 0) { double M = 0; for (int i = 0; i < len; i++) { double m = (double)(msi[i] >> 32); if (m > M) { M = m; Mi = i; } } if (M < (double)o.min_score) return; } else Mi = Mi0;
  int nMatched = 0, nAQ = 0, nAT = 0, i = Mi, j = 0; int32_t qb = 0, qe = 0, tb = 0, te = 0; int beginOfNext = 0; bool firstAnchor = true; int nAnchors = 0;
  auto emit = [&]() { double pid = (double)nMatched / (double)std::max(nAQ, nAT) * 100; if (pid > 100) pid = 100; Chain2 c; c.n_anchors = nAnchors; c.aligned_q = nAQ; c.aligned_t = nAT; c.matched = nMatched; c.pident = pid; c.qb = qb; c.qe = qe; c.tb = tb; c.te = te; paths.push_back(c); };
  for (;;) {
    j = (int)(msi[i] & 0xffffffffu) - offset; if (j < 0) break;
    const Sub& sub = subs[i]; nAnchors++;
    if (firstAnchor) { firstAnchor = false; qe = sub.q + sub.len - 1; te = sub.t + sub.len - 1; qb = sub.q; tb = sub.t; nMatched += sub.len; }
    else { qb = sub.q; tb = sub.t; if (sub.q + (int)sub.len - 1 >= beginOfNext) nMatched += beginOfNext - sub.q; else nMatched += sub.len; }
    beginOfNext = sub.q;
    if (i == j) { nAQ += qe - qb + 1; if (nAQ < o.min_align_len) break; nAT += te - tb + 1;
      double pid = (double)nMatched / (double)std::max(nAQ, nAT) * 100; if (pid < o.kmer_pident_threshold) break; emit(); break; }
    i = j;
  }
  if (j < 0 && nAnchors > 0) { nAQ += qe - qb + 1; nAT += te - tb + 1; if (nAQ >= o.min_align_len) { double pid = (double)nMatched / (double)std::max(nAQ, nAT) * 100; if (pid >= o.kmer_pident_threshold) emit(); } }
  if (Mi != len - 1) chain_a_region(subs + Mi + 1, msi + Mi + 1, len - Mi - 1, offset + Mi + 1, o, paths, -1);
  if (i > 0) chain_a_region(subs, msi, i, offset, o, paths, -1);
}
static inline std::vector<Chain2> chain2(const std::vector<Sub>& subs, const Chain2Opts& o) {  // Chainer2.Chain lib-chaining2.go:152-358
  std::vector<Chain2> paths; int n = (int)subs.size();
  if (n == 1) { const Sub& s = subs[0]; int sl = s.len; if (sl >= o.min_score && sl >= o.min_align_len) { Chain2 c; c.qb = s.q; c.qe = s.q + sl - 1; c.tb = s.t; c.te = s.t + sl - 1; c.matched = sl; c.pident = 100; c.aligned_q = sl; c.n_anchors = 1; paths.push_back(c); } return paths; }
  std::vector<uint64_t> msi(n); msi[0] = (uint64_t)subs[0].len << 32; double M = 0; int Mi = 0;
  for (int i = 1; i < n; i++) { const Sub& a = subs[i]; double m = a.len; int mj = i, cnt = 0;
    for (int j = i - 1; j >= 0; j--) { const Sub& b = subs[j]; if (b.q == a.q || b.t > a.t) continue; cnt++;
      int32_t base = a.q - b.q - (int32_t)b.len; if (!(base <= o.band_base || cnt <= o.band_count)) break;
      int32_t qd = std::abs(a.q - b.q), td = std::abs(a.t - b.t); double g = (double)std::abs(qd - td); if (g > (double)o.max_gap) continue;
      double s = (double)(msi[j] >> 32) + (double)b.len - g; if (s >= m) { m = s; mj = j; } }
    msi[i] = ((uint64_t)m << 32) | (uint32_t)mj; if (m > M) { M = m; Mi = i; } }
  if (M < (double)o.min_score) return paths;
  chain_a_region(subs.data(), msi.data(), n, 0, o, paths, Mi); return paths;
}
// Compare lib-seq_compare.go:335-522. tseq = 2-bit codes -> ASCII already (upper-case ACGT)
static inline std::vector<Chain2> compare(const QueryTable& T, uint32_t begin, uint32_t end, const std::string& tseq, const Chain2Opts& o, int min_prefix) {
  const int k = T.k; int m = min_prefix; size_t L = tseq.size(); if (L >= 1000000) m += 8; else if (L >= 250000) m += 6; else if (L >= 50000) m += 4; else if (L >= 10000) m += 2;
  std::vector<uint64_t> fw, rc; all_kmers((const uint8_t*)tseq.data(), (int)L, k, fw, rc); std::vector<Sub> subs; uint64_t ccc = ns(1, k), ggg = ns(2, k), ttt = ns(3, k); size_t lo, hi;
  for (size_t idx = 0; idx < fw.size(); idx++) { uint64_t km = fw[idx]; if (km == 0 || km == ccc || km == ggg || km == ttt) continue;
    if (tree_search(T, km, m, &lo, &hi)) for (size_t u = lo; u < hi; u++) { int lp = lcp_bases(km, T.keys[u], k); for (uint32_t vi = T.voff[u]; vi < T.voff[u + 1]; vi++) { uint32_t v = T.vals[vi], p = v >> 1;
          if ((v & 1) == 1 || p < begin || p + (uint32_t)lp > end) continue; subs.push_back({(int32_t)p, (int32_t)idx, (uint8_t)lp, false, false}); } }
    uint64_t kr = rc[idx];
    if (tree_search(T, kr, m, &lo, &hi)) for (size_t u = lo; u < hi; u++) { int lp = lcp_bases(kr, T.keys[u], k); for (uint32_t vi = T.voff[u]; vi < T.voff[u + 1]; vi++) { uint32_t v = T.vals[vi], p = (v >> 1) + (uint32_t)k - (uint32_t)lp;
          if ((v & 1) == 0 || p + (uint32_t)lp < begin || p > end) continue; subs.push_back({(int32_t)p, (int32_t)(idx + k - lp), (uint8_t)lp, true, true}); } }
  }
  std::vector<Chain2> none; if (subs.empty()) return none;
  if (subs.size() > 1) clear_subs(subs, k);
  trim_subs(subs, 100); if (subs.empty()) return none;
  std::vector<Chain2> chains = chain2(subs, o);
  if (chains.size() > 1) std::stable_sort(chains.begin(), chains.end(), [](const Chain2& a, const Chain2& b) { return a.qb < b.qb; });  // :501-508 (insertion sort for n<=12 in Go => stable)
  return chains;
}

// =====================================================================================================
// extendMatch / _extendRight / Chainer3 (lib-index-search-util.go:34-201, lib-chaining3.go:111-299)
// =====================================================================================================
static inline void extend_right(const char* s1, int n1, const char* s2, int n2, int* e1, int* e2) {
  *e1 = *e2 = 0; if (n1 < 2 || n2 < 2) return; std::vector<Sub> subs;
  for (int i2 = 0; i2 + 1 < n2; i2++) { int c2 = b2b(s2[i2]) << 2 | b2b(s2[i2 + 1]); for (int i1 = 0; i1 + 1 < n1; i1++) if ((b2b(s1[i1]) << 2 | b2b(s1[i1 + 1])) == c2) subs.push_back({i1, i2, 2, false, false}); }
  if (subs.empty()) return;
  std::sort(subs.begin(), subs.end(), [](const Sub& a, const Sub& b) { if (a.q != b.q) return a.q < b.q; return a.t < b.t; });
  int n = (int)subs.size(); std::vector<int64_t> sc(n); std::vector<int> pj(n); Sub s0{0, 0, 0, false, false};
  auto d2 = [](const Sub& a, const Sub& b) { return std::max(std::fabs((double)(a.q - b.q)), std::fabs((double)(a.t - b.t))); };
  double M = 0; int Mi = 0; { const Sub& a = subs[0]; double m = (double)a.len - d2(s0, a) - gap2(s0, a); sc[0] = (int64_t)m; pj[0] = 0; }
  for (int i = 1; i < n; i++) { const Sub& a = subs[i]; double m = (double)a.len - d2(s0, a) - gap2(s0, a); int mj = i, cnt = 0;
    for (int j = i - 1; j >= 0; j--) { const Sub& b = subs[j]; if (b.q == a.q || b.t > a.t) continue; cnt++; int32_t base = a.q - b.q - (int32_t)b.len; if (!(base <= 10 || cnt <= 20)) break;
      double d = d2(a, b); if (d > 10) continue; double g = gap2(a, b); if (g > 5) continue; double s = (double)sc[j] + (double)b.len - d - g; if (s >= m) { m = s; mj = j; } }
    sc[i] = (int64_t)m; pj[i] = mj; if (m > M) { M = m; Mi = i; } }
  if (M < 1) return;
  int i = Mi, nMatched = 0, nAQ = 0, nAT = 0, beginOfNext = 0; int32_t qb = 0, qe = 0, tb = 0, te = 0; bool firstA = true;
  for (;;) { int j = pj[i]; const Sub& sub = subs[i];
    if (firstA) { firstA = false; qe = sub.q + sub.len - 1; te = sub.t + sub.len - 1; qb = sub.q; tb = sub.t; nMatched += sub.len; }
    else { qb = sub.q; tb = sub.t; if (sub.q + (int)sub.len - 1 >= beginOfNext) nMatched += beginOfNext - sub.q; else nMatched += sub.len; }
    beginOfNext = sub.q;
    if (i == j) { nAQ += qe - qb + 1; if (nAQ < 2) return; nAT += te - tb + 1; double pid = (double)nMatched / (double)std::max(nAQ, nAT) * 100; if (pid < 15) return; *e1 = qe + 1; *e2 = te + 1; return; }
    i = j; }
}
struct Extended { int start1, end1, start2, end2, s1, e1, s2, e2; };
static inline Extended extend_match(const std::string& seq1, const std::string& seq2, int start1, int end1, int start2, int end2, int extLen, int tBegin, int maxExtLen, bool rc) {
  Extended R{start1, end1, start2, end2, 0, 0, 0, 0}; const int m = 2; int _s1 = start1, _e1 = end1, _s2 = start2, _e2 = end2; int n1 = (int)seq1.size(), n2 = (int)seq2.size();
  if (end1 + m < n1 && end2 + m < n2) { int ext = rc ? std::min(extLen, tBegin) : std::min(extLen, maxExtLen);
    if (ext > 2) { int e1 = std::min(end1 + ext, n1), e2 = std::min(end2 + ext, n2); int a, b; extend_right(seq1.data() + end1, e1 - end1, seq2.data() + end2, e2 - end2, &a, &b); R.e1 = a; R.e2 = b; if (a > 0 || b > 0) { end1 += a; end2 += b; } } }
  if (start1 > m && start2 > m) { int ext = rc ? std::min(extLen, maxExtLen) : std::min(extLen, tBegin);
    if (ext > 2) { int s1 = std::max(start1 - ext, 0), s2 = std::max(start2 - ext, 0); std::string r1(seq1.begin() + s1, seq1.begin() + start1), r2(seq2.begin() + s2, seq2.begin() + start2); std::reverse(r1.begin(), r1.end()); std::reverse(r2.begin(), r2.end());
      int a, b; extend_right(r1.data(), (int)r1.size(), r2.data(), (int)r2.size(), &a, &b); R.s1 = a; R.s2 = b; if (a > 0 || b > 0) { start1 -= a; start2 -= b; } } }
  if (start1 < 0 || start2 < 0) { start1 = _s1; start2 = _s2; } if (end1 > n1 || end2 > n2) { end1 = _e1; end2 = _e2; }
  R.start1 = start1; R.end1 = end1; R.start2 = start2; R.end2 = end2; return R;
}

// =====================================================================================================
// (5) WFA — github.com/shenwei356/wfa v0.5.0 (EXTERNAL; call sites lib-index-search.go:1842,1910,2261,2528).
// Gap-affine wavefront alignment, end-to-end, penalties x=4 o=6 e=2 (validated on demo rows),
// exact (no adaptive pruning), WFA2-lib backtrace priority: mismatch > D ext > D open > I ext > I open.
// Op letters follow the wfa module: 'I' consumes target (text), 'D' consumes query (pattern); LexicMap swaps them
// for SAM (lib-index-search.go:2331-2338).
// =====================================================================================================
struct WfaResult { std::vector<uint64_t> ops; int qbegin = 0, qend = 0, tbegin = 0, tend = 0; int align_len = 0, matches = 0, gaps = 0; int score = 0; };
static const int32_t WF_NULL = INT32_MIN / 2;
struct WF { int lo = 0, hi = -1; std::vector<int32_t> off; bool null = true; int32_t get(int k) const { return (null || k < lo || k > hi) ? WF_NULL : off[k - lo]; } };
// adaptive != 0: WFA-adaptive wavefront reduction as in WFA2-lib (wavefront_heuristic_wfadaptive): after extending the M wavefront of a
// score, if it spans >= min_wf_len diagonals, drop diagonals from both ends whose distance to the end max(plen-v, tlen-h) exceeds the
// best one by more than max_dist_diff; the target diagonal is preserved. The reference enables wfa.DefaultAdaptiveOption
// (lib-index-search.go:1911; MinWFLen 10, MaxDistDiff 50 per the commented alternative at :1912-1915).
static inline WfaResult wfa_align(const char* q, int plen, const char* t, int tlen, int adaptive = 1, int min_wf_len = 10, int max_dist_diff = 50, int X = 4, int O = 6, int E = 2) {
  WfaResult R; std::vector<WF> Mw, Iw, Dw; const int kend = tlen - plen;
  auto extend = [&](int k, int32_t h) { int v = h - k; while (v < plen && h < tlen && q[v] == t[h]) { v++; h++; } return h; };
  auto at = [&](std::vector<WF>& W, int s) -> const WF* { static const WF nullwf; return (s < 0 || s >= (int)W.size()) ? &nullwf : &W[s]; };
  Mw.emplace_back(); Iw.emplace_back(); Dw.emplace_back(); Mw[0].null = false; Mw[0].lo = Mw[0].hi = 0; Mw[0].off = {extend(0, 0)};
  int s = 0;
  while (!(Mw[s].get(kend) >= tlen)) {
    s++; Mw.emplace_back(); Iw.emplace_back(); Dw.emplace_back();
    const WF *mx = at(Mw, s - X), *mo = at(Mw, s - O - E), *ie = at(Iw, s - E), *de = at(Dw, s - E);
    if (mx->null && mo->null && ie->null && de->null) continue;
    int lo = INT32_MAX, hi = INT32_MIN; if (!mx->null) { lo = std::min(lo, mx->lo); hi = std::max(hi, mx->hi); }
    for (const WF* w : {mo, ie, de}) if (!w->null) { lo = std::min(lo, w->lo - 1); hi = std::max(hi, w->hi + 1); }
    WF &M = Mw[s], &I = Iw[s], &D = Dw[s]; M.lo = I.lo = D.lo = lo; M.hi = I.hi = D.hi = hi; int w = hi - lo + 1; M.off.assign(w, WF_NULL); I.off.assign(w, WF_NULL); D.off.assign(w, WF_NULL);
    bool anyM = false, anyI = false, anyD = false;
    for (int k = lo; k <= hi; k++) {
      int32_t ins = std::max(mo->get(k - 1), ie->get(k - 1)); ins = (ins <= WF_NULL) ? WF_NULL : ins + 1;
      int32_t del = std::max(mo->get(k + 1), de->get(k + 1)); if (del < WF_NULL) del = WF_NULL;
      int32_t mis = mx->get(k); mis = (mis <= WF_NULL) ? WF_NULL : mis + 1;
      auto valid = [&](int32_t h) { if (h <= WF_NULL) return false; int v = h - k; return h >= 0 && v >= 0 && h <= tlen && v <= plen; };
      if (!valid(ins)) ins = WF_NULL; if (!valid(del)) del = WF_NULL; if (!valid(mis)) mis = WF_NULL;
      I.off[k - lo] = ins; D.off[k - lo] = del; int32_t mm = std::max(mis, std::max(ins, del));
      if (mm > WF_NULL) { mm = extend(k, mm); anyM = true; } M.off[k - lo] = mm; anyI |= ins > WF_NULL; anyD |= del > WF_NULL;
    }
    M.null = !anyM; I.null = !anyI; D.null = !anyD;
    if (adaptive && !M.null && hi - lo + 1 >= min_wf_len) {
      int mind = INT32_MAX; std::vector<int> dist(w); for (int k = lo; k <= hi; k++) { int32_t o = M.off[k - lo]; int d = (o > WF_NULL) ? std::max(plen - (o - k), tlen - o) : INT32_MAX; dist[k - lo] = d; mind = std::min(mind, d); }
      int nlo = lo, nhi = hi; int top_limit = std::min(kend, hi); for (int k = lo; k < top_limit; k++) { if (dist[k - lo] != INT32_MAX && dist[k - lo] - mind <= max_dist_diff) break; nlo++; }
      int bottom_limit = std::max(kend, nlo); for (int k = hi; k > bottom_limit; k--) { if (dist[k - lo] != INT32_MAX && dist[k - lo] - mind <= max_dist_diff) break; nhi--; }
      if (nlo != lo || nhi != hi) { int nw = nhi - nlo + 1; for (WF* W : {&M, &I, &D}) { std::vector<int32_t> o2(W->off.begin() + (nlo - lo), W->off.begin() + (nlo - lo) + nw); W->off.swap(o2); W->lo = nlo; W->hi = nhi; }
        bool aM = false, aI = false, aD = false; for (int i = 0; i < nw; i++) { aM |= M.off[i] > WF_NULL; aI |= I.off[i] > WF_NULL; aD |= D.off[i] > WF_NULL; } M.null = !aM; I.null = !aI; D.null = !aD; }
    }
  }
  R.score = s;
  // backtrace (WFA2 wavefront_backtrace_affine)
  std::string ops; int k = kend; int32_t off = tlen; int sc = s; enum { MM, II, DD } mat = MM; int v = off - k, h = off;
  auto pig = [](int32_t o, int type) -> int64_t { return o <= WF_NULL ? INT64_MIN : (((int64_t)o << 4) | type); };
  while (v > 0 && h > 0 && sc > 0) {
    int s_mis = sc - X, s_open = sc - O - E, s_ext = sc - E; int64_t best;
    int64_t c_mis = INT64_MIN, c_io = INT64_MIN, c_ie = INT64_MIN, c_do = INT64_MIN, c_de = INT64_MIN;
    if (mat == MM) { int32_t o = at(Mw, s_mis)->get(k); c_mis = pig(o <= WF_NULL ? WF_NULL : o + 1, 9); }
    if (mat == MM || mat == II) { int32_t o = at(Mw, s_open)->get(k - 1); c_io = pig(o <= WF_NULL ? WF_NULL : o + 1, 1); o = at(Iw, s_ext)->get(k - 1); c_ie = pig(o <= WF_NULL ? WF_NULL : o + 1, 2); }
    if (mat == MM || mat == DD) { c_do = pig(at(Mw, s_open)->get(k + 1), 5); c_de = pig(at(Dw, s_ext)->get(k + 1), 6); }
    best = std::max(c_mis, std::max(std::max(c_io, c_ie), std::max(c_do, c_de)));
    if (best == INT64_MIN) throw std::runtime_error("oracle wfa: backtrace failed");
    if (mat == MM) { int32_t mo = (int32_t)(best >> 4); int nm = off - mo; ops.append(nm, 'M'); off = mo; v = off - k; h = off; if (v <= 0 || h <= 0) continue; }
    int type = (int)(best & 15);
    switch (type) { case 9: sc = s_mis; mat = MM; ops.push_back('X'); off--; break;
      case 1: sc = s_open; mat = MM; ops.push_back('I'); k--; off--; break; case 2: sc = s_ext; mat = II; ops.push_back('I'); k--; off--; break;
      case 5: sc = s_open; mat = MM; ops.push_back('D'); k++; break; case 6: sc = s_ext; mat = DD; ops.push_back('D'); k++; break; }
    v = off - k; h = off;
  }
  if (sc == 0) ops.append(off, 'M'); else { while (v > 0) { ops.push_back('D'); v--; } while (h > 0) { ops.push_back('I'); h--; } }
  std::reverse(ops.begin(), ops.end());
  // run-length ops + first-M..last-M statistics (AlignmentResult fields as used at lib-index-search.go:2278-2302)
  for (size_t i = 0; i < ops.size();) { size_t j = i; while (j < ops.size() && ops[j] == ops[i]) j++; R.ops.push_back(((uint64_t)ops[i] << 32) | (uint64_t)(j - i)); i = j; }
  int qi = 0, ti = 0; bool seen = false; int alen_run = 0, gaps_run = 0;
  for (char c : ops) { if (c == 'M') { if (!seen) { seen = true; R.qbegin = qi + 1; R.tbegin = ti + 1; alen_run = 0; gaps_run = 0; } }
    if (seen) { alen_run++; if (c == 'I' || c == 'D') gaps_run++; }
    if (c == 'M' || c == 'X') { qi++; ti++; } else if (c == 'I') ti++; else qi++;
    if (c == 'M') { R.matches++; R.qend = qi; R.tend = ti; R.align_len = alen_run; R.gaps = gaps_run; } }
  return R;
}
// trimOps lib-index-search-util.go:239-258
static inline std::vector<uint64_t> trim_ops(const std::vector<uint64_t>& ops) { int st = -1, en = -1; for (size_t i = 0; i < ops.size(); i++) if ((ops[i] >> 32) == 'M') { st = (int)i; break; } for (int i = (int)ops.size() - 1; i >= 0; i--) if ((ops[i] >> 32) == 'M') { en = i; break; } if (st < 0) return {}; return std::vector<uint64_t>(ops.begin() + st, ops.begin() + en + 1); }
// scoreAndEvalue lib-index-search-util.go:260-304 with (2,-3,5,2,totalBases,0.625,0.41)
static inline void score_evalue(const WfaResult& w, int qlen, int64_t totalBases, int* score, int* bitscore, double* evalue) {
  std::vector<uint64_t> ops = trim_ops(w.ops); if (ops.empty()) { *score = 0; *bitscore = 0; *evalue = DBL_MAX; return; }
  int sc = 0; for (uint64_t op : ops) { int n = (int)(op & 0xffffffffu); switch (op >> 32) { case 'M': sc += n * 2; break; case 'X': sc += n * -3; break; case 'I': case 'D': case 'H': sc -= 5 + n * 2; break; } }
  int _s = sc; if (_s & 1) _s--; double bs = (0.625 * (double)_s - std::log(0.41)) / M_LN2; *score = sc; *bitscore = (int)bs; *evalue = (double)totalBases * std::pow(2, -bs) * (double)qlen;
}

}  // namespace lmo
