// lmi_format.hpp — host-side reader/writer for the LexicMap on-disk index (.lmi directory).
//
// Product code (NOT the oracle): used by the index builder (lmi_build.cpp) and by the GPU image
// loader (engine.cu). Everything is big-endian, as in the reference (`var be = binary.BigEndian`).
//
// Format sources (reference file:line, relative to /root/reference/lexicmap/cmd):
//   kv-data  (.bin)      kv/kv-data.go:66-125 (layout), :261-304 (headers), :328-602 (per-mask records)
//   kv-index (.bin.idx)  kv/kv-data.go:566-599 (writer), :631-769 (reader)
//   VARINT-GB            util/varint-GB.go:28-44 (PutUint64s), :84-106 (Uint64s)
//   7-byte values        kv/kv-encoding.go:30-46
//   genomes.bin(.idx)    genome/genome.go:218-295 (records), :298-358 (index), :1427-1500 (2-bit packing)
//   info.toml            lib-index-build.go:1914-1932
//   genomes.map.bin      lib-index-build.go:1969-2017
//   seed value bits      lib-index-build.go:412-455
//   masks.bin            lexichash v0.5.5 (source absent from the reference tree): 32-byte header + 8 B/mask is
//                        size-verified only (demo/README.md:144); the header field layout below is OURS.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <stdexcept>
#include <algorithm>
#include <map>
#include <sys/stat.h>

namespace lmi {

// ---------------------------------------------------------------- constants (lib-index-build.go:412-455)
constexpr int BITS_BATCH_IDX = 17, BITS_GENOME_IDX = 17, BITS_POSITION = 28;
constexpr int BITS_NONE_IDX = 64 - BITS_BATCH_IDX - BITS_GENOME_IDX;  // 30
constexpr uint64_t MASK_NONE_IDX = (1ull << BITS_NONE_IDX) - 1;
constexpr uint64_t MASK_GENOME_IDX = (1ull << BITS_GENOME_IDX) - 1;

inline void die(const std::string& m) { throw std::runtime_error(m); }

// ---------------------------------------------------------------- big-endian helpers
inline void put_be(uint8_t* b, uint64_t v, int n) { for (int i = 0; i < n; i++) b[i] = (uint8_t)(v >> (8 * (n - 1 - i))); }
inline uint64_t get_be(const uint8_t* b, int n) { uint64_t v = 0; for (int i = 0; i < n; i++) v = v << 8 | b[i]; return v; }

struct FileW {
  FILE* f = nullptr; uint64_t n = 0;
  explicit FileW(const std::string& p) { f = fopen(p.c_str(), "wb"); if (!f) die("cannot create " + p); setvbuf(f, nullptr, _IOFBF, 1 << 20); }
  ~FileW() { if (f) fclose(f); }
  void w(const void* p, size_t len) { if (len && fwrite(p, 1, len, f) != len) die("write failed"); n += len; }
  void be(uint64_t v, int nb) { uint8_t b[8]; put_be(b, v, nb); w(b, nb); }
  void close() { if (f) { if (fclose(f)) die("close failed"); f = nullptr; } }
};

inline std::vector<uint8_t> read_file(const std::string& p) {
  FILE* f = fopen(p.c_str(), "rb"); if (!f) die("cannot open " + p);
  fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> b((size_t)sz);
  if (sz && fread(b.data(), 1, (size_t)sz, f) != (size_t)sz) { fclose(f); die("short read " + p); }
  fclose(f); return b;
}
inline bool file_exists(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0; }
inline void mkdir_p(const std::string& p) {
  std::string cur; for (size_t i = 0; i <= p.size(); i++) { if (i == p.size() || p[i] == '/') { if (!cur.empty()) mkdir(cur.c_str(), 0755); } if (i < p.size()) cur += p[i]; }
}

// ---------------------------------------------------------------- k-mer helpers
// 2-bit codes A0 C1 G2 T3, degenerate bases as genome.base2bit (genome/genome.go:1427-1444)
inline uint8_t base2bit(uint8_t c) {
  switch (c) {
    case 'C': case 'c': case 'B': case 'b': case 'S': case 's': case 'Y': case 'y': return 1;
    case 'G': case 'g': case 'K': case 'k': return 2;
    case 'T': case 't': case 'U': case 'u': return 3;
    default: return 0;
  }
}
inline uint64_t kmer_reverse(uint64_t c, int k) {  // kmers.MustReverse: reverse base order (no complement)
  uint64_t r = 0; for (int i = 0; i < k; i++) { r = r << 2 | (c & 3); c >>= 2; } return r;
}
inline uint64_t kmer_revcomp(uint64_t c, int k) {
  uint64_t r = 0; for (int i = 0; i < k; i++) { r = r << 2 | (3 - (c & 3)); c >>= 2; } return r;
}
inline uint64_t kmer_ns(uint64_t b, int k) { uint64_t c = b; for (int i = 1; i < k; i++) c = c << 2 | b; return c; }  // util.Ns
// util.IsLowComplexityDust (util/kmers.go:162-328): 3-mer windows i=0..k-2 of code>>(2i)&63, score = sum c(c-1)/2 > 50
inline bool is_low_complexity_dust(uint64_t code, int k) {
  uint8_t cnt[64]; memset(cnt, 0, sizeof cnt);
  for (int i = 0; i <= k - 2; i++) cnt[(code >> (2 * i)) & 63]++;
  uint16_t score = 0; for (int i = 0; i < 64; i++) { uint16_t c = cnt[i]; score += (uint16_t)((uint16_t)(c - 1) * c) >> 1; }
  return score > 50;
}
inline bool is_low_complexity(uint64_t kmer, int k) {  // lib-index-search.go:1222-1238 / lib-index-build.go:1033-1046
  uint64_t ttt = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
  return kmer == kmer_ns(1, k) || kmer == kmer_ns(2, k) || kmer == ttt || is_low_complexity_dust(kmer, k);
}

// ---------------------------------------------------------------- VARINT-GB (util/varint-GB.go)
inline int byte_len_u64(uint64_t v) { int n = 1; while (v >>= 8) n++; return n; }
inline int put_u64s(uint8_t* buf, uint64_t v1, uint64_t v2, uint8_t* ctrl) {
  int n = 0, b1 = byte_len_u64(v1), b2 = byte_len_u64(v2);
  for (int i = b1 - 1; i >= 0; i--) buf[n++] = (uint8_t)(v1 >> (8 * i));
  for (int i = b2 - 1; i >= 0; i--) buf[n++] = (uint8_t)(v2 >> (8 * i));
  *ctrl = (uint8_t)(((b1 - 1) << 3) | (b2 - 1));
  return n;
}
inline int get_u64s(uint8_t ctrl, const uint8_t* buf, uint64_t* v1, uint64_t* v2) {
  int b1 = ((ctrl >> 3) & 7) + 1, b2 = (ctrl & 7) + 1;
  *v1 = get_be(buf, b1); *v2 = get_be(buf + b1, b2); return b1 + b2;
}

// ---------------------------------------------------------------- kv-data writer
struct KvEntry { uint64_t kmer; std::vector<uint64_t> values; };

struct KvWriter {
  int k, mask_prefix, anchor_prefix; bool use7; FileW fd, fi; std::vector<uint64_t> p2o;
  KvWriter(const std::string& file, int k_, int mask_offset, int chunk_size, int mask_prefix_, int anchor_prefix_, bool use7bytes)
      : k(k_), mask_prefix(mask_prefix_), anchor_prefix(anchor_prefix_), use7(use7bytes), fd(file), fi(file + ".idx") {
    uint8_t cfg = use7 ? 1 : 0;
    fd.w(".kv-data", 8); uint8_t m1[8] = {1, 1, (uint8_t)k, cfg, 0, 0, 0, 0}; fd.w(m1, 8); fd.be(mask_offset, 8); fd.be(chunk_size, 8);
    fi.w(".kvindex", 8); uint8_t m2[8] = {1, 1, (uint8_t)k, (uint8_t)mask_prefix, (uint8_t)anchor_prefix, cfg, 0, 0}; fi.w(m2, 8); fi.be(mask_offset, 8); fi.be(chunk_size, 8);
    p2o.resize(2 + (2ull << (2 * anchor_prefix)));
  }
  uint64_t anchor_of(uint64_t kmer) const { return (kmer >> (2 * (k - mask_prefix - anchor_prefix))) & ((1ull << (2 * anchor_prefix)) - 1); }
  void put_values(const std::vector<uint64_t>& v) { for (uint64_t x : v) { if (use7) fd.be(x & 0x00ffffffffffffffull, 7); else fd.be(x, 8); } }
  // entries must be sorted by kmer, distinct (kv-data.go:328-602)
  void write_mask(const std::vector<KvEntry>& e) {
    size_t n = e.size();
    fd.be(n, 8);
    if (n == 0) { fi.be(0, 8); return; }
    std::fill(p2o.begin(), p2o.end(), 0);
    p2o[1] = fd.n << 1;
    bool first = true; uint64_t prefix_pre = 0, offset = 0; uint8_t buf[40], ctrl;
    size_t i = 0;
    for (; i + 1 < n; i += 2) {
      const KvEntry &a = e[i], &b = e[i + 1];
      uint64_t pre = anchor_of(a.kmer);
      if (first || pre != prefix_pre) { first = false; size_t j = (pre << 1) + 2; p2o[j] = a.kmer; p2o[j + 1] = fd.n << 1; prefix_pre = pre; }
      pre = anchor_of(b.kmer);
      if (pre != prefix_pre) { size_t j = (pre << 1) + 2; p2o[j] = b.kmer; p2o[j + 1] = (fd.n << 1) | 1; prefix_pre = pre; }
      int nb = put_u64s(buf + 1, a.kmer - offset, b.kmer - a.kmer, &ctrl);
      if ((n & 1) == 0 && i + 2 == n) ctrl |= 1 << 7;
      buf[0] = ctrl; int len = nb + 1;
      nb = put_u64s(buf + len + 1, a.values.size(), b.values.size(), &ctrl); buf[len] = ctrl; len += nb + 1;
      fd.w(buf, len); put_values(a.values); put_values(b.values);
      offset = b.kmer;
    }
    if (i < n) {  // last single one
      const KvEntry& a = e[i];
      uint64_t pre = anchor_of(a.kmer);
      if (first || pre != prefix_pre) { size_t j = (pre << 1) + 2; p2o[j] = a.kmer; p2o[j + 1] = fd.n << 1; }
      int nb = put_u64s(buf + 1, a.kmer - offset, 0, &ctrl); ctrl |= (1 << 7) | (1 << 6);
      buf[0] = ctrl; int len = nb + 1;
      nb = put_u64s(buf + len + 1, a.values.size(), 0, &ctrl); buf[len] = ctrl; len += nb + 1;
      fd.w(buf, len); put_values(a.values);
    }
    uint64_t nrec = 0; size_t E = p2o.size() >> 1;
    for (size_t j = 0; j < E; j++) if (p2o[2 * j + 1] > 0) nrec++;
    fi.be(nrec, 8); p2o[0] = nrec;
    for (size_t j = 0; j < E; j++) if (p2o[2 * j + 1] > 0) { fi.be(p2o[2 * j], 8); fi.be(p2o[2 * j + 1], 8); }
  }
  void close() { fd.close(); fi.close(); }
};

// ---------------------------------------------------------------- kv-data decoder (whole chunk → flat arrays)
struct KvMaskData {            // one mask bucket, decoded
  std::vector<uint64_t> keys;      // sorted distinct k-mers
  std::vector<uint32_t> val_off;   // CSR, size keys+1
  std::vector<uint64_t> vals;
  std::vector<uint64_t> rec_off;   // file offset of each 2-k-mer record (record r holds keys 2r, 2r+1)
};
struct KvChunk { int k = 0, mask_offset = 0, chunk_size = 0, mask_prefix = 0, anchor_prefix = 0; bool use7 = false; std::vector<KvMaskData> masks;
  // anchor table from the .idx file: per mask, 4^anchor_prefix entries = bucket-relative index of the first
  // key to scan from (0xFFFFFFFF = anchor absent). Inherits the "last run wins" behaviour of the writer.
  std::vector<std::vector<uint32_t>> anchor_start; };

inline KvChunk read_kv_chunk(const std::string& file) {
  KvChunk c; std::vector<uint8_t> d = read_file(file), x = read_file(file + ".idx");
  if (d.size() < 32 || memcmp(d.data(), ".kv-data", 8)) die("not a kv-data file: " + file);
  if (x.size() < 32 || memcmp(x.data(), ".kvindex", 8)) die("not a kv-index file: " + file + ".idx");
  if (d[8] != 1 || x[8] != 1) die("kv-data: version mismatch");
  c.k = d[10]; c.use7 = d[11] & 1; c.mask_offset = (int)get_be(&d[16], 8); c.chunk_size = (int)get_be(&d[24], 8);
  c.mask_prefix = x[11]; c.anchor_prefix = x[12];
  const int vb = c.use7 ? 7 : 8; size_t p = 32; c.masks.resize(c.chunk_size);
  for (int m = 0; m < c.chunk_size; m++) {
    KvMaskData& md = c.masks[m];
    uint64_t nk = get_be(&d[p], 8); p += 8;
    md.keys.reserve(nk); md.val_off.reserve(nk + 1); md.val_off.push_back(0);
    uint64_t prev = 0;
    while (md.keys.size() < nk) {
      md.rec_off.push_back(p);
      uint8_t ctrl = d[p++]; bool last = ctrl & 128, single = ctrl & 64; ctrl &= 63;
      uint64_t v1, v2, n1, n2; p += get_u64s(ctrl, &d[p], &v1, &v2);
      uint64_t k1 = prev + v1, k2 = k1 + v2; prev = k2;
      ctrl = d[p++]; p += get_u64s(ctrl, &d[p], &n1, &n2);
      md.keys.push_back(k1); for (uint64_t j = 0; j < n1; j++, p += vb) md.vals.push_back(get_be(&d[p], vb)); md.val_off.push_back((uint32_t)md.vals.size());
      if (!(last && single)) { md.keys.push_back(k2); for (uint64_t j = 0; j < n2; j++, p += vb) md.vals.push_back(get_be(&d[p], vb)); md.val_off.push_back((uint32_t)md.vals.size()); }
      if (p > d.size()) die("kv-data: broken file " + file);
    }
  }
  // anchor table
  size_t q = 32; const size_t NA = 1ull << (2 * c.anchor_prefix);
  const int sh = 2 * (c.k - c.mask_prefix - c.anchor_prefix);
  c.anchor_start.resize(c.chunk_size);
  for (int m = 0; m < c.chunk_size; m++) {
    uint64_t nrec = get_be(&x[q], 8); q += 8;
    std::vector<uint32_t>& as = c.anchor_start[m]; as.assign(NA, 0xFFFFFFFFu);
    const KvMaskData& md = c.masks[m];
    for (uint64_t r = 0; r < nrec; r++, q += 16) {
      if (r == 0) continue;  // (nRecords, offset of the first k-mer): informational
      uint64_t kmer = get_be(&x[q], 8), off = get_be(&x[q + 8], 8);
      size_t a = (kmer >> sh) & (NA - 1);
      bool second = off & 1; off >>= 1;
      auto it = std::lower_bound(md.rec_off.begin(), md.rec_off.end(), off);
      if (it == md.rec_off.end() || *it != off) die("kv-index: offset does not point at a record");
      as[a] = (uint32_t)(2 * (it - md.rec_off.begin()) + (second ? 1 : 0));
    }
  }
  return c;
}

// ---------------------------------------------------------------- genomes.bin
struct GenomeRec {
  std::string id; uint32_t genome_size = 0, concat_len = 0; std::vector<uint32_t> seq_sizes; std::vector<std::string> seq_ids;
  std::vector<uint8_t> twobit;  // first base in bits 7-6
};
inline std::vector<uint8_t> seq_to_2bit(const uint8_t* s, size_t n) {
  std::vector<uint8_t> b((n + 3) / 4, 0);
  for (size_t i = 0; i < n; i++) b[i >> 2] |= base2bit(s[i]) << (6 - 2 * (i & 3));
  return b;
}
struct GenomeWriter {
  FileW f; std::string path; uint32_t batch; std::vector<std::pair<uint64_t, uint32_t>> index;
  GenomeWriter(const std::string& p, uint32_t batch_) : f(p), path(p), batch(batch_) { f.w(".genomes", 8); uint8_t v[8] = {0, 1, 0, 0, 0, 0, 0, 0}; f.w(v, 8); }
  void write(const GenomeRec& g) {
    index.push_back({f.n, g.concat_len});
    f.be(g.id.size(), 2); f.w(g.id.data(), g.id.size());
    f.be(g.genome_size, 4); f.be(g.concat_len, 4); f.be(g.seq_sizes.size(), 4);
    for (size_t i = 0; i < g.seq_sizes.size(); i++) { f.be(g.seq_sizes[i], 4); f.be(g.seq_ids[i].size(), 2); f.w(g.seq_ids[i].data(), g.seq_ids[i].size()); }
    f.be(g.twobit.size(), 4); f.be(g.concat_len, 4); f.w(g.twobit.data(), g.twobit.size());
  }
  void close() {
    f.close(); FileW x(path + ".idx"); x.w(".genomei", 8); uint8_t v[8] = {0, 1, 0, 0, 0, 0, 0, 0}; x.w(v, 8);
    x.be(batch, 4); x.be(index.size(), 4); for (auto& e : index) { x.be(e.first, 8); x.be(e.second, 4); } x.close();
  }
};
inline std::vector<GenomeRec> read_genomes(const std::string& file) {
  std::vector<uint8_t> d = read_file(file), x = read_file(file + ".idx");
  if (d.size() < 16 || memcmp(d.data(), ".genomes", 8) || memcmp(x.data(), ".genomei", 8)) die("not a genome data file: " + file);
  uint32_t n = (uint32_t)get_be(&x[20], 4); std::vector<GenomeRec> out(n);
  for (uint32_t i = 0; i < n; i++) {
    size_t p = get_be(&x[24 + 12 * i], 8); GenomeRec& g = out[i];
    size_t l = get_be(&d[p], 2); p += 2; g.id.assign((const char*)&d[p], l); p += l;
    g.genome_size = (uint32_t)get_be(&d[p], 4); g.concat_len = (uint32_t)get_be(&d[p + 4], 4); uint32_t ns = (uint32_t)get_be(&d[p + 8], 4); p += 12;
    for (uint32_t s = 0; s < ns; s++) { g.seq_sizes.push_back((uint32_t)get_be(&d[p], 4)); l = get_be(&d[p + 4], 2); p += 6; g.seq_ids.emplace_back((const char*)&d[p], l); p += l; }
    size_t nb = get_be(&d[p], 4); p += 8; g.twobit.assign(d.begin() + p, d.begin() + p + nb);
  }
  return out;
}

// ---------------------------------------------------------------- masks.bin (header layout is ours; see top of file)
inline void write_masks(const std::string& file, const std::vector<uint64_t>& masks, int k, int64_t seed) {
  FileW f(file); f.w(".lexhash", 8); uint8_t v[8] = {0, 5, (uint8_t)k, 0, 0, 0, 0, 0}; f.w(v, 8); f.be(masks.size(), 8); f.be((uint64_t)seed, 8);
  for (uint64_t m : masks) f.be(m, 8); f.close();
}
inline std::vector<uint64_t> read_masks(const std::string& file, int* k) {
  std::vector<uint8_t> d = read_file(file); if (d.size() < 32) die("masks.bin too short");
  size_t n = (d.size() - 32) / 8; if (k) *k = d[10]; std::vector<uint64_t> m(n); for (size_t i = 0; i < n; i++) m[i] = get_be(&d[32 + 8 * i], 8); return m;
}

// ---------------------------------------------------------------- info.toml
struct IndexInfo {
  int main_version = 3, minor_version = 5, k = 31, masks = 20000; int64_t rand_seed = 1; int max_desert = 100, seed_dist_in_desert = 50;
  int chunks = 16, partitions = 4096, input_genomes = 0; int64_t input_bases = 0; int genomes = 0, genome_batch_size = 0, genome_batches = 1, contig_interval = 1000;
  bool soft_masking = false; int max_kmer_freq = 0;
};
inline void write_info(const std::string& file, const IndexInfo& i) {
  FILE* f = fopen(file.c_str(), "w"); if (!f) die("cannot create " + file);
  fprintf(f, "# Index format\nmain-version = %d\nminor-version = %d\n# LexicHash\nmax-K = %d\nmasks = %d\nrand-seed = %lld\n# Seed distance\nmax-seed-dist = %d\nseed-dist-in-desert = %d\n"
             "# Seeds (k-mer-value data) files\nchunks = %d\nindex-partitions = %d\n# Input genomes\ninput-genomes = %d\n# Input bases\ninput-bases = %lld\n# Genome data.\ngenomes = %d\n"
             "genome-batch-size = %d\ngenome-batches = %d\ncontig-interval = %d\nsoft-masking = %s\nmax-kmer-freq = %d\n",
          i.main_version, i.minor_version, i.k, i.masks, (long long)i.rand_seed, i.max_desert, i.seed_dist_in_desert, i.chunks, i.partitions, i.input_genomes,
          (long long)i.input_bases, i.genomes, i.genome_batch_size, i.genome_batches, i.contig_interval, i.soft_masking ? "true" : "false", i.max_kmer_freq);
  fclose(f);
}
inline IndexInfo read_info(const std::string& file) {
  IndexInfo i; FILE* f = fopen(file.c_str(), "r"); if (!f) die("cannot open " + file); char line[512];
  std::map<std::string, std::string> kv;
  while (fgets(line, sizeof line, f)) { char key[128], val[256]; if (line[0] == '#') continue; if (sscanf(line, " %127[^ =] = %255[^\n]", key, val) == 2) kv[key] = val; }
  fclose(f);
  auto I = [&](const char* k, long long d) { auto it = kv.find(k); return it == kv.end() ? d : atoll(it->second.c_str()); };
  i.main_version = (int)I("main-version", 3); i.minor_version = (int)I("minor-version", 5); i.k = (int)I("max-K", 31); i.masks = (int)I("masks", 0); i.rand_seed = I("rand-seed", 1);
  i.max_desert = (int)I("max-seed-dist", 100); i.seed_dist_in_desert = (int)I("seed-dist-in-desert", 50); i.chunks = (int)I("chunks", 1); i.partitions = (int)I("index-partitions", 4096);
  i.input_genomes = (int)I("input-genomes", 0); i.input_bases = I("input-bases", 0); i.genomes = (int)I("genomes", 0); i.genome_batch_size = (int)I("genome-batch-size", 0);
  i.genome_batches = (int)I("genome-batches", 1); i.contig_interval = (int)I("contig-interval", 1000); i.max_kmer_freq = (int)I("max-kmer-freq", 0);
  i.soft_masking = kv.count("soft-masking") && kv["soft-masking"].find("true") != std::string::npos;
  if (i.main_version != 3) die("index main-version must be 3");
  return i;
}

// ---------------------------------------------------------------- genomes.map.bin
inline void write_genome_map(const std::string& file, const std::vector<std::pair<std::string, uint64_t>>& m) {
  FileW f(file); for (auto& e : m) { f.be(e.first.size(), 2); f.w(e.first.data(), e.first.size()); f.be(e.second, 8); } f.close();
}
inline std::vector<std::pair<std::string, uint64_t>> read_genome_map(const std::string& file) {
  std::vector<uint8_t> d = read_file(file); std::vector<std::pair<std::string, uint64_t>> m; size_t p = 0;
  while (p + 2 <= d.size()) { size_t l = get_be(&d[p], 2); p += 2; std::string id((const char*)&d[p], l); p += l; m.push_back({id, get_be(&d[p], 8)}); p += 8; }
  return m;
}

// genomes.chunks.bin (lib-index-build.go:1795-1812, readGenomeChunksLists :2193-2246): per split genome a big-endian u64 count followed by
// that many batch+genome indexes. A missing or empty file means no genome was split.
inline std::vector<std::vector<uint64_t>> read_genome_chunks(const std::string& file) {
  std::vector<std::vector<uint64_t>> out; FILE* f = fopen(file.c_str(), "rb"); if (!f) return out; fclose(f);
  std::vector<uint8_t> d = read_file(file); size_t p = 0;
  while (p + 8 <= d.size()) { uint64_t n = get_be(&d[p], 8); p += 8; if (p + 8 * n > d.size()) die("broken genome chunk file"); std::vector<uint64_t> l; for (uint64_t i = 0; i < n; i++, p += 8) l.push_back(get_be(&d[p], 8)); out.push_back(std::move(l)); }
  return out;
}

inline std::string chunk_file(const std::string& dir, int i) { char b[64]; snprintf(b, sizeof b, "/seeds/chunk_%03d.bin", i); return dir + b; }
inline std::string batch_dir(const std::string& dir, int b) { char s[64]; snprintf(s, sizeof s, "/genomes/batch_%04d", b); return dir + s; }

}  // namespace lmi
