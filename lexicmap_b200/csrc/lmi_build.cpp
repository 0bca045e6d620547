// lmi_build.cpp — host-side tools: minimal format-exact LexicMap index writer + synthetic data generators.
//
// Why it exists: there is no Go toolchain here, so nobody else can produce a `.lmi` index.
// Scope of the writer (reference: lexicmap/cmd/lib-index-build.go): contig concatenation with
// `contig-interval` A's (:924,1662-1667); first-round LexicHash capture — one k-mer per mask with all tied
// positions, both strands (:1028); DUST / homopolymer filter (:1033-1046); value packing (:708-716); every
// captured k-mer stored a second time base-reversed under argmin_j(mask_j XOR rev) with the reversed flag
// (:776-890); per-chunk kv-data + anchor index (:1856-1901); genomes.bin, genomes.map.bin, masks.bin, info.toml.
// Seed-desert filling (:1086-1500) is restated below and ON by default, as in `lexicmap index`
// (--no-fill-deserts writes max-seed-dist = 0 in info.toml to mark its absence).
//
// LexicHash masks: the reference generates them with lexichash.NewWithSeed(k, m, seed, p) (Go math/rand; not
// reproducible here). We generate our own mask set with the same structural properties documented in
// docs/content/usage/utils/masks.md:67-111: sorted ascending, every p-base prefix present (p = floor(log4 m)),
// the surplus masks share a p-prefix with another mask but differ at base p (distinct (p+1)-prefixes).
#include "lmi_format.hpp"
#include <omp.h>
#include <parallel/algorithm>
#include <zlib.h>
#include <cmath>
#include <unordered_map>

using namespace lmi;

static inline uint64_t splitmix64(uint64_t& s) { uint64_t z = (s += 0x9E3779B97F4A7C15ull); z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }

// ------------------------------------------------------------------ masks
std::vector<uint64_t> gen_masks(int k, int m, int64_t seed) {
  int p = std::max((int)(std::log2((double)m) / 2), 1);  // lib-index-search.go:467
  uint64_t np = 1ull << (2 * p); if ((uint64_t)m < np) { p = 0; while ((1ull << (2 * (p + 1))) <= (uint64_t)m) p++; np = 1ull << (2 * p); }
  uint64_t s = (uint64_t)seed * 0x2545F4914F6CDD1Dull + 12345; std::vector<uint64_t> masks; masks.reserve(m);
  const int rest = 2 * (k - p); const uint64_t rmask = (rest >= 64) ? ~0ull : ((1ull << rest) - 1);
  auto draw = [&](uint64_t prefix, int fix_next /* -1 none, else value of base p */) {
    for (;;) {
      uint64_t r = splitmix64(s) & rmask; uint64_t c = (p ? prefix << rest : 0) | r;
      if (fix_next >= 0) c = (c & ~(3ull << (rest - 2))) | ((uint64_t)fix_next << (rest - 2));
      if (!is_low_complexity(c, k) && c != 0) return c;
    }
  };
  for (uint64_t pre = 0; pre < np; pre++) masks.push_back(draw(pre, -1));
  // surplus masks: pick prefixes round-robin in a scrambled order, base p differs from the existing mask's
  uint64_t extra = (uint64_t)m - np, step = 0x9E3779B1ull % np | 1, cur = splitmix64(s) % np;
  std::vector<uint8_t> used(np, 0);
  for (uint64_t e = 0; e < extra; e++) {
    while (used[cur] >= 3) cur = (cur + 1) % np;
    uint64_t base = masks[cur]; int b0 = (int)((base >> (rest - 2)) & 3); int nb = (b0 + 1 + used[cur]) & 3;
    used[cur]++; masks.push_back(draw(cur, nb)); cur = (cur + step) % np;
  }
  std::sort(masks.begin(), masks.end());
  masks.erase(std::unique(masks.begin(), masks.end()), masks.end());
  if ((int)masks.size() != m) die("mask generation produced duplicates");
  return masks;
}

// ------------------------------------------------------------------ LexicHash capture of one sequence (global argmin per mask)
struct Capture { std::vector<uint64_t> kmer; std::vector<std::vector<uint32_t>> locs; };  // loc = pos<<1 | strand

// masks sorted ascending. Definition: every mask captures argmin over all k-mers of both strands of
// (kmer XOR mask); all positions where that k-mer occurs are kept. K-mers overlapping a skip region are ignored.
static void capture_sequence(const uint8_t* seq /*2-bit codes 0..3, one per byte*/, size_t n, int k, const std::vector<uint64_t>& masks,
                             const std::vector<std::pair<int64_t, int64_t>>& skip, Capture& out) {
  const size_t m = masks.size(); out.kmer.assign(m, 0); out.locs.assign(m, {});
  if (n < (size_t)k) return;
  // prefix index over masks: p = largest prefix length such that 4^p <= m
  int p = 0; while ((1ull << (2 * (p + 1))) <= m) p++;
  const uint64_t np = 1ull << (2 * p); std::vector<uint32_t> pstart(np + 1, 0);
  for (size_t j = 0; j < m; j++) pstart[(masks[j] >> (2 * (k - p))) + 1]++;
  for (uint64_t i = 0; i < np; i++) pstart[i + 1] += pstart[i];
  std::vector<uint64_t> best(m, ~0ull);
  const uint64_t kmask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
  uint64_t fw = 0, rc = 0; size_t si = 0; int64_t blocked_until = -1;  // k-mer start positions <= blocked_until are skipped
  std::vector<uint64_t> all_f, all_r; std::vector<uint32_t> all_pos; bool need_fallback_store = true;
  all_f.reserve(n); all_r.reserve(n); all_pos.reserve(n);
  for (size_t i = 0; i < n; i++) {
    fw = ((fw << 2) | seq[i]) & kmask; rc = (rc >> 2) | ((uint64_t)(3 - seq[i]) << (2 * (k - 1)));
    if (i + 1 < (size_t)k) continue;
    int64_t pos = (int64_t)i + 1 - k;
    while (si < skip.size() && skip[si].second < pos) si++;
    if (si < skip.size() && skip[si].first <= pos + k - 1) continue;  // window [pos,pos+k-1] overlaps skip region
    (void)blocked_until;
    if (need_fallback_store) { all_f.push_back(fw); all_r.push_back(rc); all_pos.push_back((uint32_t)pos); }
    for (int st = 0; st < 2; st++) {
      uint64_t km = st ? rc : fw; uint64_t pre = km >> (2 * (k - p)); uint32_t loc = (uint32_t)(pos << 1 | st);
      for (uint32_t j = pstart[pre]; j < pstart[pre + 1]; j++) {
        uint64_t h = km ^ masks[j];
        if (h < best[j]) { best[j] = h; out.kmer[j] = km; out.locs[j].clear(); out.locs[j].push_back(loc); }
        else if (h == best[j]) out.locs[j].push_back(loc);
      }
    }
  }
  // masks whose p-prefix is shared by no k-mer: exhaustive scan (exact global argmin)
  for (size_t j = 0; j < m; j++) if (best[j] == ~0ull && !all_f.empty()) {
    for (size_t t = 0; t < all_f.size(); t++) for (int st = 0; st < 2; st++) {
      uint64_t km = st ? all_r[t] : all_f[t]; uint64_t h = km ^ masks[j]; uint32_t loc = all_pos[t] << 1 | st;
      if (h < best[j]) { best[j] = h; out.kmer[j] = km; out.locs[j].clear(); out.locs[j].push_back(loc); }
      else if (h == best[j]) out.locs[j].push_back(loc);
    }
  }
}

// argmin_j (masks[j] XOR x) over sorted masks (lexichash MaskKmer + the loop at lib-index-build.go:813-821)
static uint32_t xor_argmin_mask(const std::vector<uint64_t>& masks, uint64_t x) {
  size_t lo = 0, hi = masks.size();
  for (int bit = 63; bit >= 0 && hi - lo > 1; bit--) {
    uint64_t b = 1ull << bit;
    // masks[lo..hi) share all bits above `bit`; find split
    size_t a = lo, z = hi; while (a < z) { size_t mid = (a + z) / 2; if (masks[mid] & b) z = mid; else a = mid + 1; }
    if (a == lo || a == hi) continue;  // all same at this bit
    if (x & b) lo = a; else hi = a;
  }
  return (uint32_t)lo;
}

// ------------------------------------------------------------------ input genomes
struct InGenome { std::string id; std::vector<std::string> seq_ids; std::vector<std::string> seqs; };

static std::string gz_slurp(const std::string& path) {
  gzFile g = gzopen(path.c_str(), "rb"); if (!g) die("cannot open " + path); std::string s; char buf[1 << 16]; int n;
  while ((n = gzread(g, buf, sizeof buf)) > 0) s.append(buf, n); gzclose(g); return s;
}
static void parse_fasta(const std::string& txt, std::vector<std::string>& ids, std::vector<std::string>& seqs) {
  size_t p = 0, n = txt.size();
  while (p < n) {
    size_t e = txt.find('\n', p); if (e == std::string::npos) e = n;
    if (txt[p] == '>') { size_t q = p + 1; while (q < e && txt[q] != ' ' && txt[q] != '\t' && txt[q] != '\r') q++; ids.emplace_back(txt.substr(p + 1, q - p - 1)); seqs.emplace_back(); }
    else if (!seqs.empty()) { for (size_t q = p; q < e; q++) { char c = txt[q]; if (c != '\r' && c != ' ') seqs.back().push_back(c); } }
    p = e + 1;
  }
}

// synthetic genomes: F families x S members; ancestor uniform ACGT of length G; member = ancestor with
// per-base substitution rate u~U(0,0.10) and indel rate u/10 (geometric lengths, mean 2), split into 1..max_contigs contigs.
static InGenome synth_genome(int fam, int mem, int S, int G, uint64_t seed, int max_contigs) {
  static const char B[] = "ACGT"; uint64_t sa = seed + 1000003ull * (uint64_t)fam; std::string anc(G, 'A');
  for (int i = 0; i < G; i += 32) { uint64_t r = splitmix64(sa); for (int j = 0; j < 32 && i + j < G; j++) { anc[i + j] = B[r & 3]; r >>= 2; } }
  int gid = fam * S + mem; uint64_t s = seed + 7919ull * (uint64_t)(gid + 1) + 0xABCDEFull;
  double u = (mem == 0) ? 0.0 : (double)(splitmix64(s) >> 11) / 9007199254740992.0 * 0.10; double ind = u / 10;
  std::string g; g.reserve(G + G / 50);
  const uint64_t tu = (uint64_t)(u * 4294967296.0), ti = (uint64_t)(ind * 4294967296.0);
  for (int i = 0; i < G; i++) {
    uint64_t r = splitmix64(s); uint32_t a = (uint32_t)r, b = (uint32_t)(r >> 32);
    if (b < ti) {  // indel
      uint64_t r2 = splitmix64(s); int len = 1; while ((r2 & 1) && len < 16) { len++; r2 >>= 1; }
      if (r2 & 2) { i += len - 1; continue; }                       // deletion of len bases
      for (int j = 0; j < len; j++) { g.push_back(B[(r2 >> (8 + 2 * j)) & 3]); }  // insertion before base i
    }
    char c = anc[i]; if (a < tu) { int x = (int)((r >> 20) % 3); const char* alt = "ACGT"; int ci = (c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3); c = alt[(ci + 1 + x) & 3]; }
    g.push_back(c);
  }
  InGenome out; char nm[64]; snprintf(nm, sizeof nm, "SYN_%05d_%03d", fam, mem); out.id = nm;
  int nc = 1 + (int)(splitmix64(s) % (uint64_t)std::max(1, max_contigs)); size_t L = g.size(); if ((size_t)nc * 2000 > L) nc = (int)std::max<size_t>(1, L / 2000);
  std::vector<size_t> cuts; for (int c = 1; c < nc; c++) cuts.push_back(1000 + splitmix64(s) % (L - 2000)); cuts.push_back(0); cuts.push_back(L); std::sort(cuts.begin(), cuts.end());
  cuts.erase(std::unique(cuts.begin(), cuts.end()), cuts.end());
  for (size_t c = 0; c + 1 < cuts.size(); c++) { if (cuts[c + 1] - cuts[c] < 100 && c + 2 < cuts.size()) { cuts.erase(cuts.begin() + c + 1); c--; continue; } }
  for (size_t c = 0; c + 1 < cuts.size(); c++) { char sn[96]; snprintf(sn, sizeof sn, "%s.c%zu", nm, c + 1); out.seq_ids.push_back(sn); out.seqs.push_back(g.substr(cuts[c], cuts[c + 1] - cuts[c])); }
  return out;
}

// ------------------------------------------------------------------ the builder
struct Tuple { uint32_t dmask; uint32_t src; uint64_t kmer; uint64_t value; };  // src: bit31 = reversed, low bits = source mask
static inline bool tuple_less(const Tuple& a, const Tuple& b) {
  if (a.dmask != b.dmask) return a.dmask < b.dmask; if (a.kmer != b.kmer) return a.kmer < b.kmer;
  uint64_t ga = a.value >> 30, gb = b.value >> 30; if (ga != gb) return ga < gb;
  if (a.src != b.src) return a.src < b.src; return a.value < b.value;
}

// ------------------------------------------------------------------ seed-desert filling (lib-index-build.go:1086-1413)
// After the first round (one captured k-mer per mask, all its positions) consecutive seeds may be far apart. Every gap of >= max_desert
// bases is walked in steps of seed_dist: around each step (seed_dist/2 bases upstream, then downstream) the first position is taken whose
// k-mer (either strand) is captured by some mask when the REGION (gap +- 1000 bp) is masked on its own without the shorter-prefix fallback
// (lh.MaskKnownDistinctPrefixes(region, nil, false), :1199), is not low-complexity and does not touch a contig interval / gap region.
struct ExtraSeed { uint32_t mask; uint64_t kmer; uint32_t loc; };   // loc = pos<<1 | strand
struct DesertScratch { std::vector<uint64_t> best; std::vector<uint32_t> stamp, touched; uint32_t epoch = 0; std::vector<uint64_t> kl; std::vector<int32_t> l2m, l2mrc; };
static void fill_deserts(const uint8_t* seq, size_t n, int k, const std::vector<uint64_t>& masks, const std::vector<uint32_t>& pstart, int p,
                         const std::vector<std::pair<int64_t, int64_t>>& skip, const Capture& cap, int maxDesert, int seedDist, std::vector<ExtraSeed>& out, DesertScratch& S) {
  if (n < (size_t)k) return; const int seedPosR = seedDist / 2; const size_t m = masks.size(); const uint64_t kmask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
  if (S.best.size() != m) { S.best.assign(m, 0); S.stamp.assign(m, 0); S.epoch = 0; }
  std::vector<uint32_t> locs; for (size_t j = 0; j < m; j++) if (!cap.locs[j].empty() && !is_low_complexity(cap.kmer[j], k)) for (uint32_t l : cap.locs[j]) locs.push_back(l);
  std::sort(locs.begin(), locs.end()); locs.push_back((uint32_t)(n - k) << 1);   // pseudo position at the end (:1139)
  auto in_itree = [&](int64_t pos) { for (const auto& r : skip) { if (r.first - k + 1 > pos) break; if (pos <= r.second) return true; } return false; };   // :978, :1001 (regions widened by k-1 to the left)
  const uint64_t ccc = kmer_ns(1, k), ggg = kmer_ns(2, k), ttt = kmask;
  auto usable = [&](uint64_t km) { return km != 0 && km != ccc && km != ggg && km != ttt && !is_low_complexity_dust(km, k); };
  uint32_t pre = 0;
  for (uint32_t pos2str : locs) { const uint32_t pos = pos2str >> 1, d = pos - pre; if (d < (uint32_t)maxDesert) { pre = pos; continue; }
    int start = (int)pre - 1000, posOfPre = 1000; if (start < 0) { posOfPre += start; start = 0; } int end = (int)pos + 1000 + k; if (end > (int)n) end = (int)n;
    const int posOfCur = posOfPre + (int)d, nk = end - start - k + 1; if (nk <= 0) { pre = pos; continue; }
    // k-mers of the region, both strands (kmerList, :1184-1192)
    S.kl.resize(2 * (size_t)nk); { uint64_t fw = 0, rc = 0; for (int i = start; i < end; i++) { fw = ((fw << 2) | seq[i]) & kmask; rc = (rc >> 2) | ((uint64_t)(3 - seq[i]) << (2 * (k - 1))); if (i - start + 1 >= k) { int q = i - start + 1 - k; S.kl[2 * q] = fw; S.kl[2 * q + 1] = rc; } } }
    // region masking without fallback: per mask the XOR-argmin over the region k-mers sharing its p-base prefix; loc2maskidx = the LAST mask capturing a position (:1220-1233)
    S.epoch++; S.touched.clear();
    for (int q = 0; q < nk; q++) for (int st = 0; st < 2; st++) { uint64_t km = S.kl[2 * q + st]; uint64_t pre_ = km >> (2 * (k - p));
      for (uint32_t j = pstart[pre_]; j < pstart[pre_ + 1]; j++) { uint64_t h = km ^ masks[j]; if (S.stamp[j] != S.epoch) { S.stamp[j] = S.epoch; S.best[j] = h; S.touched.push_back(j); } else if (h < S.best[j]) S.best[j] = h; } }
    S.l2m.assign(nk, -1); S.l2mrc.assign(nk, -1);
    for (int q = 0; q < nk; q++) for (int st = 0; st < 2; st++) { uint64_t km = S.kl[2 * q + st]; uint64_t pre_ = km >> (2 * (k - p));
      for (uint32_t j = pstart[pre_]; j < pstart[pre_ + 1]; j++) if ((km ^ masks[j]) == S.best[j]) (st ? S.l2mrc : S.l2m)[q] = (int32_t)j; }   // ascending j: the last one stays
    auto probe = [&](int j, int& im, uint64_t& km, uint32_t& kpos) -> bool { if (j < 0 || j >= nk || in_itree((int64_t)start + j)) return false;
      uint64_t a = S.kl[2 * j]; if (usable(a) && S.l2m[j] >= 0) { im = S.l2m[j]; km = a; kpos = (uint32_t)(start + j) << 1; return true; }
      uint64_t b = S.kl[2 * j + 1]; if (usable(b) && S.l2mrc[j] >= 0) { im = S.l2mrc[j]; km = b; kpos = ((uint32_t)(start + j) << 1) | 1u; return true; } return false; };
    int j = posOfPre + seedDist;
    for (;;) { if (j >= posOfCur) break;
      int dstart = j + 1, uend = j - seedPosR; bool ok = false; int im = -1; uint64_t km = 0; uint32_t kpos = 0;
      for (; j > uend; j--) if (probe(j, im, km, kpos)) { ok = true; break; }                       // upstream scan (:1241-1296)
      if (ok) { out.push_back({(uint32_t)im, km, kpos}); j += seedDist; continue; }
      if (dstart >= posOfCur) break;
      int dend = dstart + seedPosR; if (dend >= posOfCur) dend = posOfCur - 1;
      for (j = dstart; j < dend; j++) if (probe(j, im, km, kpos)) { ok = true; break; }               // downstream scan (:1322-1372)
      if (ok) { out.push_back({(uint32_t)im, km, kpos}); j += seedDist; continue; }
      j += seedDist; }                                                                                // could not fill here (gap, interval, repeats), :1392-1401
    pre = pos; }
}

struct BuildOpts { int k = 31, masks = 20000, chunks = 16, partitions = 4096, contig_interval = 1000; int64_t seed = 1; int threads = 0;
  bool fill_deserts = true; int max_desert = 100, seed_dist = 50; int64_t max_genome = 15000000; int batch_size = 0; };   // -g/--max-genome (index.go:146), -b/--batch-size (0 = all genomes in one batch)
  //   // on by default like `lexicmap index` (index.go:115-119,210-211); --no-fill-deserts, -D/--seed-max-desert, -d/--seed-in-desert-dist (index.go:582-586)

template <class GetGenome>
static void build_index(const std::string& out, size_t n_genomes, GetGenome get, const BuildOpts& o) {
  if (n_genomes == 0) die("no genomes");
  const int k = o.k; mkdir_p(out); mkdir_p(out + "/seeds");
  std::vector<uint64_t> masks = gen_masks(k, o.masks, o.seed); write_masks(out + "/masks.bin", masks, k, o.seed);
  const int mask_prefix = std::max((int)(std::log2((double)o.masks) / 2), 1), anchor_prefix = std::max((int)(std::log2((double)o.partitions) / 2), 1);
  int nt = o.threads > 0 ? o.threads : omp_get_max_threads();
  int mask_p = 0; while ((1ull << (2 * (mask_p + 1))) <= masks.size()) mask_p++;   // same prefix directory as capture_sequence
  std::vector<uint32_t> mask_pstart(((size_t)1 << (2 * mask_p)) + 1, 0); for (uint64_t mk : masks) mask_pstart[(mk >> (2 * (k - mask_p))) + 1]++; for (size_t i = 0; i + 1 < mask_pstart.size(); i++) mask_pstart[i + 1] += mask_pstart[i];
  // Genome units: input genome gi is unit gi; a genome whose concatenated contigs exceed --max-genome is split at contig boundaries
  // (lib-index-build.go:1573-1660) and its further chunks become extra units (provisional serial >= n_genomes, renumbered below in (gi, chunk) order).
  // Unit u lives in genome batch u / batch_size as genome u % batch_size (batches of -b/--batch-size genomes, lib-index-build.go:560-720 + merge).
  struct Extra { size_t gi; int chunk; uint32_t prov; GenomeRec rec; };
  std::vector<std::vector<Tuple>> tl(nt); std::vector<GenomeRec> recs(n_genomes); std::vector<Extra> extras; std::vector<char> dropped(n_genomes, 0); int64_t total_bases = 0; long long n_extra = 0; uint32_t next_prov = (uint32_t)n_genomes;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt) reduction(+ : total_bases, n_extra)
  for (size_t gi = 0; gi < n_genomes; gi++) {
    InGenome g = get(gi);
    std::vector<std::pair<size_t, size_t>> chunks; { size_t c0 = 0; int64_t sz = 0; bool too_big = false;
      for (size_t c = 0; c < g.seqs.size(); c++) { int64_t L = (int64_t)g.seqs[c].size(); if (o.max_genome > 0 && L > o.max_genome) { too_big = true; break; }    // a single sequence above the limit: the genome is skipped (:1596-1612)
        if (o.max_genome > 0 && c > c0 && sz + L > o.max_genome) { chunks.push_back({c0, c}); c0 = c; sz = 0; }
        sz += L + (c > c0 ? o.contig_interval : 0); }
      if (too_big || g.seqs.empty()) { dropped[gi] = 1; fprintf(stderr, "[lmi-build] skipping %s: %s\n", g.id.c_str(), too_big ? "a sequence is larger than --max-genome" : "no valid sequences"); continue; }
      chunks.push_back({c0, g.seqs.size()}); }
    for (size_t ch = 0; ch < chunks.size(); ch++) {
      GenomeRec rl; GenomeRec& r = ch == 0 ? recs[gi] : rl; uint32_t unit = (uint32_t)gi;
      if (ch > 0) {
#pragma omp critical(lmi_extra)
        unit = next_prov++; }
      r.id = g.id; std::vector<uint8_t> codes; std::vector<std::pair<int64_t, int64_t>> skip;
      for (size_t c = chunks[ch].first; c < chunks[ch].second; c++) {
        if (c > chunks[ch].first) { skip.push_back({(int64_t)codes.size(), (int64_t)codes.size() + o.contig_interval - 1}); codes.insert(codes.end(), o.contig_interval, 0); }
        const std::string& s = g.seqs[c]; size_t b0 = codes.size(); codes.resize(b0 + s.size());
        // gap regions: runs of >= 5 N's are skip regions (lib-gaps.go; lib-index-build.go:992-1016)
        size_t run = 0; for (size_t i = 0; i < s.size(); i++) { codes[b0 + i] = base2bit((uint8_t)s[i]); bool isn = (s[i] == 'N' || s[i] == 'n');
          if (isn) run++; if ((!isn || i + 1 == s.size()) && run) { size_t end = isn ? i + 1 : i; if (run >= 5) skip.push_back({(int64_t)(b0 + end - run), (int64_t)(b0 + end - 1)}); run = 0; } }
        r.seq_sizes.push_back((uint32_t)s.size()); r.seq_ids.push_back(g.seq_ids[c]); r.genome_size += (uint32_t)s.size();
      }
      std::sort(skip.begin(), skip.end());
      if (codes.size() >= (1u << BITS_POSITION)) die("genome too large: " + g.id);
      r.concat_len = (uint32_t)codes.size(); total_bases += r.genome_size;
      r.twobit.assign((codes.size() + 3) / 4, 0); for (size_t i = 0; i < codes.size(); i++) r.twobit[i >> 2] |= codes[i] << (6 - 2 * (i & 3));
      Capture cap; capture_sequence(codes.data(), codes.size(), k, masks, skip, cap);
      std::vector<Tuple>& T = tl[omp_get_thread_num()]; const uint64_t gshift = (uint64_t)unit << BITS_NONE_IDX;  // unit serial for now; batch | index after renumbering
      if (o.fill_deserts) { static thread_local DesertScratch scratch; std::vector<ExtraSeed> ex; fill_deserts(codes.data(), codes.size(), k, masks, mask_pstart, mask_p, skip, cap, o.max_desert, o.seed_dist, ex, scratch);
        for (const ExtraSeed& e : ex) { uint64_t rv = kmer_reverse(e.kmer, k); uint32_t dj = xor_argmin_mask(masks, rv);   // extra k-mers enter like captured ones, incl. the reversed copy (:726-760, :845-890)
          T.push_back({e.mask, e.mask, e.kmer, gshift | (((uint64_t)e.loc << 1) & MASK_NONE_IDX)});
          T.push_back({dj, e.mask | 0x80000000u, rv, gshift | ((((uint64_t)e.loc << 1) | 1) & MASK_NONE_IDX)}); }
        n_extra += ex.size(); }
      for (size_t j = 0; j < masks.size(); j++) {
        uint64_t km = cap.kmer[j]; if (cap.locs[j].empty() || is_low_complexity(km, k)) continue;
        uint64_t rv = kmer_reverse(km, k); uint32_t dj = xor_argmin_mask(masks, rv);
        for (uint32_t loc : cap.locs[j]) {
          T.push_back({(uint32_t)j, (uint32_t)j, km, gshift | (((uint64_t)loc << 1) & MASK_NONE_IDX)});
          T.push_back({dj, (uint32_t)j | 0x80000000u, rv, gshift | ((((uint64_t)loc << 1) | 1) & MASK_NONE_IDX)});
        }
      }
      if (ch > 0) {
#pragma omp critical(lmi_extra)
        extras.push_back(Extra{gi, (int)ch, unit, std::move(rl)}); }
    }
  }
  // final unit numbering: kept input genomes in input order, then the extra chunks in (genome, chunk) order
  std::sort(extras.begin(), extras.end(), [](const Extra& a, const Extra& b) { return a.gi != b.gi ? a.gi < b.gi : a.chunk < b.chunk; });
  std::vector<uint32_t> unit_of(next_prov, 0xFFFFFFFFu); std::vector<GenomeRec*> units; for (size_t gi = 0; gi < n_genomes; gi++) if (!dropped[gi]) { unit_of[gi] = (uint32_t)units.size(); units.push_back(&recs[gi]); }
  for (Extra& e : extras) { unit_of[e.prov] = (uint32_t)units.size(); units.push_back(&e.rec); }
  if (units.empty()) die("no genomes left to index");
  const size_t NU = units.size(); const size_t bsz = o.batch_size > 0 ? (size_t)o.batch_size : NU; if (bsz > (1u << BITS_GENOME_IDX)) die("more than 131072 genomes per batch: use --batch-size"); const size_t nb = (NU + bsz - 1) / bsz; if (nb > (1u << BITS_BATCH_IDX)) die("too many genome batches");
  auto bgi_of = [&](uint32_t unit) { return ((uint64_t)(unit / bsz) << BITS_GENOME_IDX) | (uint64_t)(unit % bsz); };
  { bool ident = extras.empty() && nb == 1; for (size_t gi = 0; gi < n_genomes && ident; gi++) if (dropped[gi]) ident = false;
    if (!ident) {
#pragma omp parallel for schedule(static) num_threads(nt)
      for (int t = 0; t < nt; t++) for (Tuple& x : tl[t]) x.value = (bgi_of(unit_of[x.value >> BITS_NONE_IDX]) << BITS_NONE_IDX) | (x.value & MASK_NONE_IDX); } }
  // genomes.bin per batch + map + chunk lists
  { std::vector<std::pair<std::string, uint64_t>> gm;
    for (size_t b = 0; b < nb; b++) { mkdir_p(batch_dir(out, (int)b)); GenomeWriter gw(batch_dir(out, (int)b) + "/genomes.bin", (int)b);
      for (size_t u = b * bsz; u < std::min(NU, (b + 1) * bsz); u++) { gw.write(*units[u]); gm.push_back({units[u]->id, bgi_of((uint32_t)u)}); units[u]->twobit.clear(); units[u]->twobit.shrink_to_fit(); } gw.close(); }
    write_genome_map(out + "/genomes.map.bin", gm);
    FileW gc(out + "/genomes.chunks.bin");   // per split genome: #chunks, then the batch+genome index of every chunk (big endian u64s, :1795-1812)
    for (size_t x = 0; x < extras.size();) { size_t y = x; while (y < extras.size() && extras[y].gi == extras[x].gi) y++; gc.be((uint64_t)(y - x + 1), 8); gc.be(bgi_of(unit_of[extras[x].gi]), 8); for (size_t z = x; z < y; z++) gc.be(bgi_of(unit_of[extras[z].prov]), 8); x = y; }
    gc.close(); }
  const size_t n_units = NU, n_batches = nb; size_t n_input = 0; for (size_t gi = 0; gi < n_genomes; gi++) if (!dropped[gi]) n_input++;
  // seeds
  std::vector<Tuple> all; { size_t tot = 0; for (auto& v : tl) tot += v.size(); all.reserve(tot); for (auto& v : tl) { all.insert(all.end(), v.begin(), v.end()); std::vector<Tuple>().swap(v); } }
  __gnu_parallel::sort(all.begin(), all.end(), tuple_less);
  std::vector<size_t> mstart(o.masks + 1, 0); for (const Tuple& t : all) mstart[t.dmask + 1]++; for (int j = 0; j < o.masks; j++) mstart[j + 1] += mstart[j];
  const int chunk_size = (o.masks + o.chunks - 1) / o.chunks;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt)
  for (int c = 0; c < o.chunks; c++) {
    int begin = c * chunk_size, end = std::min(begin + chunk_size, o.masks); if (begin >= end) continue;
    KvWriter w(chunk_file(out, c), k, begin, end - begin, mask_prefix, anchor_prefix, n_batches <= 512 /* 7-byte values, kv-data.go:126-137 */);
    std::vector<KvEntry> ent;
    for (int j = begin; j < end; j++) {
      ent.clear();
      for (size_t t = mstart[j]; t < mstart[j + 1]; t++) { if (ent.empty() || ent.back().kmer != all[t].kmer) ent.push_back({all[t].kmer, {}}); ent.back().values.push_back(all[t].value); }
      w.write_mask(ent);
    }
    w.close();
  }
  IndexInfo info; info.k = k; info.masks = o.masks; info.rand_seed = o.seed; info.max_desert = o.fill_deserts ? o.max_desert : 0; info.seed_dist_in_desert = o.fill_deserts ? o.seed_dist : 0; info.chunks = o.chunks; info.partitions = o.partitions;
  info.input_genomes = (int)n_input; info.input_bases = total_bases; info.genomes = (int)n_units; info.genome_batch_size = (int)bsz; info.genome_batches = (int)n_batches; info.contig_interval = o.contig_interval;
  write_info(out + "/info.toml", info);
  fprintf(stderr, "[lmi-build] %zu genomes (%zu units in %zu batches), %lld bases, %zu seed values (%lld from desert filling), %d masks -> %s\n", n_input, n_units, n_batches, (long long)total_bases, all.size(), 2 * n_extra, o.masks, out.c_str());
}

// ------------------------------------------------------------------ synthetic queries from an index's genomes, or straight from the synthetic collection
// (SURVEY.md §8d): uniform random genome, uniform start, `len` bases, 0..max_sub substitutions + 0..max_indel indels, 50 % reverse-complemented.
// `from_index` empty: genomes [glo, ghi) of the collection --synth F,S,G,seed,mc are regenerated on the fly (a shard's queries without its index).
static void synth_queries(const std::string& index, int n, int len, uint64_t seed, const std::string& out, double max_sub, double max_indel,
                          const std::string& synth = "", long long glo = 0, long long ghi = 0) {
  std::vector<GenomeRec> gs; IndexInfo info; int F = 0, S = 0, G = 0, mc = 20; unsigned long long sd = 0; const bool direct = index.empty();
  if (direct) { if (sscanf(synth.c_str(), "%d,%d,%d,%llu,%d", &F, &S, &G, &sd, &mc) < 4) die("synth-queries: --index or --synth F,S,G,seed[,max_contigs] needed"); if (ghi <= glo) { glo = 0; ghi = (long long)F * S; } }
  else { gs = read_genomes(batch_dir(index, 0) + "/genomes.bin"); info = read_info(index + "/info.toml"); }
  FILE* f = fopen(out.c_str(), "w"); if (!f) die("cannot create " + out); uint64_t s = seed; static const char B[] = "ACGT";
  for (int q = 0; q < n; q++) {
    std::string seq, gid, contig; int ci = 0; size_t st = 0; int L = len;
    if (direct) { InGenome g; for (int tries = 0;; tries++) { long long gi = glo + (long long)(splitmix64(s) % (uint64_t)(ghi - glo)); g = synth_genome((int)(gi / S), (int)(gi % S), S, G, sd, mc); ci = (int)(splitmix64(s) % g.seqs.size()); if ((int)g.seqs[ci].size() >= len || tries > 1000) break; }
      gid = g.id; contig = g.seqs[ci]; L = std::min<int>(len, (int)contig.size()); st = splitmix64(s) % (contig.size() - L + 1); }
    else { int gi = 0, tries = 0; for (;; tries++) { gi = (int)(splitmix64(s) % gs.size()); const GenomeRec& g = gs[gi]; ci = (int)(splitmix64(s) % g.seq_sizes.size()); if ((int)g.seq_sizes[ci] >= len || tries > 1000) break; }
      const GenomeRec& g = gs[gi]; size_t off = 0; for (int c = 0; c < ci; c++) off += g.seq_sizes[c] + info.contig_interval; L = std::min<int>(len, g.seq_sizes[ci]); st = splitmix64(s) % (g.seq_sizes[ci] - L + 1); gid = g.id;
      contig.resize(L); for (int i = 0; i < L; i++) { size_t p = off + st + i; contig[i] = B[(g.twobit[p >> 2] >> (6 - 2 * (p & 3))) & 3]; } }
    const size_t c0 = direct ? st : 0;
    double u = (double)(splitmix64(s) >> 11) / 9007199254740992.0 * max_sub, ind = (double)(splitmix64(s) >> 11) / 9007199254740992.0 * max_indel;
    for (int i = 0; i < L; i++) {
      int c = (int)base2bit((uint8_t)contig[c0 + i]); double r = (double)(splitmix64(s) >> 11) / 9007199254740992.0;
      if (r < ind) { if (splitmix64(s) & 1) continue; seq.push_back(B[splitmix64(s) & 3]); }
      if (r >= ind && r < ind + u) c = (c + 1 + (int)(splitmix64(s) % 3)) & 3;
      seq.push_back(B[c]);
    }
    bool rc = splitmix64(s) & 1; if (rc) { std::reverse(seq.begin(), seq.end()); for (char& c : seq) c = (c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : 'A'); }
    fprintf(f, ">q%06d g=%s c=%d s=%zu rc=%d\n%s\n", q, gid.c_str(), ci, st, (int)rc, seq.c_str());
  }
  fclose(f);
}

// ------------------------------------------------------------------ CLI
static const char* arg(int argc, char** argv, const char* name, const char* def) { for (int i = 2; i + 1 < argc; i++) if (!strcmp(argv[i], name)) return argv[i + 1]; return def; }

int main(int argc, char** argv) {
  try {
    if (argc < 2) { fprintf(stderr, "usage: lmi-tools <index|synth-queries|synth-fasta> [options]\n"); return 2; }
    std::string cmd = argv[1];
    BuildOpts o; o.masks = atoi(arg(argc, argv, "--masks", "20000")); o.chunks = atoi(arg(argc, argv, "--chunks", "16")); o.partitions = atoi(arg(argc, argv, "--partitions", "4096"));
    o.seed = atoll(arg(argc, argv, "--rand-seed", "1")); o.threads = atoi(arg(argc, argv, "--threads", "0")); o.contig_interval = atoi(arg(argc, argv, "--contig-interval", "1000"));
    for (int i = 2; i < argc; i++) { if (!strcmp(argv[i], "--fill-deserts")) o.fill_deserts = true; if (!strcmp(argv[i], "--no-fill-deserts")) o.fill_deserts = false; }
    o.max_genome = atoll(arg(argc, argv, "--max-genome", "15000000")); o.batch_size = atoi(arg(argc, argv, "--batch-size", "0"));
    o.max_desert = atoi(arg(argc, argv, "--seed-max-desert", "100")); o.seed_dist = atoi(arg(argc, argv, "--seed-in-desert-dist", "50"));
    if (o.fill_deserts && (o.seed_dist * 2 > o.max_desert || o.seed_dist < 2)) die("value of --seed-in-desert-dist should be smaller than 0.5 * --seed-max-desert");   // index.go:213
    if (cmd == "index") {
      std::string out = arg(argc, argv, "--out", ""); if (out.empty()) die("--out needed");
      std::string synth = arg(argc, argv, "--synth", ""), list = arg(argc, argv, "--in-list", "");
      if (!synth.empty()) {  // F,S,G,seed[,max_contigs]
        int F, S, G, mc = 20; unsigned long long sd; int n = sscanf(synth.c_str(), "%d,%d,%d,%llu,%d", &F, &S, &G, &sd, &mc); if (n < 4) die("--synth F,S,G,seed[,max_contigs]");
        // --synth-range lo,hi: only genomes [lo, hi) of the collection (one shard of a genome-sharded collection; same masks, same genomes as the full build)
        long long lo = 0, hi = (long long)F * S; std::string rg = arg(argc, argv, "--synth-range", ""); if (!rg.empty() && (sscanf(rg.c_str(), "%lld,%lld", &lo, &hi) != 2 || lo < 0 || hi > (long long)F * S || lo >= hi)) die("--synth-range lo,hi within [0, F*S)");
        build_index(out, (size_t)(hi - lo), [&](size_t gi) { size_t g = gi + (size_t)lo; return synth_genome((int)(g / S), (int)(g % S), S, G, sd, mc); }, o);
      } else if (!list.empty()) {  // text file: one FASTA(.gz) path per line; genome id = file name up to first ".fa"/".fna"/".fasta"
        std::vector<std::string> files; { FILE* f = fopen(list.c_str(), "r"); if (!f) die("cannot open " + list); char l[4096]; while (fgets(l, sizeof l, f)) { std::string s(l); while (!s.empty() && (s.back() == '\n' || s.back() == '\r')) s.pop_back(); if (!s.empty()) files.push_back(s); } fclose(f); }
        build_index(out, files.size(), [&](size_t gi) {
          InGenome g; std::string b = files[gi]; size_t sl = b.rfind('/'); if (sl != std::string::npos) b = b.substr(sl + 1);
          for (const char* ext : {".fasta", ".fna", ".fa", ".fastq", ".fq"}) { size_t e = b.find(ext); if (e != std::string::npos) { b = b.substr(0, e); break; } }
          g.id = b; std::vector<std::string> ids, seqs; parse_fasta(gz_slurp(files[gi]), ids, seqs);
          for (size_t i = 0; i < seqs.size(); i++) if ((int)seqs[i].size() >= o.k) { g.seq_ids.push_back(ids[i]); g.seqs.push_back(std::move(seqs[i])); }
          return g; }, o);
      } else die("index: need --synth or --in-list");
    } else if (cmd == "synth-queries") {
      long long lo = 0, hi = 0; std::string rg = arg(argc, argv, "--genome-range", ""); if (!rg.empty() && sscanf(rg.c_str(), "%lld,%lld", &lo, &hi) != 2) die("--genome-range lo,hi");
      synth_queries(arg(argc, argv, "--index", ""), atoi(arg(argc, argv, "--n", "100")), atoi(arg(argc, argv, "--len", "1000")), strtoull(arg(argc, argv, "--seed", "20260925"), 0, 10),
                    arg(argc, argv, "--out", "queries.fasta"), atof(arg(argc, argv, "--max-sub", "0.10")), atof(arg(argc, argv, "--max-indel", "0.01")), arg(argc, argv, "--synth", ""), lo, hi);
    } else if (cmd == "synth-fasta") {  // dump synthetic genomes as FASTA files (small test sets)
      int F, S, G, mc = 20; unsigned long long sd; std::string synth = arg(argc, argv, "--synth", ""); if (sscanf(synth.c_str(), "%d,%d,%d,%llu,%d", &F, &S, &G, &sd, &mc) < 4) die("--synth F,S,G,seed[,max_contigs]");
      std::string out = arg(argc, argv, "--out", "refs"); mkdir_p(out);
      for (int gi = 0; gi < F * S; gi++) { InGenome g = synth_genome(gi / S, gi % S, S, G, sd, mc); FILE* f = fopen((out + "/" + g.id + ".fa").c_str(), "w"); for (size_t c = 0; c < g.seqs.size(); c++) fprintf(f, ">%s\n%s\n", g.seq_ids[c].c_str(), g.seqs[c].c_str()); fclose(f); }
    } else if (cmd == "kat") {   // the dataset of the reference's known-answer test kv/kv-data_test.go:30-59
      std::string out = arg(argc, argv, "--out", "t.kv"); const int k = 5, lenPrefix = 2; uint64_t prefix = 0b0111ull << ((k - lenPrefix) << 1), n = 1ull << ((k - lenPrefix) << 1);
      KvWriter w(out, k, 0, 1 << lenPrefix, lenPrefix, 2, true); std::vector<KvEntry> e; for (uint64_t i = 0; i < n; i++) e.push_back({prefix | i, {i}}); for (int j = 0; j < (1 << lenPrefix); j++) w.write_mask(e); w.close();
    } else if (cmd == "kv-dump") {   // decode a chunk with the product decoder: mask, key, values..., and the anchor table
      KvChunk c = read_kv_chunk(arg(argc, argv, "--file", "")); printf("k=%d mask_offset=%d chunk_size=%d mask_prefix=%d anchor_prefix=%d use7=%d\n", c.k, c.mask_offset, c.chunk_size, c.mask_prefix, c.anchor_prefix, (int)c.use7);
      for (int m = 0; m < c.chunk_size; m++) { const KvMaskData& md = c.masks[m]; for (size_t t = 0; t < md.keys.size(); t++) { printf("K\t%d\t%llu", m, (unsigned long long)md.keys[t]); for (uint32_t v = md.val_off[t]; v < md.val_off[t + 1]; v++) printf("\t%llu", (unsigned long long)md.vals[v]); printf("\n"); }
        for (size_t a = 0; a < c.anchor_start[m].size(); a++) if (c.anchor_start[m][a] != 0xFFFFFFFFu) printf("A\t%d\t%zu\t%u\n", m, a, c.anchor_start[m][a]); }
    } else if (cmd == "varint-test") {   // util/varint-GB_test.go:54-100 restated: round trip + control-byte length
      uint64_t s = strtoull(arg(argc, argv, "--seed", "1"), 0, 10); int bad = 0; for (int it = 0; it < 2000000; it++) { uint64_t a = splitmix64(s) >> (splitmix64(s) & 63), b = splitmix64(s) >> (splitmix64(s) & 63), x, y; uint8_t buf[16], ctrl; int n = put_u64s(buf, a, b, &ctrl);
        int m = get_u64s(ctrl, buf, &x, &y); if (x != a || y != b || n != m || n != ((ctrl >> 3) & 7) + (ctrl & 7) + 2) bad++; } printf("varint mismatches: %d\n", bad); return bad ? 1 : 0;
    } else die("unknown command " + cmd);
  } catch (std::exception& e) { fprintf(stderr, "[lmi-tools] error: %s\n", e.what()); return 1; }
  return 0;
}
