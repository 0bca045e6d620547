// oracle.cpp — Index.Search restated on the CPU + C exports for ctypes.  TEST INFRASTRUCTURE ONLY (see oracle_core.hpp).
// Follows lexicmap/cmd/lib-index-search.go:1191-2940 stage by stage. Record layouts mirror include/lexicmap_gpu.h so the
// tests can compare byte-for-byte, but nothing here is shared with the product.
#include "oracle_core.hpp"
#include <omp.h>
#include <chrono>

using namespace lmo;

extern "C" {
typedef struct lmo_params { int32_t min_prefix, min_single_prefix, top_n_genomes, top_n_chains; float max_gap, max_distance; int32_t ext_len, ext_len2; double min_qcov_genome, max_evalue;
  int32_t align_max_gap, align_min_len, align_band, output_seq; double min_pident, min_qcov_hsp; int32_t wfa_adaptive, reserved; } lmo_params;
typedef struct lmo_pa { uint64_t genome; uint32_t query; int32_t t_begin, t_end, rc; int32_t qb, qe, tb, te, aligned_q, aligned_t, matched, n_anchors; } lmo_pa;   // one Chain2Result of SeqComparator.Compare in window coordinates (lib-seq_compare.go:335-522)
typedef struct lmo_hsp { uint32_t query, hits; uint64_t genome; uint32_t seq_idx, n_seqs, chunk_idx, n_chunks; int32_t seq_len, cls, hsp, qb, qe, tb, te, rc, alen, matches, gaps, score, bitscore, pad0;
  double evalue, qcov_hsp, pident, qcov_gnm; uint64_t cigar_off; uint32_t cigar_len, pad; } lmo_hsp;
typedef struct lmo_anchor { uint64_t genome; uint32_t query; int32_t qbegin, tbegin; uint8_t len, qrc, trc, pad; } lmo_anchor;
typedef struct lmo_chain { uint64_t genome; uint32_t query; float score; int32_t n_seeds, q0, t0, len0, q1, t1, len1, rc; } lmo_chain;
}

namespace {

struct SD { bool rc; double sim; int nseeds; int seq_idx, nseqs, seqlen; std::string seqid; std::vector<Chain2> chains; uint32_t chunk_idx = 0, n_chunks = 1; };
struct GenomeRes { uint64_t bgi; std::vector<Sub> subs; std::vector<std::vector<int32_t>> chains; float score = 0; std::vector<SD> sds; double af = 0; };

struct StageSink { std::vector<lmo_anchor>* anchors = nullptr; std::vector<lmo_chain>* chains = nullptr; std::vector<lmo_pa>* pas = nullptr; };

// genome.Reader.SubSeq3 genome/genome.go:931-1143 (+ meta parse)
static GenomeMeta genome_meta(const GenomeBatchFile& b, int idx) {
  GenomeMeta g; const uint8_t* d = b.data.data(); size_t p = (size_t)b.rec_off[idx]; size_t l = rd_be(d + p, 2); p += 2 + l;
  g.genome_size = (int)rd_be(d + p, 4); g.num_seqs = (int)rd_be(d + p + 8, 4); p += 12;
  for (int i = 0; i < g.num_seqs; i++) { g.seq_sizes.push_back((int)rd_be(d + p, 4)); l = rd_be(d + p + 4, 2); p += 6; g.seq_ids.emplace_back((const char*)d + p, l); p += l; }
  g.seq_offset = p; return g;
}
static std::string subseq3(const GenomeBatchFile& b, int idx, const GenomeMeta& g, int start, int end) {
  int nBases = (int)b.nbases[idx]; if (start < 0) start = 0; if (end >= nBases - 1) end = nBases - 1; if (end < start) end = start;
  const uint8_t* d = b.data.data() + g.seq_offset + 8; std::string s; s.resize(end - start + 1); static const char B[] = "ACGT";
  for (int i = start; i <= end; i++) s[i - start] = B[(d[i >> 2] >> (6 - 2 * (i & 3))) & 3];
  return s;
}
static void rc_inplace(std::string& s) { std::reverse(s.begin(), s.end()); for (char& c : s) c = (c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c); }
static int coverage_len(std::vector<std::array<int, 2>> r) {  // coverageLen lib-seq_compare.go:270-308
  if (r.empty()) return 0; if (r.size() == 1) return r[0][1] - r[0][0] + 1; std::stable_sort(r.begin(), r.end(), [](const std::array<int, 2>& a, const std::array<int, 2>& b) { return a[0] < b[0]; });
  int start = r[0][0], end = r[0][1], tot = 0; for (size_t i = 1; i < r.size(); i++) { if (r[i][0] > end) { tot += end - start + 1; start = r[i][0]; end = r[i][1]; continue; } if (r[i][1] <= end) continue; end = r[i][1]; } return tot + end - start + 1;
}

static void search_one(const Index& ix, const Params& P, const std::string& qseq_in, uint32_t qidx, std::vector<GenomeRes>& out, StageSink* sink) {
  out.clear(); std::string s = qseq_in; for (char& c : s) if (c >= 'a' && c <= 'z') c -= 32;  // search.go:582-587
  const int K = ix.k, qlen = (int)s.size(); if (qlen < K) return;                                // search.go:571-575
  // ---- (1) mask + low-complexity filter
  MaskResult mr = mask_fast(ix, (const uint8_t*)s.data(), qlen);
  for (int i = 0; i < ix.n_masks; i++) if (low_complexity(mr.kmers[i], K)) mr.kmers[i] = 0;
  // ---- (1b) reversed k-mers for suffix matching :1268-1350 (arrival order made deterministic: ascending old mask)
  std::vector<std::vector<uint64_t>> kmersR(ix.n_masks); std::vector<std::vector<int>> locsR(ix.n_masks);
  for (int i = 0; i < ix.n_masks; i++) { uint64_t km = mr.kmers[i]; if (km == 0) continue; uint64_t rv = kmer_reverse(km, K);
    size_t nm = xor_argmin_sorted(ix.masks.data(), 0, ix.masks.size(), rv);
    bool existed = false; for (uint64_t v : kmersR[nm]) if (v == rv) { existed = true; break; } if (!existed) { kmersR[nm].push_back(rv); locsR[nm].push_back(i); } }
  // ---- (2) seed lookup + anchor materialisation :1357-1643
  std::map<uint64_t, GenomeRes> m; std::vector<KvHit> hits;
  for (const KvChunkFile& c : ix.chunks) { hits.clear();
    for (int iq = 0; iq < c.chunk_size; iq++) kv_probe(c, iq, mr.kmers[c.chunk_index + iq], P.min_prefix, false, 0, hits);
    for (int iq = 0; iq < c.chunk_size; iq++) { const auto& l = kmersR[c.chunk_index + iq]; for (size_t j = 0; j < l.size(); j++) kv_probe(c, iq, l[j], P.min_prefix, true, (int)j, hits); }
    for (const KvHit& sr : hits) { int kPrefix = sr.len; const std::vector<int>& locs = sr.is_suffix ? mr.locs[locsR[sr.iquery][sr.iquery2]] : mr.locs[sr.iquery];
      for (int posQ0 : locs) { bool rcQ = posQ0 & 1; int posQ = posQ0 >> 1;
        for (uint64_t refpos : sr.values) { uint64_t bgi = refpos >> 30; int posT = (int)((refpos << 34) >> 36); bool rvT = refpos & 1, rcT = (refpos >> 1) & 1; int beginQ, beginT;
          if (!rvT) { beginQ = rcQ ? posQ + K - kPrefix : posQ; beginT = rcT ? posT + K - kPrefix : posT; } else { beginQ = rcQ ? posQ : posQ + K - kPrefix; beginT = rcT ? posT : posT + K - kPrefix; }
          GenomeRes& r = m[bgi]; r.bgi = bgi; r.subs.push_back({(int32_t)beginQ, (int32_t)beginT, (uint8_t)kPrefix, rcT, rcQ}); } } } }
  if (m.empty()) return;
  if (sink && sink->anchors) for (auto& e : m) for (const Sub& a : e.second.subs) sink->anchors->push_back({e.first, qidx, a.q, a.t, a.len, (uint8_t)a.qrc, (uint8_t)a.trc, 0});
  // ---- (3) chaining :1702-1805
  ChainOpts co{P.max_gap, seed_weight((float)P.min_single_prefix), P.max_distance, P.top_n_chains}; std::vector<GenomeRes*> rs;
  for (auto& e : m) { GenomeRes& r = e.second; if (r.subs.size() > 1) clear_subs(r.subs, K); r.chains = chain1(r.subs, co, &r.score); if (r.score < co.min_score) continue; rs.push_back(&r); }
  if (P.top_n_genomes > 0 && (int)rs.size() > P.top_n_genomes) { std::stable_sort(rs.begin(), rs.end(), [](GenomeRes* a, GenomeRes* b) { return a->score > b->score; }); rs.resize(P.top_n_genomes); }
  if (rs.empty()) return;
  std::stable_sort(rs.begin(), rs.end(), [](GenomeRes* a, GenomeRes* b) { return (a->bgi & 131071) < (b->bgi & 131071); });  // :1848-1853
  // ---- (4)+(5) pseudo-alignment and alignment :1834-2763
  QueryTable T = build_query_table((const uint8_t*)s.data(), qlen, 31);  // SeqComparatorOptions.K = 31 search.go:361
  Chain2Opts c2{P.align_max_gap, (int)((double)P.align_min_len * P.min_pident / 100), P.align_min_len, P.align_band / 2, P.align_band, 15};
  const int extLen = P.ext_len, contigInterval = ix.contig_interval; const bool has_chunks = !ix.genome_chunks.empty();
  for (GenomeRes* rp : rs) {
    GenomeRes& r = *rp; int refBatch = (int)(r.bgi >> 17), refID = (int)(r.bgi & 131071); const GenomeBatchFile& gb = ix.batches[refBatch]; GenomeMeta gm = genome_meta(gb, refID);
    std::stable_sort(r.chains.begin(), r.chains.end(), [&](const std::vector<int32_t>& a, const std::vector<int32_t>& b) { return r.subs[a[0]].t < r.subs[b[0]].t; });  // :1967-1974
    if (sink && sink->chains) for (auto& ch : r.chains) { const Sub &f = r.subs[ch[0]], &l = r.subs[ch.back()]; bool rc = ch.size() == 1 ? (l.qrc != l.trc) : (f.t > l.t);
        sink->chains->push_back({r.bgi, qidx, r.score, (int32_t)ch.size(), f.q, f.t, f.len, l.q, l.t, l.len, (int32_t)rc}); }
    std::set<std::array<int, 6>> alignmentKeys; int iSeq = 0, iSeqPre = -1;
    for (auto& chain : r.chains) {
      int nSeeds = (int)chain.size(); const Sub& f = r.subs[chain[0]]; int qb = f.q, tb = f.t; const Sub& sub = r.subs[chain[nSeeds - 1]]; int qe = sub.q + sub.len - 1, te = sub.t + sub.len - 1;
      bool rc = (nSeeds == 1) ? (sub.qrc != sub.trc) : (tb > sub.t); int tBegin, tEnd;
      if (rc) { tBegin = sub.t - extLen; if (tBegin < 0) tBegin = 0; tEnd = tb + sub.len - 1 + extLen; } else { tBegin = tb - extLen; if (tBegin < 0) tBegin = 0; tEnd = te + extLen; }
      int qBegin = qb - std::min(qb, extLen), qEnd = qe + std::min(qlen - qe - 1, extLen);
      std::string tseq = subseq3(gb, refID, gm, tBegin, tEnd); if ((int)tseq.size() < tEnd - tBegin + 1) tEnd -= tEnd - tBegin + 1 - (int)tseq.size();
      if (rc) rc_inplace(tseq);
      std::vector<Chain2> crChains = compare(T, (uint32_t)qBegin, (uint32_t)qEnd, tseq, c2, 11); if (crChains.empty()) continue;
      if (sink && sink->pas) for (const Chain2& c : crChains) sink->pas->push_back({r.bgi, qidx, tBegin, tEnd, (int32_t)rc, c.qb, c.qe, c.tb, c.te, c.aligned_q, c.aligned_t, c.matched, c.n_anchors});
      iSeqPre = -1; std::vector<Chain2> cur; const int tlenSeq = (int)tseq.size();
      auto flush = [&](std::vector<Chain2>& chains2, bool variantA, int iSeqUse) {
        bool hasResult = false; double maxSim = 0;
        for (Chain2& c : chains2) {
          c.af = (double)c.aligned_q / (double)qlen * 100;  // Update2
          if (c.qb >= c.qe + 1) { c.dead = true; continue; }
          int start, end; if (rc) { start = tEnd - c.te - c.t_pos_offset_begin; end = tEnd - c.tb - c.t_pos_offset_begin + 1; } else { start = c.t_pos_offset_begin + c.tb - tBegin; end = c.t_pos_offset_begin + c.te - tBegin + 1; }
          if (start >= end) { c.dead = true; continue; }
          int ext2 = P.ext_len2; if (c.aligned_q > 1000000) ext2 += 80; else if (c.aligned_q > 250000) ext2 += 40; else if (c.aligned_q > 50000) ext2 += 20; else if (c.aligned_q > 10000) ext2 += 10;
          if (start < 0 || end > tlenSeq || c.qb < 0 || c.qe + 1 > qlen) { c.dead = true; continue; }  // Go would panic on the slice; never observed
          Extended ex = extend_match(s, tseq, c.qb, c.qe + 1, start, end, ext2, c.tb, c.max_ext_len, rc);
          int ql = ex.end1 - ex.start1, tl = ex.end2 - ex.start2; WfaResult cg = wfa_align(s.data() + ex.start1, ql, tseq.data() + ex.start2, tl, P.wfa_adaptive);
          score_evalue(cg, ql, ix.total_bases, &c.score, &c.bitscore, &c.evalue); if (c.evalue > P.max_evalue) { c.dead = true; continue; }
          c.qb -= ex.s1; c.qe += ex.e1; c.qb = c.qb + cg.qbegin - 1; c.qe = c.qe - (ql - cg.qend);
          if (rc) { c.tb -= ex.e2; c.te += ex.s2; c.tb = c.tb + (tl - cg.tend); c.te = variantA ? (c.te - cg.tbegin - 1) : (c.te - (cg.tbegin - 1)); }   // :2284-2285 vs :2551-2552
          else { c.tb -= ex.s2; c.te += ex.e2; c.tb = c.tb + cg.tbegin - 1; c.te = c.te - (tl - cg.tend); }
          c.aligned_q = c.qe - c.qb + 1; c.aligned_len = cg.align_len; c.matched = cg.matches; c.gaps = cg.gaps; c.af = (double)c.aligned_q / (double)qlen * 100; if (c.af > 100) c.af = 100;
          c.pident = (double)c.matched / (double)cg.align_len * 100;
          if (c.af < P.min_qcov_hsp || c.pident < P.min_pident) { c.dead = true; continue; }
          if (P.output_seq) { c.cigar.clear(); c.qseq.clear(); c.tseq.clear(); c.align.clear(); int qi = ex.start1 + cg.qbegin - 1, ti = ex.start2 + cg.tbegin - 1;   // AlignmentText(&_qseq, &_tseq, true): first M .. last M
            for (uint64_t op : trim_ops(cg.ops)) { char o = (char)(op >> 32); uint32_t n = (uint32_t)(op & 0xffffffffu);
              for (uint32_t x = 0; x < n; x++) {   // WFA ops: I consumes the target only, D the query only (swapped below for SAM, :2331-2338)
                if (o == 'M' || o == 'X') { c.qseq.push_back((char)s[qi++]); c.tseq.push_back((char)tseq[ti++]); c.align.push_back(o == 'M' ? '|' : ' '); }
                else if (o == 'I') { c.qseq.push_back('-'); c.tseq.push_back((char)tseq[ti++]); c.align.push_back(' '); }
                else { c.qseq.push_back((char)s[qi++]); c.tseq.push_back('-'); c.align.push_back(' '); } }
              if (o == 'D') o = 'I'; else if (o == 'I') o = 'D'; c.cigar += std::to_string(n); c.cigar.push_back(o); } }
          double sim = (double)c.bitscore * c.pident; if (sim > maxSim) maxSim = sim; hasResult = true;
        }
        if (hasResult) { SD sd; sd.rc = rc; sd.nseeds = nSeeds; sd.sim = maxSim; sd.seq_idx = iSeqUse; sd.nseqs = (int)gm.seq_ids.size(); sd.seqlen = gm.seq_sizes[iSeqUse]; sd.seqid = gm.seq_ids[iSeqUse]; sd.chains = chains2;
          { auto ci = ix.genome_chunks.find(r.bgi); if (ci != ix.genome_chunks.end()) { sd.n_chunks = ci->second.n; sd.chunk_idx = ci->second.i; } }   // :2375-2385 / :2643-2654
          r.sds.push_back(std::move(sd)); }
      };
      auto convert = [&](Chain2& c, int qb_, int qe_, int tb_, int te_, int tPosOffsetBegin, int iS) {  // :2167-2200 / :2423-2454
        c.qb = qb_; c.qe = qe_; c.t_pos_offset_begin = tPosOffsetBegin;
        if (rc) { c.tb = tBegin - tPosOffsetBegin + (tlenSeq - te_ - 1); if (c.tb < 0) { c.qe += c.tb; c.aligned_q += c.tb; c.tb = 0; }
          c.te = tBegin - tPosOffsetBegin + (tlenSeq - tb_ - 1); if (c.te > gm.seq_sizes[iS] - 1) { c.qb += c.te - (gm.seq_sizes[iS] - 1); c.te = gm.seq_sizes[iS] - 1; } }
        else { c.tb = tBegin - tPosOffsetBegin + tb_; if (c.tb < 0) { c.qb -= c.tb; c.aligned_q += c.tb; c.tb = 0; }
          c.te = tBegin - tPosOffsetBegin + te_; if (c.te > gm.seq_sizes[iS] - 1) { c.qe -= c.te - (gm.seq_sizes[iS] - 1); c.te = gm.seq_sizes[iS] - 1; } }
        c.max_ext_len = gm.seq_sizes[iS] - 1 - c.te;
      };
      for (Chain2& c : crChains) {
        int cqb = c.qb, cqe = c.qe, ctb = c.tb, cte = c.te; iSeq = 0; int tPosOffsetBegin = 0, tPosOffsetEnd = 0;
        if (gm.num_seqs > 1) {
          iSeq = -1; int _begin, _end; if (rc) { _begin = tEnd - cte + K; _end = tEnd - ctb - K; } else { _begin = tBegin + ctb + K; _end = tBegin + cte - K; }
          if (_begin >= _end) { if (rc) { _begin = tEnd - cte; _end = tEnd - ctb; } else { _begin = tBegin + ctb; _end = tBegin + cte; } }
          for (int j = 0; j < (int)gm.seq_sizes.size(); j++) { int l = gm.seq_sizes[j]; tPosOffsetEnd += l - 1;
            if (_begin + K >= tPosOffsetBegin && _end - K <= tPosOffsetEnd) { iSeq = j; break; } else if (_end < tPosOffsetBegin) { iSeq = -1; break; }
            tPosOffsetEnd += contigInterval + 1; tPosOffsetBegin = tPosOffsetEnd; }
          if (iSeq < 0) continue;
          if (iSeqPre >= 0 && iSeq != iSeqPre) {
            int iSeq0 = iSeq; iSeq = iSeqPre; convert(c, cqb, cqe, ctb, cte, tPosOffsetBegin, iSeq);
            if (!cur.empty()) flush(cur, true, iSeq);
            iSeqPre = -1; cur.clear();
            std::array<int, 6> key{c.qb, c.qe, c.tb, c.te, iSeq, (int)rc}; if (!alignmentKeys.count(key)) { cur.push_back(c); alignmentKeys.insert(key); }
            iSeq = iSeq0; continue;
          }
        }
        iSeqPre = iSeq; convert(c, cqb, cqe, ctb, cte, tPosOffsetBegin, iSeq);
        std::array<int, 6> key{c.qb, c.qe, c.tb, c.te, iSeq, (int)rc}; if (!alignmentKeys.count(key)) { cur.push_back(c); alignmentKeys.insert(key); }
      }
      if (iSeq >= 0 && !cur.empty()) flush(cur, false, iSeq);
    }
    if (r.sds.empty()) continue;
    if (has_chunks) continue;   // :2701 "if hasGenomeChunks, do not filter results now"
    std::vector<std::array<int, 2>> regions; for (SD& sd : r.sds) for (Chain2& c : sd.chains) if (!c.dead) regions.push_back({c.qb, c.qe});
    r.af = (double)coverage_len(regions) / (double)qlen * 100; if (r.af > 100) r.af = 100; if (r.af < P.min_qcov_genome) { r.sds.clear(); continue; }
    std::stable_sort(r.sds.begin(), r.sds.end(), [](const SD& a, const SD& b) { return a.sim > b.sim; });  // :2745-2747
  }
  // ---- (5b) merge the results of the chunks of a split genome :2797-2913. The reference merges into whichever chunk's result arrived first
  // (goroutine order); made deterministic here: into the chunk that comes first in `rs` (genome-index order), the others appended in that order.
  if (has_chunks) {
    std::map<uint32_t, GenomeRes*> first;
    for (GenomeRes* rp : rs) { if (rp->sds.empty()) continue; auto ci = ix.genome_chunks.find(rp->bgi); if (ci == ix.genome_chunks.end()) continue;
      auto f = first.find(ci->second.group); if (f == first.end()) { first[ci->second.group] = rp; continue; }
      for (SD& sd : rp->sds) f->second->sds.push_back(std::move(sd)); rp->sds.clear(); }
    for (GenomeRes* rp : rs) { GenomeRes& r = *rp; if (r.sds.empty()) continue;   // recompute the query coverage per genome, filter, sort :2856-2897
      std::vector<std::array<int, 2>> regions; for (SD& sd : r.sds) for (Chain2& c : sd.chains) if (!c.dead) regions.push_back({c.qb, c.qe});
      r.af = (double)coverage_len(regions) / (double)qlen * 100; if (r.af > 100) r.af = 100; if (r.af < P.min_qcov_genome) { r.sds.clear(); continue; }
      std::stable_sort(r.sds.begin(), r.sds.end(), [](const SD& a, const SD& b) { return a.sim > b.sim; }); }
  }
  // ---- (6) finish :2919-2932
  std::vector<GenomeRes*> rs2; for (GenomeRes* r : rs) if (!r->sds.empty()) rs2.push_back(r);
  std::stable_sort(rs2.begin(), rs2.end(), [](GenomeRes* a, GenomeRes* b) { return a->sds[0].sim > b->sds[0].sim; });
  for (GenomeRes* r : rs2) {  // SortBySeqID :1042-1096 == stable grouping by first appearance of each seqid
    std::vector<SD> g; std::vector<char> used(r->sds.size(), 0);
    for (size_t i = 0; i < r->sds.size(); i++) { if (used[i]) continue; for (size_t j = i; j < r->sds.size(); j++) if (!used[j] && r->sds[j].seqid == r->sds[i].seqid) { used[j] = 1; g.push_back(r->sds[j]); } }
    r->sds.swap(g);
  }
  for (GenomeRes* r : rs2) { GenomeRes g; g.bgi = r->bgi; g.af = r->af; g.score = r->score; g.sds = std::move(r->sds); out.push_back(std::move(g)); }
}

struct Handle { Index ix; };
struct Rows { std::vector<lmo_hsp> rows; std::string pool; std::vector<std::string> seqids; };

static Params to_params(const lmo_params* p) { Params P; if (!p) return P; P.min_prefix = p->min_prefix; P.min_single_prefix = p->min_single_prefix; P.top_n_genomes = p->top_n_genomes; P.top_n_chains = p->top_n_chains;
  P.max_gap = p->max_gap; P.max_distance = p->max_distance; P.ext_len = p->ext_len; P.ext_len2 = p->ext_len2; P.min_qcov_genome = p->min_qcov_genome; P.max_evalue = p->max_evalue;
  P.align_max_gap = p->align_max_gap; P.align_min_len = p->align_min_len; P.align_band = p->align_band; P.min_pident = p->min_pident; P.min_qcov_hsp = p->min_qcov_hsp; P.output_seq = p->output_seq; P.wfa_adaptive = p->wfa_adaptive; return P; }

static void rows_of(uint32_t q, const std::vector<GenomeRes>& res, Rows& R) {  // printResult search.go:437-533
  for (const GenomeRes& r : res) { int cls = 1, j = 1;
    for (const SD& sd : r.sds) { for (const Chain2& c : sd.chains) { if (c.dead) continue; lmo_hsp h; memset(&h, 0, sizeof h); h.query = q; h.hits = (uint32_t)res.size(); h.genome = r.bgi; h.seq_idx = sd.seq_idx; h.n_seqs = sd.nseqs; h.chunk_idx = sd.chunk_idx; h.n_chunks = sd.n_chunks; h.seq_len = sd.seqlen;
        h.cls = cls; h.hsp = j; h.qb = c.qb; h.qe = c.qe; h.tb = c.tb; h.te = c.te; h.rc = sd.rc; h.alen = c.aligned_len; h.matches = c.matched; h.gaps = c.gaps; h.score = c.score; h.bitscore = c.bitscore; h.evalue = c.evalue;
        h.qcov_hsp = c.af; h.pident = c.pident; h.qcov_gnm = r.af; h.cigar_off = R.pool.size(); h.cigar_len = (uint32_t)c.cigar.size(); R.pool += c.cigar; R.pool += c.qseq; R.pool += c.tseq; R.pool += c.align; /* pool entry: cigar | qseq | sseq | align (alen bytes each) */ R.rows.push_back(h); R.seqids.push_back(sd.seqid); j++; } cls++; } }
}
}  // namespace

extern "C" {
void lmo_default_params(lmo_params* p) { Params d; p->min_prefix = d.min_prefix; p->min_single_prefix = d.min_single_prefix; p->top_n_genomes = 0; p->top_n_chains = 0; p->max_gap = d.max_gap; p->max_distance = d.max_distance; p->ext_len = d.ext_len; p->ext_len2 = d.ext_len2;
  p->min_qcov_genome = 0; p->max_evalue = 10; p->align_max_gap = 20; p->align_min_len = 50; p->align_band = 100; p->output_seq = 0; p->min_pident = 70; p->min_qcov_hsp = 0; p->wfa_adaptive = 1; p->reserved = 0; }
static thread_local std::string g_err;
const char* lmo_last_error() { return g_err.c_str(); }
void* lmo_open(const char* dir) { try { Handle* h = new Handle; h->ix.open(dir); return h; } catch (std::exception& e) { g_err = e.what(); return nullptr; } }
void lmo_close(void* h) { delete (Handle*)h; }
const char* lmo_genome_name(void* hh, uint64_t g) { Handle* h = (Handle*)hh; auto it = h->ix.id2name.find(g); return it == h->ix.id2name.end() ? "" : it->second.c_str(); }
int64_t lmo_total_bases(void* hh) { return ((Handle*)hh)->ix.total_bases; }

// full search of a batch; threads>1 uses OpenMP over queries (the CPU baseline). Returns rows handle.
void* lmo_search_batch(void* hh, const lmo_params* p, const uint8_t* seqs, const uint64_t* off, int32_t n, int threads) {
  Handle* h = (Handle*)hh; Params P = to_params(p); std::vector<std::vector<GenomeRes>> res(n); std::string err;
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 0 ? threads : 1)
  for (int q = 0; q < n; q++) { try { search_one(h->ix, P, std::string((const char*)seqs + off[q], off[q + 1] - off[q]), q, res[q], nullptr); } catch (std::exception& e) {
#pragma omp critical
      err = e.what(); } }
  if (!err.empty()) { g_err = err; return nullptr; }
  Rows* R = new Rows; for (int q = 0; q < n; q++) rows_of(q, res[q], *R); return R;
}
uint64_t lmo_rows(void* r, const lmo_hsp** rows, const char** pool) { Rows* R = (Rows*)r; *rows = R->rows.data(); if (pool) *pool = R->pool.data(); return R->rows.size(); }
const char* lmo_row_seqid(void* r, uint64_t i) { return ((Rows*)r)->seqids[i].c_str(); }
void lmo_rows_free(void* r) { delete (Rows*)r; }

// ---- stage-wise
int lmo_mask_batch(void* hh, const uint8_t* seqs, const uint64_t* off, int32_t n, uint64_t* kmers, uint32_t* nlocs, uint32_t* minloc, uint64_t* suf, uint64_t suf_cap, uint64_t* n_suf, int bruteforce) {
  Handle* h = (Handle*)hh; const Index& ix = h->ix; uint64_t ns_ = 0;
  for (int q = 0; q < n; q++) { std::string s((const char*)seqs + off[q], off[q + 1] - off[q]); for (char& c : s) if (c >= 'a' && c <= 'z') c -= 32;
    MaskResult mr = bruteforce ? mask_bruteforce(ix, (const uint8_t*)s.data(), (int)s.size()) : mask_fast(ix, (const uint8_t*)s.data(), (int)s.size());
    std::vector<std::array<uint64_t, 4>> trip; std::set<std::pair<uint64_t, uint64_t>> seen;
    for (int i = 0; i < ix.n_masks; i++) { uint64_t km = mr.kmers[i]; if ((int)s.size() < ix.k) km = 0; if (km != 0 && low_complexity(km, ix.k)) km = 0; if (mr.locs[i].empty()) km = 0;
      size_t o = (size_t)q * ix.n_masks + i; kmers[o] = km; nlocs[o] = km ? (uint32_t)mr.locs[i].size() : 0; minloc[o] = km ? (uint32_t)*std::min_element(mr.locs[i].begin(), mr.locs[i].end()) : 0;
      if (km) { uint64_t rv = kmer_reverse(km, ix.k); uint64_t nm = xor_argmin_sorted(ix.masks.data(), 0, ix.masks.size(), rv); if (seen.insert({nm, rv}).second) trip.push_back({(uint64_t)q, nm, (uint64_t)i, rv}); } }
    std::sort(trip.begin(), trip.end()); for (auto& t : trip) { if (ns_ < suf_cap) memcpy(suf + 4 * ns_, t.data(), 32); ns_++; } }
  *n_suf = ns_; return 0;
}
static void* stage(void* hh, const lmo_params* p, const uint8_t* seqs, const uint64_t* off, int32_t n, int which, uint64_t* n_out) {
  Handle* h = (Handle*)hh; Params P = to_params(p); auto* A = new std::vector<lmo_anchor>; auto* C = new std::vector<lmo_chain>; std::vector<lmo_pa> PA; StageSink sk; sk.anchors = A; sk.chains = C; if (which == 2) sk.pas = &PA; std::vector<GenomeRes> res;
  // for the chain stage the downstream stages are irrelevant but harmless
  for (int q = 0; q < n; q++) search_one(h->ix, P, std::string((const char*)seqs + off[q], off[q + 1] - off[q]), q, res, &sk);
  if (which == 2) { *n_out = PA.size(); void* out = malloc(PA.size() * sizeof(lmo_pa) + 1); memcpy(out, PA.data(), PA.size() * sizeof(lmo_pa)); delete A; delete C; return out; }
  if (which == 0) { std::sort(A->begin(), A->end(), [](const lmo_anchor& a, const lmo_anchor& b) { if (a.query != b.query) return a.query < b.query; if (a.genome != b.genome) return a.genome < b.genome; if (a.qbegin != b.qbegin) return a.qbegin < b.qbegin;
      if (a.len != b.len) return a.len > b.len; if (a.tbegin != b.tbegin) return a.tbegin < b.tbegin; if (a.qrc != b.qrc) return a.qrc < b.qrc; return a.trc < b.trc; });
    *n_out = A->size(); void* out = malloc(A->size() * sizeof(lmo_anchor) + 1); memcpy(out, A->data(), A->size() * sizeof(lmo_anchor)); delete A; delete C; return out; }
  std::stable_sort(C->begin(), C->end(), [](const lmo_chain& a, const lmo_chain& b) { if (a.query != b.query) return a.query < b.query; return a.genome < b.genome; });
  *n_out = C->size(); void* out = malloc(C->size() * sizeof(lmo_chain) + 1); memcpy(out, C->data(), C->size() * sizeof(lmo_chain)); delete A; delete C; return out;
}
void* lmo_anchor_batch(void* hh, const lmo_params* p, const uint8_t* seqs, const uint64_t* off, int32_t n, uint64_t* n_out) { try { return stage(hh, p, seqs, off, n, 0, n_out); } catch (std::exception& e) { g_err = e.what(); return nullptr; } }
void* lmo_chain_batch(void* hh, const lmo_params* p, const uint8_t* seqs, const uint64_t* off, int32_t n, uint64_t* n_out) { try { return stage(hh, p, seqs, off, n, 1, n_out); } catch (std::exception& e) { g_err = e.what(); return nullptr; } }
void* lmo_pseudoalign_batch(void* hh, const lmo_params* p, const uint8_t* seqs, const uint64_t* off, int32_t n, uint64_t* n_out) { try { return stage(hh, p, seqs, off, n, 2, n_out); } catch (std::exception& e) { g_err = e.what(); return nullptr; } }
// WFA on pairs: off[2n+1]; returns '\n'-joined CIGARs in wfa convention (not swapped), untrimmed
char* lmo_wfa_batch(const uint8_t* seqs, const uint64_t* off, int32_t n, int32_t adaptive, uint64_t* out_len) {
  std::string o; for (int i = 0; i < n; i++) { WfaResult w = wfa_align((const char*)seqs + off[2 * i], (int)(off[2 * i + 1] - off[2 * i]), (const char*)seqs + off[2 * i + 1], (int)(off[2 * i + 2] - off[2 * i + 1]), adaptive);
    for (uint64_t op : w.ops) { o += std::to_string((uint32_t)(op & 0xffffffffu)); o.push_back((char)(op >> 32)); } o.push_back('\n'); }
  char* c = (char*)malloc(o.size() + 1); memcpy(c, o.data(), o.size() + 1); *out_len = o.size(); return c;
}
// on-disk seed lookup for one (mask,kmer): KAT helper (kv-data_test.go:208-283). Returns number of results; lens/values optional.
// standalone handle over one kv-data file (no .lmi directory) for the known-answer test
void* lmo_open_kv(const char* file) { try { Handle* h = new Handle; h->ix.chunks.resize(1); h->ix.chunks[0].data = slurp(file); Index::read_kv_index(std::string(file) + ".idx", h->ix.chunks[0]); h->ix.k = h->ix.chunks[0].k; return h; } catch (std::exception& e) { g_err = e.what(); return nullptr; } }
int lmo_kv_search(void* hh, int mask, uint64_t kmer, int p, int reversed, int check_flag, uint8_t* lens, uint64_t* first_values, int cap) {
  Handle* h = (Handle*)hh; std::vector<KvHit> hits; for (const KvChunkFile& c : h->ix.chunks) if (mask >= c.chunk_index && mask < c.chunk_index + c.chunk_size) kv_probe(c, mask - c.chunk_index, kmer, p, reversed, 0, hits, check_flag != 0);
  for (int i = 0; i < (int)hits.size() && i < cap; i++) { if (lens) lens[i] = hits[i].len; if (first_values) first_values[i] = hits[i].values[0]; } return (int)hits.size();
}
// tree.Search emulation on an explicit key list (sorted by the caller): returns the reported range
int lmo_tree_search(const uint64_t* keys, int n, int k, uint64_t key, int p, int* lo, int* hi) { QueryTable T; T.k = k; T.keys.assign(keys, keys + n); size_t a = 0, b = 0; bool ok = tree_search(T, key, p, &a, &b); *lo = (int)a; *hi = (int)b; return ok; }
// genome.Reader.SubSeq3 (genome/genome.go:931-1143)
int lmo_subseq(void* hh, uint64_t bgi, int start, int end, char* out, int cap) { Handle* h = (Handle*)hh; const GenomeBatchFile& gb = h->ix.batches[bgi >> 17]; GenomeMeta gm = genome_meta(gb, (int)(bgi & 131071)); std::string s = subseq3(gb, (int)(bgi & 131071), gm, start, end); int n = std::min<int>(cap, (int)s.size()); memcpy(out, s.data(), n); return (int)s.size(); }
int lmo_dust(uint64_t kmer, int k) { return dust(kmer, k); }
void lmo_free(void* p) { free(p); }
}
