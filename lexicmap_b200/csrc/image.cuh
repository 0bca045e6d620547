// image.cuh — the GPU-resident index image (B200: everything lives in HBM, decoded once).
//
// Replaces the reference's per-query file seeks + VARINT-GB decode (kv-searcher.go:366-395) and genome file reads
// (genome.go:1047-1062) with flat arrays:
//   seeds:   keys[E] (sorted distinct k-mers per mask bucket), val_off[E+1], vals[V], bucket_off[m+1],
//            anchor_start[m * 4^anchorPrefix] (u32 bucket-relative; from the .idx files so the writer's
//            last-run-wins anchor semantics are inherited, kv-data.go:413-434)
//   genomes: 2-bit payloads concatenated (16-byte aligned each), per-genome offsets / lengths / contig sizes
#pragma once
#include "common.cuh"
#include "lmi_format.hpp"
#include <omp.h>
#include <unordered_map>
#include <memory>

struct Image {
  // scalars
  int k = 31, m = 0, mask_prefix = 7, anchor_prefix = 6, NA = 4096, contig_interval = 1000, device = 0; i64 total_bases = 0;
  u64 E = 0, V = 0; int G = 0; size_t bytes = 0;
  // device arrays
  u64 *d_masks = nullptr, *d_bucket_off = nullptr, *d_keys = nullptr, *d_val_off = nullptr, *d_vals = nullptr; u32* d_anchor_start = nullptr;
  u32* d_anchor_bits = nullptr;   // m * NA/32 words: bit a of mask i set iff anchor_start[i][a] is present (10 MB, L2-resident filter in front of the 328 MB table)
  u32* d_mask_pstart = nullptr; int mask_pbits = 14;   // masks bucketed by their mask_prefix leading bases: [pstart[p], pstart[p+1])
  u8* d_g2bit = nullptr; u64* d_g_off = nullptr; u32 *d_g_nbases = nullptr, *d_g_seq_off = nullptr, *d_seq_sizes = nullptr; u32* d_batch_base = nullptr;
  // host metadata
  std::vector<u64> h_masks; std::vector<std::string> genome_names; std::vector<u64> genome_bgi; std::shared_ptr<std::vector<std::vector<std::string>>> seq_ids_p = std::make_shared<std::vector<std::vector<std::string>>>(); std::vector<std::vector<std::string>>& seq_ids = *seq_ids_p; std::vector<std::vector<u32>> seq_sizes; int n_shards = 1, shard = 0;
  std::vector<u8> h_g2bit; std::vector<u64> h_g_off;   // host copy of the 2-bit genomes: alignment text of the -a output
  std::vector<u32> batch_base, h_nbases; std::unordered_map<u64, u32> bgi2dense; lmi::IndexInfo info;
  // split genomes (genomes.chunks.bin): per dense genome its chunk group (0xFFFFFFFF = not split), chunk index and chunk count (lib-index-search.go:504-537)
  bool has_chunks = false; std::vector<u32> chunk_group, chunk_idx, chunk_n;

  template <class T> T* up(const std::vector<T>& h) { T* d = nullptr; size_t b = std::max<size_t>(h.size(), 1) * sizeof(T) + 64; CUDA_CHECK(cudaMalloc((void**)&d, b)); if (!h.empty()) CUDA_CHECK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice)); bytes += b; return d; }

  void load(const std::string& dir, int dev, int shard, int n_shards) {
    device = dev; CUDA_CHECK(cudaSetDevice(dev)); this->n_shards = n_shards; this->shard = shard;
    info = lmi::read_info(dir + "/info.toml"); k = info.k; m = info.masks; contig_interval = info.contig_interval; total_bases = info.input_bases;
    if (k != 31) lmi::die("only k=31 indexes are supported by the GPU path (SeqComparatorOptions.K is fixed to 31, search.go:361)");
    mask_prefix = std::max((int)(std::log2((double)m) / 2), 1); anchor_prefix = std::max((int)(std::log2((double)info.partitions) / 2), 1); NA = 1 << (2 * anchor_prefix);  // lib-index-search.go:467-469
    int kk = 0; h_masks = lmi::read_masks(dir + "/masks.bin", &kk); if ((int)h_masks.size() != m) lmi::die("masks.bin does not match info.toml");
    // ---- genomes
    std::vector<u8> g2bit; std::vector<u64> g_off; std::vector<u32> g_nbases, g_seq_off(1, 0), seqsz; batch_base.assign(info.genome_batches + 1, 0);
    for (int b = 0; b < info.genome_batches; b++) {
      std::vector<lmi::GenomeRec> recs = lmi::read_genomes(lmi::batch_dir(dir, b) + "/genomes.bin"); batch_base[b + 1] = batch_base[b] + (u32)recs.size();
      for (size_t i = 0; i < recs.size(); i++) { lmi::GenomeRec& r = recs[i]; u64 bgi = ((u64)b << 17) | i; bgi2dense[bgi] = (u32)genome_bgi.size(); genome_bgi.push_back(bgi);
        genome_names.push_back(r.id); seq_ids.push_back(r.seq_ids); seq_sizes.push_back(r.seq_sizes); g_nbases.push_back(r.concat_len);
        for (u32 s : r.seq_sizes) seqsz.push_back(s); g_seq_off.push_back((u32)seqsz.size());
        while (g2bit.size() & 15) g2bit.push_back(0); g_off.push_back(g2bit.size()); g2bit.insert(g2bit.end(), r.twobit.begin(), r.twobit.end()); }
    }
    { auto gm = lmi::read_genome_map(dir + "/genomes.map.bin"); for (auto& e : gm) { auto it = bgi2dense.find(e.second); if (it != bgi2dense.end()) genome_names[it->second] = e.first; } }
    g2bit.resize(g2bit.size() + 64, 0); g_off.push_back(g2bit.size()); G = (int)genome_bgi.size();
    { auto gc = lmi::read_genome_chunks(dir + "/genomes.chunks.bin"); chunk_group.assign(G, 0xFFFFFFFFu); chunk_idx.assign(G, 0); chunk_n.assign(G, 1);
      for (size_t gr = 0; gr < gc.size(); gr++) for (size_t i = 0; i < gc[gr].size(); i++) { auto it = bgi2dense.find(gc[gr][i]); if (it == bgi2dense.end()) lmi::die("genomes.chunks.bin names a genome that is not in the index"); has_chunks = true; chunk_group[it->second] = (u32)gr; chunk_idx[it->second] = (u32)i; chunk_n[it->second] = (u32)gc[gr].size(); } }
    h_nbases = g_nbases; h_g_off = g_off; h_g2bit = g2bit; d_g2bit = up(g2bit); d_g_off = up(g_off); d_g_nbases = up(g_nbases); d_g_seq_off = up(g_seq_off); d_seq_sizes = up(seqsz); d_batch_base = up(batch_base);
    // ---- seeds: decode every chunk (host, one thread per chunk), flatten
    std::vector<lmi::KvChunk> chunks(info.chunks);
#pragma omp parallel for schedule(dynamic, 1)
    for (int c = 0; c < info.chunks; c++) chunks[c] = lmi::read_kv_chunk(lmi::chunk_file(dir, c));
    std::vector<u64> bucket_off(m + 1, 0); std::vector<u64> vcount(m + 1, 0);
    for (auto& c : chunks) { if (c.mask_prefix != mask_prefix || c.anchor_prefix != anchor_prefix) lmi::die("seed chunk prefix lengths do not match info.toml");
      for (int j = 0; j < c.chunk_size; j++) { bucket_off[c.mask_offset + j + 1] = c.masks[j].keys.size(); vcount[c.mask_offset + j + 1] = c.masks[j].vals.size(); } }
    for (int j = 0; j < m; j++) { bucket_off[j + 1] += bucket_off[j]; vcount[j + 1] += vcount[j]; }
    E = bucket_off[m]; V = vcount[m];
    std::vector<u64> keys(E), val_off(E + 1), vals(V); std::vector<u32> anchor_start((size_t)m * NA, 0xFFFFFFFFu);
    for (auto& c : chunks) {
#pragma omp parallel for schedule(dynamic, 16)
      for (int j = 0; j < c.chunk_size; j++) { const lmi::KvMaskData& md = c.masks[j]; int gm_ = c.mask_offset + j; u64 e0 = bucket_off[gm_], v0 = vcount[gm_];
        for (size_t t = 0; t < md.keys.size(); t++) { keys[e0 + t] = md.keys[t]; val_off[e0 + t] = v0 + md.val_off[t]; }
        for (size_t t = 0; t < md.vals.size(); t++) vals[v0 + t] = md.vals[t];
        memcpy(&anchor_start[(size_t)gm_ * NA], c.anchor_start[j].data(), (size_t)NA * 4); }
    }
    val_off[E] = V; chunks.clear();
    if (n_shards > 1) {  // genome sharding: keep only values of genomes with dense % n_shards == shard; keys keep their slots (flag semantics unchanged for kept order)
      std::vector<u64> nv; nv.reserve(V / n_shards + 16); std::vector<u64> no(E + 1);
      for (u64 e = 0; e < E; e++) { no[e] = nv.size(); for (u64 t = val_off[e]; t < val_off[e + 1]; t++) { auto it = bgi2dense.find(vals[t] >> 30); if (it != bgi2dense.end() && (int)(it->second % n_shards) == shard) nv.push_back(vals[t]); } }
      no[E] = nv.size(); vals.swap(nv); val_off.swap(no); V = vals.size();
    }
    { mask_pbits = 2 * mask_prefix; std::vector<u32> ps(((size_t)1 << mask_pbits) + 1, 0); for (u64 mk : h_masks) ps[(mk >> (2 * k - mask_pbits)) + 1]++; for (size_t i = 0; i + 1 < ps.size(); i++) ps[i + 1] += ps[i]; d_mask_pstart = up(ps); }
    { std::vector<u32> bits((size_t)m * NA / 32 + 1, 0); for (size_t t = 0; t < anchor_start.size(); t++) if (anchor_start[t] != 0xFFFFFFFFu) bits[t >> 5] |= 1u << (t & 31); d_anchor_bits = up(bits); }
    d_masks = up(h_masks); d_bucket_off = up(bucket_off); d_keys = up(keys); d_val_off = up(val_off); d_vals = up(vals); d_anchor_start = up(anchor_start);
  }
  void release() { for (void* p : {(void*)d_masks, (void*)d_bucket_off, (void*)d_keys, (void*)d_val_off, (void*)d_vals, (void*)d_anchor_start, (void*)d_g2bit, (void*)d_g_off, (void*)d_g_nbases, (void*)d_g_seq_off, (void*)d_seq_sizes, (void*)d_batch_base, (void*)d_mask_pstart, (void*)d_anchor_bits}) if (p) cudaFree(p); }
};
