// image.cuh — the GPU-resident index image (B200: everything lives in HBM) and its ingest from the on-disk `.lmi` format.
//
// Replaces the reference's per-query file seeks + VARINT-GB decode (kv-searcher.go:366-395) and genome file reads
// (genome.go:1047-1062) with flat arrays:
//   seeds:   entries[E]   16-byte records {k-mer, first value (bucket-relative), value count | first-value reversed flag}, sorted by
//                         k-mer inside each mask bucket — one 32-byte sector answers "key, has values of the wanted kind, how many"
//            vals[V]      seed values in file order; bucket_off[m+1] / bucket_voff[m+1] = first entry / value of every mask bucket
//            anchor_start[m * 4^anchorPrefix] (u32 bucket-relative; from the .idx files so the writer's last-run-wins anchor
//                         semantics are inherited, kv-data.go:413-434) + a 1-bit-per-anchor presence bitmap
//   genomes: 2-bit payloads concatenated (16-byte aligned each), per-genome offsets / lengths / contig sizes
//
// Ingest: the seed chunk files are copied to the GPU as they are and decoded THERE — one thread per mask walks its
// VARINT-GB record stream (kv-data.go:328-602) twice: a counting pass (bucket sizes -> offsets) and a fill pass; a third kernel resolves the
// .idx anchor records to entry indexes. The host only reads files and walks the 8-byte record counts of the .idx blocks; it never holds
// more than one chunk file, so host RAM stays at ~1/16 of the index instead of 2x the image.
#pragma once
#include "common.cuh"
#include "lmi_format.hpp"
#include <omp.h>
#include <unordered_map>
#include <memory>
#include <chrono>
#include <sys/stat.h>

struct SeedEntry { u64 key; u32 vrel; u32 nflag; };   // nflag = number of values | (reversed flag of the key's FIRST value in the whole index) << 31
static_assert(sizeof(SeedEntry) == 16, "SeedEntry must be 16 bytes");

// Prefix Bloom filter of the seed index: one bit per hashed (mask bucket, leading maskPrefix+anchorPrefix bases of a stored k-mer). A probe can only
// match keys that share its first p >= maskPrefix+anchorPrefix bases (kv-searcher.go:202,282-304), so "no key of the bucket starts like the probe" is
// an exact reason to drop it before the index lookup; the anchor table alone (which ignores the mask prefix, kv-data.go:319-325) lets ~30 % of the
// (query, mask) slots through, most of them with nothing to find.
__host__ __device__ __forceinline__ u32 pb_hash(u32 bucket, u32 prefix) { u32 h = prefix * 0x9E3779B1u ^ (bucket * 0x85EBCA6Bu); h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 13; return h; }

__device__ __forceinline__ u64 ld_be(const u8* __restrict__ p, int n) { u64 v = 0; for (int i = 0; i < n; i++) v = (v << 8) | p[i]; return v; }

struct KvWalk {   // cursor over one mask's record stream
  const u8* d; u64 p; u64 prev; u64 left; int vb;
  // next pair record: keys k1 (and k2 unless single), value counts n1 / n2; returns the file offset of the first value
  __device__ __forceinline__ u64 next(u64& k1, u64& k2, u64& n1, u64& n2, bool& has2) {
    u8 c = d[p++]; const bool last = c & 128, single = c & 64; c &= 63; const int b1 = ((c >> 3) & 7) + 1, b2 = (c & 7) + 1;
    const u64 v1 = ld_be(d + p, b1), v2 = ld_be(d + p + b1, b2); p += b1 + b2; k1 = prev + v1; k2 = k1 + v2; prev = k2;
    c = d[p++]; const int c1 = ((c >> 3) & 7) + 1, c2 = (c & 7) + 1; n1 = ld_be(d + p, c1); n2 = ld_be(d + p + c1, c2); p += c1 + c2;
    has2 = !(last && single) && left >= 2; if (!has2) n2 = 0; const u64 v0 = p; p += (n1 + n2) * (u64)vb; left -= has2 ? 2 : 1; return v0; }
};
struct ShardSel { const u32* batch_base; int n_shards, shard; __device__ __forceinline__ bool keep(u64 v) const { if (n_shards <= 1) return true; const u64 bgi = v >> 30; const u32 g = batch_base[bgi >> 17] + (u32)(bgi & 0x1ffff); return (int)(g % (u32)n_shards) == shard; } };

// counting pass: one thread per mask of the chunk
__global__ void k_kv_count(const u8* __restrict__ d, const u8* __restrict__ x, const u64* __restrict__ xoff, const u32* __restrict__ xn, int nmasks, int vb, ShardSel S, u64* __restrict__ nkeys, u64* __restrict__ nvals) {
  int m = blockIdx.x * blockDim.x + threadIdx.x; if (m >= nmasks) return; if (xn[m] == 0) { nkeys[m] = 0; nvals[m] = 0; return; }
  const u64 rec0 = ld_be(x + xoff[m] + 8, 8) >> 1; const u64 nk = ld_be(d + rec0 - 8, 8); KvWalk w{d, rec0, 0, nk, vb}; u64 nv = 0;
  while (w.left) { u64 k1, k2, n1, n2; bool h2; const u64 v0 = w.next(k1, k2, n1, n2, h2); if (S.n_shards <= 1) nv += n1 + n2; else for (u64 i = 0; i < n1 + n2; i++) nv += S.keep(ld_be(d + v0 + i * vb, vb)) ? 1 : 0; }
  nkeys[m] = nk; nvals[m] = nv;
}
// fill pass: entries + values of every mask of the chunk
__global__ void k_kv_fill(const u8* __restrict__ d, const u8* __restrict__ x, const u64* __restrict__ xoff, const u32* __restrict__ xn, int nmasks, int vb, ShardSel S, const u64* __restrict__ bucket_off, const u64* __restrict__ bucket_voff,
                          SeedEntry* __restrict__ entries, u64* __restrict__ vals, u32* __restrict__ pbloom, u32 pbmask, int ash, int mask0) {
  int m = blockIdx.x * blockDim.x + threadIdx.x; if (m >= nmasks || xn[m] == 0) return;
  const u64 rec0 = ld_be(x + xoff[m] + 8, 8) >> 1; const u64 nk = ld_be(d + rec0 - 8, 8); KvWalk w{d, rec0, 0, nk, vb}; SeedEntry* E = entries + bucket_off[m]; u64* V = vals + bucket_voff[m]; u64 e = 0, v = 0;
  auto put = [&](u64 key, u64 v0, u64 n) { SeedEntry t; t.key = key; t.vrel = (u32)v; u32 cnt = 0, flag = 0; for (u64 i = 0; i < n; i++) { const u64 val = ld_be(d + v0 + i * vb, vb); if (i == 0) flag = (u32)(val & 1); if (S.keep(val)) { V[v++] = val; cnt++; } } t.nflag = cnt | (flag << 31); E[e++] = t; const u32 hb = pb_hash((u32)(mask0 + m), (u32)(key >> ash)) & pbmask; atomicOr(&pbloom[hb >> 5], 1u << (hb & 31)); };
  while (w.left) { u64 k1, k2, n1, n2; bool h2; const u64 v0 = w.next(k1, k2, n1, n2, h2); put(k1, v0, n1); if (h2) put(k2, v0 + n1 * (u64)vb, n2); }
}
// anchor records of the .idx block -> bucket-relative entry index (the recorded k-mer is the key the scan starts at). One warp per mask.
__global__ void k_kv_anchor(const u8* __restrict__ x, const u64* __restrict__ xoff, const u32* __restrict__ xn, int nmasks, int mask0, const u64* __restrict__ bucket_off, const SeedEntry* __restrict__ entries, int sh, u32 NA, u32* __restrict__ anchor_start, u32* __restrict__ bad) {
  const int wm = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31; if (wm >= nmasks) return; const u32 nrec = xn[wm]; const u64 b0 = bucket_off[mask0 + wm], b1 = bucket_off[mask0 + wm + 1]; const SeedEntry* E = entries + b0; const u32 n = (u32)(b1 - b0);
  for (u32 r = 1 + lane; r < nrec; r += 32) { const u64 kmer = ld_be(x + xoff[wm] + 16ull * r, 8); u32 lo = 0, hi = n; while (lo < hi) { u32 mid = (lo + hi) >> 1; if (E[mid].key < kmer) lo = mid + 1; else hi = mid; }
    if (lo >= n || E[lo].key != kmer) { atomicAdd(bad, 1u); continue; } anchor_start[(u64)wm * NA + (u32)((kmer >> sh) & (u64)(NA - 1))] = lo; }   // dense scratch table of the chunk
}
// Compact anchor table. A bucket uses a fraction of its 4^anchorPrefix anchors (~1,150 of 4,096 on the 1,000-genome benchmark index), so the dense table
// (m x 4096 x 4 B = 328 MB) is replaced by three smaller arrays: the presence bitmap (1 bit per anchor), per 32-anchor word the number of present anchors
// before it in the bucket (u16), and the starts of the present anchors only (bucket b: cstart[cbase[b] ...]); start = cstart[cbase[b] + cum[word] + popc(bits below)].
// One CTA per bucket of the chunk: `full` is the chunk's dense scratch table (0xFFFFFFFF = absent).
__global__ void __launch_bounds__(128) k_anchor_compact(const u32* __restrict__ full, int nm, int mask0, u32 NA, const u32* __restrict__ cbase, u32* __restrict__ bits, u16* __restrict__ cum, u32* __restrict__ cstart, u32* __restrict__ bad) {
  __shared__ u32 s_cnt[128]; __shared__ u32 s_run; const int b = blockIdx.x; if (b >= nm) return; const u32 nw = NA >> 5; const u32* F = full + (u64)b * NA; const u64 w0 = (u64)(mask0 + b) * nw;
  for (u32 wbase = 0; wbase < nw; wbase += 128) { const u32 w = wbase + threadIdx.x; u32 v = 0; if (w < nw) for (int i = 0; i < 32; i++) if (F[w * 32 + i] != 0xFFFFFFFFu) v |= 1u << i; if (w < nw) bits[w0 + w] = v; s_cnt[threadIdx.x] = (w < nw) ? __popc(v) : 0; __syncthreads();
    if (wbase == 0 && threadIdx.x == 0) s_run = 0; __syncthreads();
    if (threadIdx.x == 0) { u32 run = s_run; for (u32 i = 0; i < 128; i++) { const u32 c = s_cnt[i]; s_cnt[i] = run; run += c; } s_run = run; } __syncthreads();
    if (w < nw) { const u32 before = s_cnt[threadIdx.x]; cum[w0 + w] = (u16)before; u32 r = 0; const u32 base = cbase[mask0 + b]; u32 vv = v; while (vv) { const int i = __ffs(vv) - 1; vv &= vv - 1; cstart[base + before + r] = F[w * 32 + i]; r++; } } __syncthreads(); }
  (void)bad;
}
__global__ void k_cstart_narrow(const u32* __restrict__ in, u16* __restrict__ out, u64 n) { const u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; if (i < n) out[i] = (u16)in[i]; }
__global__ void k_anchor_count(const u32* __restrict__ full, int nm, u32 NA, u32* __restrict__ counts) { const int b = blockIdx.x * blockDim.x + threadIdx.x; if (b >= nm) return; const u32* F = full + (u64)b * NA; u32 c = 0; for (u32 i = 0; i < NA; i++) c += F[i] != 0xFFFFFFFFu; counts[b] = c; }
__global__ void k_anchor_verify(const u32* __restrict__ bits, const u32* __restrict__ cbase, int m, u32 nw, u32* __restrict__ bad) { const int b = blockIdx.x * blockDim.x + threadIdx.x; if (b >= m) return; u32 c = 0; for (u32 w = 0; w < nw; w++) c += __popc(bits[(u64)b * nw + w]); if (c != cbase[b + 1] - cbase[b]) atomicAdd(bad, 1u); }

// ---- synthetic seed image (BASELINE.json configs[4]): per mask `per` keys = the mask's prefix + stratified-uniform low bits (sorted
// by construction), one random value each (reversed flag = bit 0 of a hash)
__host__ __device__ inline u64 mix64(u64 z) { z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
__host__ __device__ inline u64 synth_key(u64 mask, int mask_prefix, int k, u64 per, u64 j, u64 seed, u32 bucket) { const int lowbits = 2 * (k - mask_prefix); const u64 span = 1ull << lowbits, stride = span / per; const u64 r = j * stride + mix64(seed ^ ((u64)bucket << 32) ^ j) % stride; return ((mask >> lowbits) << lowbits) | r; }
__global__ void k_synth_fill(const u64* __restrict__ masks, int m0, int nm, int mask_prefix, int k, u64 per, u64 seed, SeedEntry* __restrict__ entries, u64* __restrict__ vals, u32* __restrict__ pbloom, u32 pbmask, int ash) {
  u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x; if (t >= (u64)nm * per) return; const u32 b = (u32)(t / per); const u64 j = t % per; SeedEntry e; e.key = synth_key(masks[m0 + b], mask_prefix, k, per, j, seed, m0 + b); e.vrel = (u32)j;
  { const u32 hb = pb_hash((u32)(m0 + b), (u32)(e.key >> ash)) & pbmask; atomicOr(&pbloom[hb >> 5], 1u << (hb & 31)); }
  const u64 v = mix64(seed * 31 + ((u64)(m0 + b) * per + j)); e.nflag = 1u | ((u32)(v & 1) << 31); entries[t] = e; if (vals) vals[t] = v & ~(0x1FFFFull << 47);   // batch bits cleared: genome = bits 30..46 only
}
__global__ void k_synth_anchor(const u64* __restrict__ bucket_off, const SeedEntry* __restrict__ entries, int nm, int sh, u32 NA, u32* __restrict__ anchor_start) {   // thread per (mask, anchor): first entry with that anchor
  u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x; if (t >= (u64)nm * NA) return; const u32 b = (u32)(t / NA), a = (u32)(t % NA); const SeedEntry* E = entries + bucket_off[b]; const u32 n = (u32)(bucket_off[b + 1] - bucket_off[b]); if (!n) { anchor_start[t] = 0xFFFFFFFFu; return; }
  const int ab = __ffs((int)NA) - 1; const u64 want = ((E[0].key >> (sh + ab)) << (sh + ab)) | ((u64)a << sh);   // every key of a synthetic bucket carries the mask's prefix
  u32 lo = 0, hi = n; while (lo < hi) { u32 mid = (lo + hi) >> 1; if (E[mid].key < want) lo = mid + 1; else hi = mid; }
  anchor_start[t] = (lo < n && (E[lo].key >> sh) == (want >> sh)) ? lo : 0xFFFFFFFFu;
}

struct Image {
  // scalars
  int k = 31, m = 0, mask_prefix = 7, anchor_prefix = 6, NA = 4096, contig_interval = 1000, device = 0; i64 total_bases = 0;
  u64 E = 0, V = 0; int G = 0; size_t bytes = 0; int n_shards = 1, shard = 0;
  int mask_lo = 0, mask_hi = 0;   // masks whose buckets this image holds ([0, m) except for mask-range-partitioned synthetic images)
  // device arrays
  u64 *d_masks = nullptr, *d_bucket_off = nullptr, *d_bucket_voff = nullptr, *d_vals = nullptr; SeedEntry* d_entries = nullptr;
  u32 *d_anchor_cbase = nullptr, *d_anchor_cstart = nullptr; u16* d_anchor_cum = nullptr; u64 n_anchors = 0;   // compact anchor table (k_anchor_compact)
  u16* d_anchor_cstart16 = nullptr;   // the starts as u16 when every bucket holds < 65,536 keys (halves the table the lookup kernel wants in L2); d_anchor_cstart is released then
  u32* d_pbloom = nullptr; u32 pbmask = 0;   // prefix Bloom filter (pb_hash): pbmask + 1 bits, a power of two >= 8 bits per stored k-mer (<= 2^32)
  u32* d_anchor_bits = nullptr;   // m * NA/32 words: bit a of mask i set iff anchor_start[i][a] is present (10 MB, L2-resident filter in front of the 328 MB table)
  u32* d_mask_pstart = nullptr; int mask_pbits = 14;   // masks bucketed by their mask_prefix leading bases: [pstart[p], pstart[p+1])
  u8* d_g2bit = nullptr; u64* d_g_off = nullptr; u32 *d_g_nbases = nullptr, *d_g_seq_off = nullptr, *d_seq_sizes = nullptr; u32* d_batch_base = nullptr;
  // host metadata
  std::vector<u64> h_masks; std::vector<std::string> genome_names; std::vector<u64> genome_bgi; std::shared_ptr<std::vector<std::vector<std::string>>> seq_ids_p = std::make_shared<std::vector<std::vector<std::string>>>(); std::vector<std::vector<std::string>>& seq_ids = *seq_ids_p; std::vector<std::vector<u32>> seq_sizes;
  std::vector<u8> h_g2bit; std::vector<u64> h_g_off;   // host copy of the 2-bit genomes: alignment text of the -a output
  std::vector<u32> batch_base, h_nbases; std::unordered_map<u64, u32> bgi2dense; lmi::IndexInfo info;
  // split genomes (genomes.chunks.bin): per dense genome its chunk group (0xFFFFFFFF = not split), chunk index and chunk count (lib-index-search.go:504-537)
  bool has_chunks = false; std::vector<u32> chunk_group, chunk_idx, chunk_n;
  u64 synth_per = 0, synth_seed = 0;   // > 0: synthetic seeds-only image (lmg_index_synth)
  double load_ms[4] = {0, 0, 0, 0};   // genomes, seed count pass, seed fill pass, total (wall)

  template <class T> T* dalloc(size_t n) { T* d = nullptr; size_t b = std::max<size_t>(n, 1) * sizeof(T) + 64; CUDA_CHECK(cudaMalloc((void**)&d, b)); bytes += b; return d; }
  template <class T> T* up(const std::vector<T>& h) { T* d = dalloc<T>(h.size()); if (!h.empty()) CUDA_CHECK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice)); return d; }

  void alloc_pbloom() { u64 bits = 1ull << 20; while (bits < 8 * E && bits < (1ull << 32)) bits <<= 1; pbmask = (u32)(bits - 1); d_pbloom = dalloc<u32>(bits / 32); CUDA_CHECK(cudaMemset(d_pbloom, 0, bits / 8)); }
  void finish_masks() { mask_pbits = 2 * mask_prefix; std::vector<u32> ps(((size_t)1 << mask_pbits) + 1, 0); for (u64 mk : h_masks) ps[(mk >> (2 * k - mask_pbits)) + 1]++; for (size_t i = 0; i + 1 < ps.size(); i++) ps[i + 1] += ps[i]; d_mask_pstart = up(ps); d_masks = up(h_masks); }
  // compact anchor arrays from per-bucket counts (host): allocations + prefix sums; the chunks are compacted afterwards
  void alloc_anchors(const std::vector<u32>& counts) { if (NA < 32) lmi::die("indexes with fewer than 64 partitions are not supported by the GPU path"); std::vector<u32> cb(m + 1, 0); for (int j = 0; j < m; j++) cb[j + 1] = cb[j] + counts[j]; n_anchors = cb[m];
    d_anchor_cbase = up(cb); d_anchor_cstart = dalloc<u32>(n_anchors); const u64 nwords = (u64)m * (NA >> 5); d_anchor_bits = dalloc<u32>(nwords); d_anchor_cum = dalloc<u16>(nwords); CUDA_CHECK(cudaMemset(d_anchor_bits, 0, nwords * 4)); CUDA_CHECK(cudaMemset(d_anchor_cum, 0, nwords * 2)); }
  // starts are bucket-relative key indexes: 16 bits are enough unless a bucket holds 65,536 keys or more (LMG_CSTART32 keeps the wide table)
  void pack_cstart16(const std::vector<u64>& bucket_off) { u64 mx = 0; for (int j = 0; j < m; j++) mx = std::max(mx, bucket_off[j + 1] - bucket_off[j]); if (mx >= 65536 || n_anchors == 0 || getenv("LMG_CSTART32")) return;
    u16* d16 = nullptr; CUDA_CHECK(cudaMalloc((void**)&d16, n_anchors * 2 + 64)); k_cstart_narrow<<<(unsigned)((n_anchors + 255) / 256), 256>>>(d_anchor_cstart, d16, n_anchors); CUDA_CHECK(cudaGetLastError()); CUDA_CHECK(cudaDeviceSynchronize());
    CUDA_CHECK(cudaFree(d_anchor_cstart)); d_anchor_cstart = nullptr; d_anchor_cstart16 = d16; bytes -= n_anchors * 2; }
  void verify_anchors() { u32* d_bad = nullptr; CUDA_CHECK(cudaMalloc((void**)&d_bad, 4)); CUDA_CHECK(cudaMemset(d_bad, 0, 4)); k_anchor_verify<<<(m + 127) / 128, 128>>>(d_anchor_bits, d_anchor_cbase, m, (u32)(NA >> 5), d_bad); CUDA_CHECK(cudaGetLastError()); u32 bad = 0; CUDA_CHECK(cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost)); cudaFree(d_bad);
    if (bad) lmi::die("kv-index: the anchor records of " + std::to_string(bad) + " masks are inconsistent (duplicate or missing anchors)"); }

  void load(const std::string& dir, int dev, int shard_, int n_shards_) {
    auto t_all = std::chrono::steady_clock::now(); auto ms_since = [](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t).count(); };
    device = dev; CUDA_CHECK(cudaSetDevice(dev)); n_shards = std::max(1, n_shards_); shard = shard_;
    info = lmi::read_info(dir + "/info.toml"); k = info.k; m = info.masks; contig_interval = info.contig_interval; total_bases = info.input_bases; mask_lo = 0; mask_hi = m;
    if (k != 31) lmi::die("only k=31 indexes are supported by the GPU path (SeqComparatorOptions.K is fixed to 31, search.go:361)");
    mask_prefix = std::max((int)(std::log2((double)m) / 2), 1); anchor_prefix = std::max((int)(std::log2((double)info.partitions) / 2), 1); NA = 1 << (2 * anchor_prefix);  // lib-index-search.go:467-469
    int kk = 0; h_masks = lmi::read_masks(dir + "/masks.bin", &kk); if ((int)h_masks.size() != m) lmi::die("masks.bin does not match info.toml");
    // ---- genomes: sizes from the .idx files, payloads streamed batch by batch into one device array (and the host copy used for -a text)
    auto t0 = std::chrono::steady_clock::now();
    std::vector<u64> g_off; std::vector<u32> g_nbases, g_seq_off(1, 0), seqsz; batch_base.assign(info.genome_batches + 1, 0);
    { u64 tot = 0; std::vector<std::vector<std::pair<u64, u32>>> bidx(info.genome_batches);
      for (int b = 0; b < info.genome_batches; b++) { std::vector<u8> x = lmi::read_file(lmi::batch_dir(dir, b) + "/genomes.bin.idx"); if (x.size() < 24 || memcmp(x.data(), ".genomei", 8)) lmi::die("not a genome index file");
        u32 n = (u32)lmi::get_be(&x[20], 4); batch_base[b + 1] = batch_base[b] + n; for (u32 i = 0; i < n; i++) { bidx[b].push_back({lmi::get_be(&x[24 + 12 * i], 8), (u32)lmi::get_be(&x[32 + 12 * i], 4)}); tot = (tot + 15) & ~15ull; g_off.push_back(tot); tot += ((u64)bidx[b].back().second + 3) / 4; } }
      tot += 64; g_off.push_back(tot); G = (int)batch_base[info.genome_batches]; h_g2bit.assign(tot, 0); d_g2bit = dalloc<u8>(tot); CUDA_CHECK(cudaMemset(d_g2bit, 0, tot));
      u32 gd = 0;
      for (int b = 0; b < info.genome_batches; b++) { std::vector<u8> d = lmi::read_file(lmi::batch_dir(dir, b) + "/genomes.bin"); if (d.size() < 16 || memcmp(d.data(), ".genomes", 8)) lmi::die("not a genome data file");
        const u64 first = g_off[gd];
        for (size_t i = 0; i < bidx[b].size(); i++, gd++) { size_t p = bidx[b][i].first; u64 bgi = ((u64)b << 17) | i; bgi2dense[bgi] = gd; genome_bgi.push_back(bgi);
          size_t l = lmi::get_be(&d[p], 2); p += 2; genome_names.emplace_back((const char*)&d[p], l); p += l; u32 concat = (u32)lmi::get_be(&d[p + 4], 4), ns = (u32)lmi::get_be(&d[p + 8], 4); p += 12;
          seq_ids.emplace_back(); seq_sizes.emplace_back(); for (u32 s = 0; s < ns; s++) { u32 sz = (u32)lmi::get_be(&d[p], 4); l = lmi::get_be(&d[p + 4], 2); p += 6; seq_ids.back().emplace_back((const char*)&d[p], l); p += l; seq_sizes.back().push_back(sz); seqsz.push_back(sz); }
          g_seq_off.push_back((u32)seqsz.size()); g_nbases.push_back(concat); size_t nb = lmi::get_be(&d[p], 4); p += 8; if (p + nb > d.size() || nb > g_off[gd + 1] - g_off[gd]) lmi::die("genomes.bin: broken record"); memcpy(&h_g2bit[g_off[gd]], &d[p], nb); }
        if (gd > 0 && !bidx[b].empty()) CUDA_CHECK(cudaMemcpy(d_g2bit + first, &h_g2bit[first], g_off[gd] - first, cudaMemcpyHostToDevice)); } }
    { auto gm = lmi::read_genome_map(dir + "/genomes.map.bin"); for (auto& e : gm) { auto it = bgi2dense.find(e.second); if (it != bgi2dense.end()) genome_names[it->second] = e.first; } }
    { auto gc = lmi::read_genome_chunks(dir + "/genomes.chunks.bin"); chunk_group.assign(G, 0xFFFFFFFFu); chunk_idx.assign(G, 0); chunk_n.assign(G, 1);
      for (size_t gr = 0; gr < gc.size(); gr++) for (size_t i = 0; i < gc[gr].size(); i++) { auto it = bgi2dense.find(gc[gr][i]); if (it == bgi2dense.end()) lmi::die("genomes.chunks.bin names a genome that is not in the index"); has_chunks = true; chunk_group[it->second] = (u32)gr; chunk_idx[it->second] = (u32)i; chunk_n[it->second] = (u32)gc[gr].size(); } }
    h_nbases = g_nbases; h_g_off = g_off; d_g_off = up(g_off); d_g_nbases = up(g_nbases); d_g_seq_off = up(g_seq_off); d_seq_sizes = up(seqsz); d_batch_base = up(batch_base);
    load_ms[0] = ms_since(t0);
    // ---- seeds: chunk files -> device, decoded there
    t0 = std::chrono::steady_clock::now();
    struct ChunkDev { u8 *d = nullptr, *x = nullptr; u64 *xoff = nullptr; u32* xn = nullptr; int mask0 = 0, nm = 0, vb = 7; size_t dsz = 0, xsz = 0; std::vector<u32> h_xn; };
    auto free_chunk = [](ChunkDev& c) { for (void* p : {(void*)c.d, (void*)c.x, (void*)c.xoff, (void*)c.xn}) if (p) cudaFree(p); c.d = c.x = nullptr; c.xoff = nullptr; c.xn = nullptr; };
    auto upload_chunk = [&](int ci) { ChunkDev c; std::string f = lmi::chunk_file(dir, ci); std::vector<u8> d = lmi::read_file(f), x = lmi::read_file(f + ".idx");
      if (d.size() < 32 || memcmp(d.data(), ".kv-data", 8)) lmi::die("not a kv-data file: " + f); if (x.size() < 32 || memcmp(x.data(), ".kvindex", 8)) lmi::die("not a kv-index file: " + f + ".idx"); if (d[8] != 1 || x[8] != 1) lmi::die("kv-data: version mismatch");
      if (d[10] != k) lmi::die("seed chunk k-mer size does not match info.toml"); c.vb = (d[11] & 1) ? 7 : 8; c.mask0 = (int)lmi::get_be(&d[16], 8); c.nm = (int)lmi::get_be(&d[24], 8); if (x[11] != mask_prefix || x[12] != anchor_prefix) lmi::die("seed chunk prefix lengths do not match info.toml");
      if (c.mask0 < 0 || c.mask0 + c.nm > m) lmi::die("seed chunk mask range outside the index");
      std::vector<u64> xoff(c.nm); std::vector<u32> xn(c.nm); size_t q = 32; for (int i = 0; i < c.nm; i++) { if (q + 8 > x.size()) lmi::die("kv-index: truncated"); u64 nrec = lmi::get_be(&x[q], 8); xoff[i] = q + 8; xn[i] = (u32)nrec; q += 8 + 16 * nrec; } if (q > x.size()) lmi::die("kv-index: truncated");
      // xoff = first 16-byte record of the mask's block: record 0 is (record count, file offset of the first k-mer << 1), record r >= 1 is (anchor k-mer, offset << 1 | second-of-pair) (kv-data.go:566-599)
      c.dsz = d.size(); c.xsz = x.size(); CUDA_CHECK(cudaMalloc((void**)&c.d, d.size() + 64)); CUDA_CHECK(cudaMalloc((void**)&c.x, x.size() + 64)); CUDA_CHECK(cudaMalloc((void**)&c.xoff, sizeof(u64) * (c.nm + 1))); CUDA_CHECK(cudaMalloc((void**)&c.xn, sizeof(u32) * (c.nm + 1)));
      CUDA_CHECK(cudaMemcpy(c.d, d.data(), d.size(), cudaMemcpyHostToDevice)); CUDA_CHECK(cudaMemcpy(c.x, x.data(), x.size(), cudaMemcpyHostToDevice)); CUDA_CHECK(cudaMemcpy(c.xoff, xoff.data(), sizeof(u64) * c.nm, cudaMemcpyHostToDevice)); CUDA_CHECK(cudaMemcpy(c.xn, xn.data(), sizeof(u32) * c.nm, cudaMemcpyHostToDevice)); c.h_xn = xn; return c; };
    // keep the raw chunk bytes on the device between the two passes when they are small next to the free memory, otherwise read the files twice
    u64 raw_total = 0; for (int c = 0; c < info.chunks; c++) { struct stat st; std::string f = lmi::chunk_file(dir, c); if (stat(f.c_str(), &st) == 0) raw_total += (u64)st.st_size; if (stat((f + ".idx").c_str(), &st) == 0) raw_total += (u64)st.st_size; }
    size_t freeb = 0, totalb = 0; CUDA_CHECK(cudaMemGetInfo(&freeb, &totalb)); const bool keep_raw = raw_total * 4 < freeb && !getenv("LMG_INGEST_REREAD");
    ShardSel S{d_batch_base, n_shards, shard}; u64 *d_nk = nullptr, *d_nv = nullptr; CUDA_CHECK(cudaMalloc((void**)&d_nk, sizeof(u64) * (m + 1))); CUDA_CHECK(cudaMalloc((void**)&d_nv, sizeof(u64) * (m + 1))); CUDA_CHECK(cudaMemset(d_nk, 0, sizeof(u64) * (m + 1))); CUDA_CHECK(cudaMemset(d_nv, 0, sizeof(u64) * (m + 1)));
    std::vector<ChunkDev> kept(info.chunks); std::vector<u32> anchor_counts(m, 0);
    for (int c = 0; c < info.chunks; c++) { ChunkDev cd = upload_chunk(c); for (int i = 0; i < cd.nm; i++) anchor_counts[cd.mask0 + i] = cd.h_xn[i] ? cd.h_xn[i] - 1 : 0;   /* one .idx record per present anchor after the header record */ if (cd.nm) { k_kv_count<<<(cd.nm + 63) / 64, 64>>>(cd.d, cd.x, cd.xoff, cd.xn, cd.nm, cd.vb, S, d_nk + cd.mask0, d_nv + cd.mask0); CUDA_CHECK(cudaGetLastError()); }
      if (keep_raw) kept[c] = cd; else { CUDA_CHECK(cudaDeviceSynchronize()); free_chunk(cd); } }
    CUDA_CHECK(cudaDeviceSynchronize());
    std::vector<u64> nk(m + 1), nv(m + 1), bucket_off(m + 1, 0), bucket_voff(m + 1, 0); CUDA_CHECK(cudaMemcpy(nk.data(), d_nk, sizeof(u64) * m, cudaMemcpyDeviceToHost)); CUDA_CHECK(cudaMemcpy(nv.data(), d_nv, sizeof(u64) * m, cudaMemcpyDeviceToHost)); cudaFree(d_nk); cudaFree(d_nv);
    for (int j = 0; j < m; j++) { bucket_off[j + 1] = bucket_off[j] + nk[j]; bucket_voff[j + 1] = bucket_voff[j] + nv[j]; if (nk[j] >= (1ull << 32) || nv[j] >= (1ull << 31)) lmi::die("a mask bucket holds more than 2^32 k-mers or 2^31 values"); }
    E = bucket_off[m]; V = bucket_voff[m]; load_ms[1] = ms_since(t0); t0 = std::chrono::steady_clock::now();
    d_bucket_off = up(bucket_off); d_bucket_voff = up(bucket_voff); d_entries = dalloc<SeedEntry>(E); d_vals = dalloc<u64>(V); alloc_anchors(anchor_counts);
    int max_nm = 1; for (int c = 0; c < info.chunks; c++) max_nm = std::max(max_nm, (m + info.chunks - 1) / info.chunks + 1); u32* d_full = nullptr; size_t full_cap = 0;   // dense anchor scratch of one chunk
    alloc_pbloom();
    u32* d_bad = nullptr; CUDA_CHECK(cudaMalloc((void**)&d_bad, 4)); CUDA_CHECK(cudaMemset(d_bad, 0, 4)); const int sh = 2 * (k - mask_prefix - anchor_prefix);
    for (int c = 0; c < info.chunks; c++) { ChunkDev cd = keep_raw ? kept[c] : upload_chunk(c); if (cd.nm) {
        k_kv_fill<<<(cd.nm + 63) / 64, 64>>>(cd.d, cd.x, cd.xoff, cd.xn, cd.nm, cd.vb, S, d_bucket_off + cd.mask0, d_bucket_voff + cd.mask0, d_entries, d_vals, d_pbloom, pbmask, sh, cd.mask0); CUDA_CHECK(cudaGetLastError());
        const size_t need = (size_t)cd.nm * NA; if (need > full_cap) { if (d_full) cudaFree(d_full); CUDA_CHECK(cudaMalloc((void**)&d_full, need * 4)); full_cap = need; } CUDA_CHECK(cudaMemset(d_full, 0xff, need * 4));
        k_kv_anchor<<<(cd.nm * 32 + 127) / 128, 128>>>(cd.x, cd.xoff, cd.xn, cd.nm, cd.mask0, d_bucket_off, d_entries, sh, (u32)NA, d_full, d_bad); CUDA_CHECK(cudaGetLastError());
        k_anchor_compact<<<cd.nm, 128>>>(d_full, cd.nm, cd.mask0, (u32)NA, d_anchor_cbase, d_anchor_bits, d_anchor_cum, d_anchor_cstart, d_bad); CUDA_CHECK(cudaGetLastError()); }
      CUDA_CHECK(cudaDeviceSynchronize()); free_chunk(cd); }
    if (d_full) cudaFree(d_full); (void)max_nm;
    u32 bad = 0; CUDA_CHECK(cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost)); cudaFree(d_bad); if (bad) lmi::die("kv-index: " + std::to_string(bad) + " anchor records do not name a stored k-mer");
    finish_masks(); verify_anchors(); pack_cstart16(bucket_off); load_ms[2] = ms_since(t0); load_ms[3] = ms_since(t_all);
  }

  // Synthetic seeds-only image for the seed-lookup microbenchmark: masks [lo, hi) of an m-mask index, `per` keys each (no genomes: only the probe kernels may run on it)
  void synth(int dev, int m_, u64 per, u64 seed, int lo, int hi, bool with_values) {
    device = dev; CUDA_CHECK(cudaSetDevice(dev)); k = 31; m = m_; mask_prefix = std::max((int)(std::log2((double)m) / 2), 1); anchor_prefix = 6; NA = 4096; mask_lo = lo; mask_hi = hi; G = 0; total_bases = 0;
    h_masks.resize(m); const u64 np = 1ull << (2 * mask_prefix); const int lowbits = 2 * (k - mask_prefix);
    for (int i = 0; i < m; i++) { const u64 p = (u64)i * np / (u64)m; u64 r = mix64(seed ^ (0xC5ull << 56) ^ (u64)i) & ((1ull << lowbits) - 1); const bool dup_lo = ((u64)(i + 1) * np / (u64)m) == p, dup_hi = i > 0 && ((u64)(i - 1) * np / (u64)m) == p;   // masks sharing a prefix differ in the next base
      if (dup_lo || dup_hi) { r &= ~(3ull << (lowbits - 2)); r |= (u64)(dup_hi ? 2 : 1) << (lowbits - 2); } h_masks[i] = (p << lowbits) | r; }
    std::sort(h_masks.begin(), h_masks.end()); for (int i = 1; i < m; i++) if (h_masks[i] == h_masks[i - 1]) lmi::die("synthetic masks collide");
    std::vector<u64> bucket_off(m + 1, 0), bucket_voff(m + 1, 0); for (int j = 0; j < m; j++) { const u64 n = (j >= lo && j < hi) ? per : 0; bucket_off[j + 1] = bucket_off[j] + n; bucket_voff[j + 1] = bucket_voff[j] + (with_values ? n : 0); }
    E = bucket_off[m]; V = bucket_voff[m]; finish_masks(); d_bucket_off = up(bucket_off); d_bucket_voff = up(bucket_voff); d_entries = dalloc<SeedEntry>(E); d_vals = dalloc<u64>(V);
    alloc_pbloom(); const int sh = 2 * (k - mask_prefix - anchor_prefix);
    const int nm = hi - lo;
    if (nm > 0 && per) { const u64 tot = (u64)nm * per; k_synth_fill<<<(unsigned)((tot + 255) / 256), 256>>>(d_masks, lo, nm, mask_prefix, k, per, seed, d_entries, with_values ? d_vals : nullptr, d_pbloom, pbmask, sh); CUDA_CHECK(cudaGetLastError());
      const u64 ta = (u64)nm * NA; u32* d_full = nullptr; CUDA_CHECK(cudaMalloc((void**)&d_full, ta * 4)); k_synth_anchor<<<(unsigned)((ta + 255) / 256), 256>>>(d_bucket_off + lo, d_entries, nm, sh, (u32)NA, d_full); CUDA_CHECK(cudaGetLastError());
      u32* d_cnt = nullptr; CUDA_CHECK(cudaMalloc((void**)&d_cnt, (size_t)nm * 4)); k_anchor_count<<<(nm + 127) / 128, 128>>>(d_full, nm, (u32)NA, d_cnt); CUDA_CHECK(cudaGetLastError()); std::vector<u32> counts(m, 0); CUDA_CHECK(cudaMemcpy(counts.data() + lo, d_cnt, (size_t)nm * 4, cudaMemcpyDeviceToHost)); cudaFree(d_cnt);
      alloc_anchors(counts); u32* d_bad = nullptr; CUDA_CHECK(cudaMalloc((void**)&d_bad, 4)); CUDA_CHECK(cudaMemset(d_bad, 0, 4)); k_anchor_compact<<<nm, 128>>>(d_full, nm, lo, (u32)NA, d_anchor_cbase, d_anchor_bits, d_anchor_cum, d_anchor_cstart, d_bad); CUDA_CHECK(cudaGetLastError()); CUDA_CHECK(cudaDeviceSynchronize()); cudaFree(d_full); cudaFree(d_bad); }
    else alloc_anchors(std::vector<u32>(m, 0));
    verify_anchors(); pack_cstart16(bucket_off);
    batch_base.assign(2, 0); d_batch_base = up(batch_base); std::vector<u64> one(2, 0); d_g_off = up(one); d_g2bit = dalloc<u8>(64);
  }
  void release() { for (void* p : {(void*)d_masks, (void*)d_bucket_off, (void*)d_bucket_voff, (void*)d_entries, (void*)d_vals, (void*)d_anchor_cbase, (void*)d_anchor_cstart, (void*)d_anchor_cstart16, (void*)d_anchor_cum, (void*)d_g2bit, (void*)d_g_off, (void*)d_g_nbases, (void*)d_g_seq_off, (void*)d_seq_sizes, (void*)d_batch_base, (void*)d_mask_pstart, (void*)d_anchor_bits, (void*)d_pbloom}) if (p) cudaFree(p); }
};
