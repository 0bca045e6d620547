// engine.cu — B200 (sm_100a) implementation of LexicMap's query-side search path behind the C ABI in
// include/lexicmap_gpu.h. No CPU fallback: every stage below runs as CUDA kernels on the index image in HBM.
//
// Stage map (reference lexicmap/cmd/lib-index-search.go:1191-2940):
//   K1 sketch      : pack queries to 2-bit, all strand k-mers, per-query sort, per-mask XOR-argmin capture (a1),
//                    DUST filter (a2), base-reversed k-mer -> suffix mask (a3)
//   K2 seed_probe  : prefix + suffix range probes of the mask buckets (a4) and anchor materialisation (a5)
//   K3 chain       : anchor sort, nested-anchor removal (a6), float32 chaining DP + backtrack (a7, a8)
//   K4 pseudo_align: target window fetch (a9), query k-mer table (a10), prefix/suffix matching (a11), Chainer2 (a12)
//   K5 extend+wfa  : contig mapping (a13), 2-mer flank extension (a14), wavefront alignment (a15), scoring (a16, a17)
//   host finish    : per-genome coverage, ordering (a18) -> rows (a19 is the TSV writer in the CLI / Python API)
#include "image.cuh"
#include "../../include/lexicmap_gpu.h"
#include <cub/cub.cuh>
#include <cmath>
#include <cfloat>
#include <algorithm>
#include <numeric>
#include <chrono>
#include <mutex>

// =====================================================================================================
// small device utilities
// =====================================================================================================
#define FULLMASK 0xffffffffu
static inline int cdiv(i64 a, i64 b) { return (int)((a + b - 1) / b); }

struct CubTemp {  // grow-only temp storage for CUB calls
  void* p = nullptr; size_t cap = 0; cudaStream_t st;
  void* get(size_t n) { if (n > cap) { if (p) cudaFreeAsync(p, st); CUDA_CHECK(cudaMallocAsync(&p, n + 256, st)); cap = n; } return p; }
  ~CubTemp() { if (p) cudaFree(p); }
};

// =====================================================================================================
// K1: sketch
// =====================================================================================================
// ASCII -> 2-bit packed (4 bases / byte, first base in bits 7-6). One thread per output byte.
__global__ void k_pack_queries(const u8* __restrict__ ascii, const u64* __restrict__ off, const u64* __restrict__ boff, u8* __restrict__ packed, int nq) {
  int q = blockIdx.y; if (q >= nq) return; u64 a0 = off[q], L = off[q + 1] - a0; u64 nb = (L + 3) >> 2; u8* out = packed + boff[q];
  for (u64 b = blockIdx.x * (u64)blockDim.x + threadIdx.x; b < nb; b += (u64)gridDim.x * blockDim.x) {
    u32 v = 0; for (int j = 0; j < 4; j++) { u64 i = b * 4 + j; u32 c = (i < L) ? base2bit(ascii[a0 + i]) : 0; v = (v << 2) | c; } out[b] = (u8)v; }
}
// all k-mers of both strands: entry 2p = forward k-mer at p, 2p+1 = reverse complement; value = p<<1|strand,
// bit31 = excluded from the pseudo-alignment table (forward k-mer is 0 / homopolymer / DUST; lib-seq_compare.go:143-146)
__global__ void k_gen_kmers(const u8* __restrict__ packed, const u64* __restrict__ boff, const u64* __restrict__ off, const u64* __restrict__ koff, u64* __restrict__ keys, u32* __restrict__ vals, int nq, int k) {
  int q = blockIdx.y; if (q >= nq) return; u64 L = off[q + 1] - off[q]; if (L < (u64)k) return; u64 np = L - k + 1; const u8* s = packed + boff[q]; u64 o = koff[q];
  for (u64 p = blockIdx.x * (u64)blockDim.x + threadIdx.x; p < np; p += (u64)gridDim.x * blockDim.x) {
    u64 f = 0, r = 0; for (int j = 0; j < k; j++) { u64 c = get_base(s, p + j); f = (f << 2) | c; r = (r >> 2) | ((3 - c) << (2 * (k - 1))); }
    u32 ex = (f == 0 || kmer_low_complexity(f, k)) ? 0x80000000u : 0u;
    keys[o + 2 * p] = f; vals[o + 2 * p] = (u32)(p << 1) | ex; keys[o + 2 * p + 1] = r; vals[o + 2 * p + 1] = (u32)(p << 1 | 1) | ex; }
}

// 1D TMA bulk copy global -> shared with mbarrier completion (sm_90+/sm_100a: cp.async.bulk, SASS UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, u32 bytes, u64* mbar) {
  u32 d = (u32)__cvta_generic_to_shared(smem_dst), b = (u32)__cvta_generic_to_shared(mbar);
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(gsrc), "r"(bytes), "r"(b) : "memory");
}
__device__ __forceinline__ void mbar_init(u64* mbar, u32 count) { u32 b = (u32)__cvta_generic_to_shared(mbar); asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(count)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(u64* mbar, u32 phase) {
  u32 b = (u32)__cvta_generic_to_shared(mbar); u32 ok = 0;
  while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(b), "r"(phase) : "memory");
}

// Per (query, mask): the k-mer of either strand minimising kmer XOR mask (= lexichash.MaskKnownDistinctPrefixes(s,nil,true),
// call site lib-index-search.go:1212-1220) found by a bitwise descent over the query's SORTED k-mer table, which one CTA
// stages into shared memory with a single TMA bulk copy (tables of <= smem_cap bytes; longer reads read the table from L2).
// Also: low-complexity filter (:1222-1238), base-reversed k-mer and its suffix mask argmin_j(mask_j XOR rev) (:1322-1341),
// and first-owner dedup of equal captured k-mers (:1288-1298).
struct Capture { u64 kmer; u32 lo, n; u32 smask; };  // kmer==0 -> nothing captured; [lo,lo+n) = rows of the query's table; smask = suffix mask
__global__ void __launch_bounds__(256) k_capture(const u64* __restrict__ qkeys, const u64* __restrict__ koff, const u64* __restrict__ masks, int m, int k, int slices,
                                                 Capture* __restrict__ cap, u32* __restrict__ owner, u32 smem_cap_entries, int use_tma) {
  extern __shared__ __align__(128) u8 smem_raw[]; u64* stab = (u64*)smem_raw; __shared__ __align__(8) u64 mbar;
  int q = blockIdx.x / slices, sl = blockIdx.x % slices; u64 o = koff[q]; u32 n = (u32)(koff[q + 1] - o);
  int per = (m + slices - 1) / slices, i0 = sl * per, i1 = min(m, i0 + per);
  if (n == 0) { for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) { Capture c; c.kmer = 0; c.lo = c.n = 0; c.smask = 0; cap[(u64)q * m + i] = c; } return; }
  const u64* tab = qkeys + o; bool in_smem = n <= smem_cap_entries;
  if (in_smem) {
    if (use_tma) {
      if (threadIdx.x == 0) { mbar_init(&mbar, 1); }
      __syncthreads();
      if (threadIdx.x == 0) { u32 bytes = ((n * 8u) + 15u) & ~15u; tma_load_1d(stab, tab, bytes, &mbar); }   // koff is even -> 16-byte aligned source
      mbar_wait(&mbar, 0);
    } else { for (u32 t = threadIdx.x; t < n; t += blockDim.x) stab[t] = tab[t]; __syncthreads(); }
    tab = stab;
  }
  for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    u64 mk = masks[i]; u32 lo = 0, hi = n; xor_argmin_range(tab, lo, hi, mk);
    u64 km = tab[lo]; Capture c; c.kmer = km; c.lo = lo; c.n = hi - lo; c.smask = 0;
    if (kmer_low_complexity(km, k)) c.kmer = 0;   // km==0 is DUST-low-complexity too, as in the reference
    else { u64 rv = kmer_reverse62(km, k); u32 a = 0, b = (u32)m; xor_argmin_range(masks, a, b, rv); c.smask = a; atomicMin(&owner[o + lo], (u32)i); }
    cap[(u64)q * m + i] = c;
  }
}

// =====================================================================================================
// K2: seed probe (kv.Searcher.Search / Search2, kv/kv-searcher.go:190-1088) + anchors (lib-index-search.go:1357-1569)
// =====================================================================================================
struct ProbeHit { u32 q, mask_dir; u64 e0; u32 ne; u32 lo, n; u64 kmer; u32 nanch; u32 pad; };  // mask_dir = mask<<1 | dir ; [lo,lo+n) = query table rows (locs)
struct ProbeParams { const u64 *bucket_off, *keys, *val_off, *vals; const u32* anchor_start; int m, k, NA, mask_prefix, anchor_prefix, p; };

// one thread per (query, mask, direction). dir 0: captured k-mer against its own mask bucket, values must have reverse flag 0;
// dir 1: base-reversed k-mer against bucket `smask`, reverse flag 1 (decided by the FIRST value of each key: on-disk searcher
// semantics, kv-searcher.go:466-488). Range = keys in [kmer & ~low, kmer | low] at/after the anchor start (:282-304, :349-355).
__global__ void __launch_bounds__(256) k_probe_find(ProbeParams P, const Capture* __restrict__ cap, const u32* __restrict__ owner, const u64* __restrict__ koff, u64 nprobe,
                                                    ProbeHit* __restrict__ hits, u32* __restrict__ nhits, u64* __restrict__ stats) {
  u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x; bool have = false; ProbeHit h; u32 steps = 0;
  if (t < nprobe) {
    u64 qi = t >> 1; int dir = (int)(t & 1); u32 q = (u32)(qi / P.m); int i = (int)(qi % P.m); Capture c = cap[qi];
    if (c.kmer != 0 && !(dir == 1 && owner[koff[q] + c.lo] != (u32)i)) {
      u64 kmer = dir ? kmer_reverse62(c.kmer, P.k) : c.kmer; int bucket = dir ? (int)c.smask : i;
      int s2 = (P.k - P.p) << 1; u64 low = (P.p < P.k) ? ((1ull << s2) - 1) : 0; u64 left = kmer & ~low, right = kmer | low;
      u32 a = (u32)((left >> ((P.k - P.mask_prefix - P.anchor_prefix) << 1)) & (u64)(P.NA - 1));
      u32 as = P.anchor_start[(u64)bucket * P.NA + a];
      if (as != 0xFFFFFFFFu) {
        u64 b0 = P.bucket_off[bucket], b1 = P.bucket_off[bucket + 1]; u64 lo = b0 + as, hi = b1;
        // galloping lower_bound(left) from the anchor start
        u64 step = 1, l = lo; while (l + step < hi && P.keys[l + step] < left) { l += step; step <<= 1; steps++; }
        u64 r = min(hi, l + step); if (P.keys[l] >= left) r = l; else l = l + 1;
        while (l < r) { u64 mid = (l + r) >> 1; if (P.keys[mid] < left) l = mid + 1; else r = mid; steps++; }
        u64 e0 = l; u32 ne = 0, na = 0; const int want = dir;
        while (e0 + ne < hi && P.keys[e0 + ne] <= right) { u64 v0 = P.val_off[e0 + ne], v1 = P.val_off[e0 + ne + 1]; if (v1 > v0 && (int)(P.vals[v0] & 1) == want) na += (u32)(v1 - v0); ne++; }
        if (na) { have = true; h.q = q; h.mask_dir = (u32)(i << 1 | dir); h.e0 = e0; h.ne = ne; h.lo = c.lo; h.n = c.n; h.kmer = kmer; h.nanch = na * c.n; h.pad = 0; }
        if (stats) { atomicAdd((unsigned long long*)&stats[1], 1ull); atomicAdd((unsigned long long*)&stats[2], (unsigned long long)steps); atomicAdd((unsigned long long*)&stats[3], (unsigned long long)ne); }
      }
      if (stats) atomicAdd((unsigned long long*)&stats[0], 1ull);
    }
  }
  // warp-aggregated append
  u32 bal = __ballot_sync(FULLMASK, have);
  if (bal) { int lane = threadIdx.x & 31; u32 base = 0; if (lane == __ffs(bal) - 1) base = atomicAdd(nhits, __popc(bal)); base = __shfl_sync(FULLMASK, base, __ffs(bal) - 1);
    if (have) hits[base + __popc(bal & ((1u << lane) - 1))] = h; }
}

__global__ void k_hit_counts(const ProbeHit* __restrict__ h, u32 n, u64* __restrict__ c) { u32 t = blockIdx.x * blockDim.x + threadIdx.x; if (t <= n) c[t] = (t < n) ? h[t].nanch : 0; }

// anchor keys: hi = query<<36 | genome(dense)<<2 | (sorted only on bits >= 2) ; lo = QBegin<<36 | (63-Len)<<30 | TBegin<<2 | qrc<<1 | trc
__device__ __forceinline__ u64 pack_lo(i32 qb, u32 len, i32 tb, u32 qrc, u32 trc) { return ((u64)(u32)qb << 36) | ((u64)(63 - len) << 30) | ((u64)((u32)tb & 0x0FFFFFFFu) << 2) | (qrc << 1) | trc; }
__device__ __forceinline__ u64 pack_hi(u32 q, u32 g) { return ((u64)q << 36) | ((u64)g << 2); }

// one thread per hit: (matched key) x (query locations) x (values) -> anchors (lib-index-search.go:1398-1557)
__global__ void k_probe_emit(ProbeParams P, const ProbeHit* __restrict__ hits, const u64* __restrict__ hoff, u32 nh, const u32* __restrict__ qvals, const u64* __restrict__ koff,
                             const u32* __restrict__ batch_base, u64* __restrict__ a_hi, u64* __restrict__ a_lo) {
  u32 t = blockIdx.x * blockDim.x + threadIdx.x; if (t >= nh) return; ProbeHit h = hits[t]; u64 w = hoff[t]; const int K = P.k, want = h.mask_dir & 1; const u32* locs = qvals + koff[h.q] + h.lo;
  for (u32 e = 0; e < h.ne; e++) { u64 key = P.keys[h.e0 + e]; u64 v0 = P.val_off[h.e0 + e], v1 = P.val_off[h.e0 + e + 1]; if (v1 == v0 || (int)(P.vals[v0] & 1) != want) continue;
    int len = (__clzll(h.kmer ^ key) >> 1) + K - 32; if (h.kmer == key) len = K;   // Len = LZ(q^kmer)/2 + k - 32 (kv-searcher.go:480)
    for (u32 li = 0; li < h.n; li++) { u32 loc = locs[li] & 0x7fffffffu; u32 rcQ = loc & 1; i32 posQ = (i32)(loc >> 1);
      for (u64 vi = v0; vi < v1; vi++) { u64 rp = P.vals[vi]; u64 bgi = rp >> 30; u32 g = batch_base[bgi >> 17] + (u32)(bgi & 0x1ffff); i32 posT = (i32)((rp << 34) >> 36); u32 rv = rp & 1, rcT = (rp >> 1) & 1; i32 bq, bt;
        if (!rv) { bq = rcQ ? posQ + K - len : posQ; bt = rcT ? posT + K - len : posT; } else { bq = rcQ ? posQ : posQ + K - len; bt = rcT ? posT : posT + K - len; }
        a_hi[w] = pack_hi(h.q, g); a_lo[w] = pack_lo(bq, (u32)len, bt, rcQ, rcT); w++; } } }
}

// =====================================================================================================
// engine object
// =====================================================================================================
struct QBatch {  // device-side query batch
  int nq = 0; u64 total_bases = 0, total_k = 0;
  DBuf<u8> ascii, packed; DBuf<u64> off, boff, koff; std::vector<u64> h_off, h_boff, h_koff;
  DBuf<u64> qkeys; DBuf<u32> qvals;    // per-query sorted (k-mer, loc) tables
};

struct lmg_index {
  Image img; cudaStream_t st = 0; CubTemp tmp; int sm_count = 148; u32 smem_optin = 0; int use_tma = 1;
  double ms[8] = {0}; u64 counters[8] = {0}; std::mutex mu;
};

static thread_local std::string g_err;

static void upload_queries(lmg_index* ix, const u8* seqs, const u64* off, int nq, QBatch& B) {
  cudaStream_t st = ix->st; const int k = ix->img.k; B.nq = nq; B.h_off.assign(off, off + nq + 1); B.total_bases = off[nq] - off[0];
  if (off[0] != 0) for (auto& x : B.h_off) x -= off[0];
  B.h_boff.resize(nq + 1); B.h_koff.resize(nq + 1); u64 b = 0, kk = 0;
  for (int q = 0; q < nq; q++) { u64 L = B.h_off[q + 1] - B.h_off[q]; if (L >= (1ull << 27)) throw std::runtime_error("query longer than 2^27 bases is not supported");
    B.h_boff[q] = b; b += (((L + 3) >> 2) + 16 + 15) & ~15ull; B.h_koff[q] = kk; kk += (L >= (u64)k) ? 2 * (L - k + 1) : 0; }
  B.h_boff[nq] = b; B.h_koff[nq] = kk; B.total_k = kk;
  B.ascii.alloc(B.total_bases + 16, st); B.ascii.from_host(seqs + off[0], B.total_bases); B.off.alloc(nq + 1, st); B.off.from_host(B.h_off.data(), nq + 1);
  B.boff.alloc(nq + 1, st); B.boff.from_host(B.h_boff.data(), nq + 1); B.koff.alloc(nq + 1, st); B.koff.from_host(B.h_koff.data(), nq + 1);
  B.packed.alloc(b + 64, st); B.packed.zero();
}

// K1a: pack + k-mers + per-query stable sort by k-mer
static void sketch_tables(lmg_index* ix, QBatch& B) {
  cudaStream_t st = ix->st; const int k = ix->img.k; int nq = B.nq; u64 maxL = 0; for (int q = 0; q < nq; q++) maxL = std::max(maxL, B.h_off[q + 1] - B.h_off[q]);
  dim3 g1((unsigned)std::max(1, std::min(64, cdiv((i64)(maxL + 3) / 4, 256))), nq); k_pack_queries<<<g1, 256, 0, st>>>(B.ascii.p, B.off.p, B.boff.p, B.packed.p, nq); KERNEL_CHECK();
  DBuf<u64> keys_in(B.total_k + 2, st); DBuf<u32> vals_in(B.total_k + 2, st); B.qkeys.alloc(B.total_k + 2, st); B.qvals.alloc(B.total_k + 2, st);
  if (B.total_k == 0) return;
  dim3 g2((unsigned)std::max(1, std::min(64, cdiv((i64)maxL, 128))), nq); k_gen_kmers<<<g2, 128, 0, st>>>(B.packed.p, B.boff.p, B.off.p, B.koff.p, keys_in.p, vals_in.p, nq, k); KERNEL_CHECK();
  size_t tb = 0; cub::DeviceSegmentedSort::StableSortPairs(nullptr, tb, keys_in.p, B.qkeys.p, vals_in.p, B.qvals.p, (int)B.total_k, nq, B.koff.p, B.koff.p + 1, st);
  cub::DeviceSegmentedSort::StableSortPairs(ix->tmp.get(tb), tb, keys_in.p, B.qkeys.p, vals_in.p, B.qvals.p, (int)B.total_k, nq, B.koff.p, B.koff.p + 1, st); KERNEL_CHECK();
}

// K1b: capture
static void sketch_capture(lmg_index* ix, QBatch& B, DBuf<Capture>& cap, DBuf<u32>& owner) {
  cudaStream_t st = ix->st; const Image& I = ix->img; cap.alloc((u64)B.nq * I.m, st); owner.alloc(B.total_k + 2, st); owner.fill_ff();
  u64 maxn = 0; for (int q = 0; q < B.nq; q++) maxn = std::max(maxn, B.h_koff[q + 1] - B.h_koff[q]);
  u32 smem_cap = (u32)std::min<u64>((ix->smem_optin - 1024) / 8, 24576);  // entries
  u32 need = (u32)std::min<u64>(maxn, smem_cap); size_t smem = ((size_t)need * 8 + 15) & ~15ull;
  CUDA_CHECK(cudaFuncSetAttribute(k_capture, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 16)));
  // slices: enough CTAs to fill 148 SMs a few times over, but keep >= 1024 masks per CTA so the staged table is reused
  int slices = std::max(1, std::min(I.m / 1024, cdiv(ix->sm_count * 8, std::max(1, B.nq))));
  k_capture<<<B.nq * slices, 256, smem, st>>>(B.qkeys.p, B.koff.p, I.d_masks, I.m, I.k, slices, cap.p, owner.p, need, ix->use_tma); KERNEL_CHECK();
}

struct Anchors { u64 n = 0; DBuf<u64> hi, lo; };

static ProbeParams probe_params(const Image& I, int p) { ProbeParams P; P.bucket_off = I.d_bucket_off; P.keys = I.d_keys; P.val_off = I.d_val_off; P.vals = I.d_vals; P.anchor_start = I.d_anchor_start; P.m = I.m; P.k = I.k; P.NA = I.NA; P.mask_prefix = I.mask_prefix; P.anchor_prefix = I.anchor_prefix; P.p = p; return P; }

template <class K, class V> static void radix_sort_pairs(lmg_index* ix, DBuf<K>& k_in, DBuf<K>& k_out, DBuf<V>& v_in, DBuf<V>& v_out, u64 n, int begin_bit, int end_bit) {
  size_t tb = 0; cub::DeviceRadixSort::SortPairs(nullptr, tb, k_in.p, k_out.p, v_in.p, v_out.p, (i64)n, begin_bit, end_bit, ix->st);
  cub::DeviceRadixSort::SortPairs(ix->tmp.get(tb), tb, k_in.p, k_out.p, v_in.p, v_out.p, (i64)n, begin_bit, end_bit, ix->st); KERNEL_CHECK();
}

static int bits_for(u64 v) { int b = 1; while ((v >> b) && b < 64) b++; return b; }

// K2: probes -> anchors sorted by (query, genome, QBegin, QEnd desc, TBegin, qrc, trc)
static void seed_probe(lmg_index* ix, QBatch& B, const lmg_params* prm, DBuf<Capture>& cap, DBuf<u32>& owner, Anchors& A, bool stats) {
  cudaStream_t st = ix->st; const Image& I = ix->img; if (prm->min_prefix < I.mask_prefix + I.anchor_prefix || prm->min_prefix > I.k) throw std::runtime_error("the minimum prefix length should be in the range of [maskPrefix+anchorPrefix, k]");  // kv-searcher.go:202
  ProbeParams P = probe_params(I, prm->min_prefix); u64 nprobe = (u64)B.nq * I.m * 2;
  DBuf<u32> nh(1, st); nh.zero(); DBuf<u64> dstats(8, st); dstats.zero();
  // hit list capacity: every probe may hit
  u64 capHits = std::min<u64>(nprobe, 1ull << 31); DBuf<ProbeHit> hits;
  // first pass with a bounded list; typical hit rates are a few % of probes
  u64 tryCap = std::min<u64>(capHits, std::max<u64>(1u << 20, nprobe / 4));
  for (;;) { hits.alloc(tryCap, st); nh.zero(); if (stats) dstats.zero();
    k_probe_find<<<cdiv((i64)nprobe, 256), 256, 0, st>>>(P, cap.p, owner.p, B.koff.p, nprobe, hits.p, nh.p, stats ? dstats.p : nullptr); KERNEL_CHECK();
    u32 h = nh.to_host()[0]; if (h <= tryCap) { tryCap = h; break; } tryCap = capHits; }
  u32 nhit = (u32)tryCap; if (stats) { auto s = dstats.to_host(); for (int i = 0; i < 4; i++) ix->counters[i] = s[i]; ix->counters[4] = nhit; }
  A.n = 0; if (nhit == 0) return;
  DBuf<u64> hoff(nhit + 1, st);
  { DBuf<u64> cnt(nhit + 1, st); k_hit_counts<<<cdiv(nhit + 1, 256), 256, 0, st>>>(hits.p, nhit, cnt.p); KERNEL_CHECK();
    size_t tb = 0; cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt.p, hoff.p, (int)(nhit + 1), st); cub::DeviceScan::ExclusiveSum(ix->tmp.get(tb), tb, cnt.p, hoff.p, (int)(nhit + 1), st); KERNEL_CHECK(); }
  u64 total; CUDA_CHECK(cudaMemcpyAsync(&total, hoff.p + nhit, 8, cudaMemcpyDeviceToHost, st)); CUDA_CHECK(cudaStreamSynchronize(st));
  A.n = total; if (stats) ix->counters[5] = total; if (total == 0) return; if (total >= (1ull << 31)) throw std::runtime_error("more than 2^31 anchors in one batch; use smaller batches");
  DBuf<u64> hi0(total, st), lo0(total, st); A.hi.alloc(total, st); A.lo.alloc(total, st);
  k_probe_emit<<<cdiv(nhit, 128), 128, 0, st>>>(P, hits.p, hoff.p, nhit, B.qvals.p, B.koff.p, I.d_batch_base, hi0.p, lo0.p); KERNEL_CHECK();
  radix_sort_pairs(ix, lo0, A.lo, hi0, A.hi, total, 0, 64);             // by lo
  int gb = bits_for((u64)std::max(1, I.G)), qb = bits_for((u64)B.nq);
  radix_sort_pairs(ix, A.hi, hi0, A.lo, lo0, total, 2, 36 + qb); (void)gb;  // stable by (query, genome)
  std::swap(A.hi, hi0); std::swap(A.lo, lo0);
}

// =====================================================================================================
// K3: ClearSubstrPairs (lib-index-search.go:864-990) + Chainer.Chain (lib-chaining.go:122-633)
// =====================================================================================================
__device__ __forceinline__ i32 a_q(u64 lo) { return (i32)(lo >> 36); }
__device__ __forceinline__ i32 a_len(u64 lo) { return 63 - (i32)((lo >> 30) & 63); }
__device__ __forceinline__ i32 a_t(u64 lo) { return (i32)((lo >> 2) & 0x0FFFFFFF); }
__device__ __forceinline__ float seedw(float l) { return __fmul_rn(__fmul_rn(0.1f, l), l); }   // seedWeight lib-chaining.go:635 (no FMA contraction, as Go/amd64)

struct ChainParams { float max_gap, min_score, max_distance; int top_chains, k; const float* gap_score; int gap_tab; };

// one warp per (query, genome) segment. Anchors arrive sorted by (QBegin asc, QEnd desc, TBegin asc, qrc, trc).
// Phase 1: mark anchors nested in an earlier anchor within the k-window, compact in place order into c_lo.
// Phase 2: DP. For anchor i the predecessors are j < i with |TBegin diff| <= max_distance (RangeIndex query, :380-385),
//          QBegin/TBegin different, QBegin diff <= max_distance (break, :416), gap <= max_gap; lanes evaluate 32 candidates at a
//          time in descending j, the warp keeps the best score with ties to the larger j (= first strict improvement, :462).
__global__ void __launch_bounds__(128) k_clear_chain(const u64* __restrict__ lo_in, const u64* __restrict__ seg_off, u32 nseg, ChainParams P,
                                                     u64* __restrict__ c_lo, u32* __restrict__ c_n, float* __restrict__ score, u32* __restrict__ pred, signed char* __restrict__ dirs, u64* __restrict__ s2i) {
  u32 seg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; int lane = threadIdx.x & 31; if (seg >= nseg) return;
  u64 b = seg_off[seg]; u32 n = (u32)(seg_off[seg + 1] - b); const u64* A = lo_in + b; u64* C = c_lo + b; const int k = P.k;
  // ---- phase 1
  u32 kept = 0;
  for (u32 base = 0; base < n; base += 32) {
    u32 i = base + lane; bool keep = false;
    if (i < n) { keep = true; if (i > 0) { u64 v = A[i]; i32 vq = a_q(v), vl = a_len(v), vt = a_t(v); i32 vQEnd = vq + vl, up = max(vQEnd - k, 0), vTEnd = vt + vl;
        for (i32 j = (i32)i - 1; j >= 0; j--) { u64 p = A[j]; i32 pq = a_q(p); if (pq < up) break; i32 pl = a_len(p), pt = a_t(p); if (vQEnd <= pq + pl && vt >= pt && vTEnd <= pt + pl) { keep = false; break; } } } }
    u32 bal = __ballot_sync(FULLMASK, keep); if (keep) C[kept + __popc(bal & ((1u << lane) - 1))] = A[i]; kept += __popc(bal);
  }
  __syncwarp(); n = kept; if (lane == 0) c_n[seg] = n;
  float* S = score + b; u32* Pd = pred + b; signed char* D = dirs + b; u64* K2 = s2i + b;
  if (n == 1) { if (lane == 0) { float w = seedw((float)a_len(C[0])); S[0] = w; Pd[0] = 0; D[0] = 0; K2[0] = ((u64)__float_as_uint(w) << 32); } return; }
  if (lane == 0) { float w = seedw((float)a_len(C[0])); S[0] = w; Pd[0] = 0; D[0] = 0; K2[0] = ((u64)__float_as_uint(w) << 32); }
  __syncwarp();
  const i32 maxDist = (i32)P.max_distance;
  for (u32 i = 1; i < n; i++) {
    u64 av = C[i]; i32 aq = a_q(av), al = a_len(av), at = a_t(av); float m0 = seedw((float)al);
    float bs = -1.0f; i32 bj = -1; int bdir = 0; bool stop = false;
    for (i32 top = (i32)i - 1; top >= 0 && !stop; top -= 32) {
      i32 j = top - lane; float s = -1.0f; int dir = 0; bool valid = false;
      if (j >= 0) { u64 bv = C[j]; i32 bq = a_q(bv), bl = a_len(bv), bt = a_t(bv);
        if (aq - bq > maxDist) stop = true;   // all smaller j are at least as far
        else if (aq != bq && at != bt && bt >= (at < maxDist ? 0 : at - maxDist) && bt <= at + maxDist) {
          i32 dq = abs(aq - bq), dt; if (at >= bt) dt = abs(at - bt); else dt = abs(at + al - bt - bl);
          i32 g = abs(dq - dt);
          if ((float)g <= P.max_gap) {
            float w; if (aq > bq + bl) w = seedw((float)al); else if (g == 0) w = __fadd_rn(-seedw((float)bl), seedw((float)(aq + al - bq))); else w = seedw((float)(aq + al - (bq + bl)));
            dir = (at >= bt) ? 1 : -1; int dj = D[j]; float prev = (dj == 0 || dj == dir) ? S[j] : seedw((float)bl);
            float gs = (g == 0) ? 0.0f : P.gap_score[min(g, P.gap_tab - 1)];
            s = __fsub_rn(__fadd_rn(prev, w), gs); valid = (s >= P.min_score);
          } } }
      if (!valid) s = -1.0f;
      // warp arg-max, ties -> larger j (smaller lane)
      float rs = s; i32 rj = valid ? j : -1; int rd = dir;
      for (int o = 16; o; o >>= 1) { float os = __shfl_xor_sync(FULLMASK, rs, o); i32 oj = __shfl_xor_sync(FULLMASK, rj, o); int od = __shfl_xor_sync(FULLMASK, rd, o); if (os > rs || (os == rs && oj > rj)) { rs = os; rj = oj; rd = od; } }
      if (rj >= 0 && rs > bs) { bs = rs; bj = rj; bdir = rd; }
      stop = __any_sync(FULLMASK, stop);
    }
    if (lane == 0) { float m = m0; u32 mj = i; int md = 0; if (bj >= 0 && bs > m0) { m = bs; mj = (u32)bj; md = bdir; } S[i] = m; Pd[i] = mj; D[i] = (signed char)md; K2[i] = ((u64)__float_as_uint(m) << 32) | i; }
    __syncwarp();
  }
}

struct ChainRec { u32 seg, ord; i32 q0, t0, len0, q1, t1, len1; u32 flags1; i32 nseeds; float score; u32 pad; };  // flags1: bit1 qrc, bit0 trc of the LAST anchor

// one thread per segment: backtrack (lib-chaining.go:490-629). s2i_sorted ascending within the segment.
__global__ void k_backtrack(const u64* __restrict__ seg_off, const u32* __restrict__ c_n, u32 nseg, const u64* __restrict__ c_lo, const u32* __restrict__ pred, const signed char* __restrict__ dirs,
                            const u64* __restrict__ s2i_sorted, u8* __restrict__ visited, ChainParams P, ChainRec* __restrict__ out, u32* __restrict__ nout, u32 cap, float* __restrict__ seg_score) {
  u32 seg = blockIdx.x * blockDim.x + threadIdx.x; if (seg >= nseg) return; u64 b = seg_off[seg]; i32 n = (i32)c_n[seg]; const u64* C = c_lo + b; const u32* Pd = pred + b; const signed char* D = dirs + b; const u64* K2 = s2i_sorted + b; u8* V = visited + b;
  u32 ord = 0;
  auto emit = [&](i32 first, i32 last, i32 cnt, float sc) { u32 w = atomicAdd(nout, 1u); if (w < cap) { ChainRec r; r.seg = seg; r.ord = ord; u64 f = C[first], l = C[last]; r.q0 = a_q(f); r.t0 = a_t(f); r.len0 = a_len(f); r.q1 = a_q(l); r.t1 = a_t(l); r.len1 = a_len(l); r.flags1 = (u32)(l & 3); r.nseeds = cnt; r.score = sc; r.pad = 0; out[w] = r; } ord++; };
  if (n == 1) { float w = __uint_as_float((u32)(K2[0] >> 32)); seg_score[seg] = w; if (w >= P.min_score) emit(0, 0, 1, w); return; }
  i32 iMax = n - 1; float maxScore = 0; bool first = true; int nChecked = 0;
  for (;;) {
    nChecked++; if (P.top_chains > 0 && nChecked > P.top_chains) break;
    float M = 0; u32 Mi = 0;
    while (iMax >= 0) { u64 e = K2[iMax]; M = __uint_as_float((u32)(e >> 32)); Mi = (u32)e; if (!V[Mi]) { iMax--; break; } iMax--; }
    if (M < P.min_score) break;
    i32 i = (i32)Mi; if (first) { maxScore = M; first = false; }
    i32 cnt = 0, lastA = -1, firstA = -1;
    for (;;) { i32 j = (i32)Pd[i]; bool change = (i != j && D[j] != 0 && D[i] != D[j]);
      if (V[j] && !change) { cnt = 0; V[i] = 1; break; }
      if (cnt == 0) lastA = i; firstA = i; cnt++; V[i] = 1;
      if (i == j || change) { if (change) { firstA = j; cnt++; } emit(firstA, lastA, cnt, 0.0f); cnt = -1; break; } else i = j; }
  }
  seg_score[seg] = maxScore;
}

struct Segments { u32 nseg = 0; DBuf<u64> key, off; DBuf<u32> cn; DBuf<u64> c_lo; DBuf<float> score; std::vector<u64> h_key, h_off; };
struct Chains { u32 n = 0; DBuf<ChainRec> rec; std::vector<ChainRec> h; std::vector<float> seg_score; };

static void chain_stage(lmg_index* ix, const lmg_params* prm, Anchors& A, Segments& S, Chains& Cn) {
  cudaStream_t st = ix->st; u64 N = A.n; S.nseg = 0; Cn.n = 0; if (N == 0) return;
  // segments = runs of equal (query, genome)
  DBuf<u64> ukey(N, st); DBuf<u32> cnt(N + 1, st); DBuf<u32> nruns(1, st);
  { size_t tb = 0; cub::DeviceRunLengthEncode::Encode(nullptr, tb, A.hi.p, ukey.p, cnt.p, nruns.p, (int)N, st); cub::DeviceRunLengthEncode::Encode(ix->tmp.get(tb), tb, A.hi.p, ukey.p, cnt.p, nruns.p, (int)N, st); KERNEL_CHECK(); }
  u32 nseg = nruns.to_host()[0]; S.nseg = nseg; S.off.alloc(nseg + 1, st);
  { DBuf<u64> c64(nseg + 1, st); struct Dummy {}; // widen counts
    std::vector<u32> hc = cnt.to_host(nseg); S.h_off.assign(nseg + 1, 0); for (u32 i = 0; i < nseg; i++) S.h_off[i + 1] = S.h_off[i] + hc[i]; S.off.from_host(S.h_off.data(), nseg + 1); }
  S.h_key = ukey.to_host(nseg); S.key = std::move(ukey);
  // gap-score table on the host (gapScore lib-chaining.go:662: 0.1*g + 0.5*float32(log2(float64(g))), g integer <= max_gap)
  int gt = std::max(2, (int)std::floor(prm->max_gap) + 2); std::vector<float> gtab(gt, 0.0f);
  for (int g = 1; g < gt; g++) { volatile float a = 0.1f * (float)g; volatile float bb = 0.5f * (float)std::log2((double)g); volatile float c = a + bb; gtab[g] = c; }
  DBuf<float> dgt(gt, st); dgt.from_host(gtab.data(), gt);
  ChainParams P; P.max_gap = prm->max_gap; { volatile float w = 0.1f * (float)prm->min_single_prefix; volatile float w2 = w * (float)prm->min_single_prefix; P.min_score = w2; } P.max_distance = prm->max_distance; P.top_chains = prm->top_n_chains; P.k = ix->img.k; P.gap_score = dgt.p; P.gap_tab = gt;
  S.c_lo.alloc(N, st); S.cn.alloc(nseg, st); S.score.alloc(nseg, st); DBuf<float> sc(N, st); DBuf<u32> pred(N, st); DBuf<signed char> dirs(N, st); DBuf<u64> s2i(N, st), s2i_s(N, st); DBuf<u8> visited(N, st); visited.zero();
  k_clear_chain<<<cdiv((i64)nseg * 32, 128), 128, 0, st>>>(A.lo.p, S.off.p, nseg, P, S.c_lo.p, S.cn.p, sc.p, pred.p, dirs.p, s2i.p); KERNEL_CHECK();
  // per-segment ascending sort of (score bits << 32 | index) over the compacted prefix of each segment
  DBuf<u64> seg_end(nseg, st);
  { std::vector<u32> hcn = S.cn.to_host(nseg); std::vector<u64> he(nseg); for (u32 i = 0; i < nseg; i++) he[i] = S.h_off[i] + hcn[i]; seg_end.from_host(he.data(), nseg); CUDA_CHECK(cudaStreamSynchronize(st)); }
  { size_t tb = 0; cub::DeviceSegmentedSort::SortKeys(nullptr, tb, s2i.p, s2i_s.p, (int)N, (int)nseg, S.off.p, seg_end.p, st); cub::DeviceSegmentedSort::SortKeys(ix->tmp.get(tb), tb, s2i.p, s2i_s.p, (int)N, (int)nseg, S.off.p, seg_end.p, st); KERNEL_CHECK(); }
  u32 cap = (u32)std::min<u64>(N + nseg, 0x7fffffffu); Cn.rec.alloc(cap, st); DBuf<u32> nout(1, st); nout.zero();
  k_backtrack<<<cdiv(nseg, 128), 128, 0, st>>>(S.off.p, S.cn.p, nseg, S.c_lo.p, pred.p, dirs.p, s2i_s.p, visited.p, P, Cn.rec.p, nout.p, cap, S.score.p); KERNEL_CHECK();
  u32 nc = nout.to_host()[0]; if (nc > cap) throw std::runtime_error("chain list overflow"); Cn.n = nc; Cn.seg_score = S.score.to_host(nseg);
  Cn.h = Cn.rec.to_host(nc);
  // drop genomes below min score (:1724), optional top-N genomes per query (:1780-1805), order chains by (segment, first TBegin, emission order) (:1967-1974)
  std::vector<char> keep(nseg, 1); for (u32 s = 0; s < nseg; s++) if (Cn.seg_score[s] < P.min_score) keep[s] = 0;
  if (prm->top_n_genomes > 0) { u32 s0 = 0; while (s0 < nseg) { u32 q = (u32)(S.h_key[s0] >> 36), s1 = s0; std::vector<u32> v; while (s1 < nseg && (u32)(S.h_key[s1] >> 36) == q) { if (keep[s1]) v.push_back(s1); s1++; }
      if ((int)v.size() > prm->top_n_genomes) { std::stable_sort(v.begin(), v.end(), [&](u32 a, u32 b) { return Cn.seg_score[a] > Cn.seg_score[b]; }); for (size_t t = prm->top_n_genomes; t < v.size(); t++) keep[v[t]] = 0; } s0 = s1; } }
  std::vector<ChainRec> kept; kept.reserve(nc); for (auto& r : Cn.h) if (keep[r.seg]) { r.score = Cn.seg_score[r.seg]; kept.push_back(r); }
  std::sort(kept.begin(), kept.end(), [](const ChainRec& a, const ChainRec& b) { if (a.seg != b.seg) return a.seg < b.seg; if (a.t0 != b.t0) return a.t0 < b.t0; return a.ord < b.ord; });
  Cn.h.swap(kept); Cn.n = (u32)Cn.h.size();
}

// =====================================================================================================
// C ABI (part 1)
// =====================================================================================================
extern "C" {

void lmg_default_params(lmg_params* p) { p->min_prefix = 15; p->min_single_prefix = 17; p->top_n_genomes = 0; p->top_n_chains = 0; p->max_gap = 50; p->max_distance = 1000; p->ext_len = 1000; p->ext_len2 = 50;
  p->min_qcov_genome = 0; p->max_evalue = 10; p->align_max_gap = 20; p->align_min_len = 50; p->align_band = 100; p->output_seq = 0; p->min_pident = 70; p->min_qcov_hsp = 0; }
const char* lmg_last_error(void) { return g_err.c_str(); }

int lmg_index_open(const char* dir, int device, int shard, int n_shards, lmg_index** out) {
  try { int ndev = 0; CUDA_CHECK(cudaGetDeviceCount(&ndev)); if (ndev == 0) throw std::runtime_error("no CUDA device: the LexicMap GPU path has no CPU fallback");
    lmg_index* ix = new lmg_index; ix->img.load(dir, device, shard, std::max(1, n_shards)); CUDA_CHECK(cudaStreamCreateWithFlags(&ix->st, cudaStreamNonBlocking)); ix->tmp.st = ix->st;
    cudaDeviceProp pr; CUDA_CHECK(cudaGetDeviceProperties(&pr, device)); ix->sm_count = pr.multiProcessorCount; ix->smem_optin = (u32)pr.sharedMemPerBlockOptin; if (pr.major < 9) ix->use_tma = 0;
    if (getenv("LMG_NO_TMA")) ix->use_tma = 0;
    cudaMemPool_t pool; CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, device)); u64 thr = ~0ull; CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    *out = ix; return 0; } catch (std::exception& e) { g_err = e.what(); return -1; }
}
int lmg_index_info(const lmg_index* ix, lmg_info* o) { const Image& I = ix->img; o->k = I.k; o->masks = I.m; o->chunks = I.info.chunks; o->partitions = I.info.partitions; o->genomes = I.G; o->genome_batches = I.info.genome_batches;
  o->contig_interval = I.contig_interval; o->mask_prefix = I.mask_prefix; o->anchor_prefix = I.anchor_prefix; o->input_bases = I.total_bases; o->seed_keys = I.E; o->seed_values = I.V; o->image_bytes = I.bytes; return 0; }
int lmg_genome_name(const lmg_index* ix, uint64_t genome, const char** name) { auto it = ix->img.bgi2dense.find(genome); if (it == ix->img.bgi2dense.end()) { *name = ""; return -1; } *name = ix->img.genome_names[it->second].c_str(); return 0; }
void lmg_index_close(lmg_index* ix) { if (!ix) return; cudaSetDevice(ix->img.device); cudaStreamSynchronize(ix->st); ix->img.release(); if (ix->tmp.p) { cudaFree(ix->tmp.p); ix->tmp.p = nullptr; } cudaStreamDestroy(ix->st); delete ix; }
void lmg_free(void* p) { free(p); }
int lmg_last_timing(const lmg_index* ix, double* ms8, uint64_t* c8) { for (int i = 0; i < 8; i++) { if (ms8) ms8[i] = ix->ms[i]; if (c8) c8[i] = ix->counters[i]; } return 0; }

int lmg_mask_batch(lmg_index* ix, const uint8_t* seqs, const uint64_t* off, int32_t n, uint64_t* kmers, uint32_t* nlocs, uint32_t* minloc, uint64_t* suf, uint64_t suf_cap, uint64_t* n_suf) {
  try { std::lock_guard<std::mutex> lk(ix->mu); CUDA_CHECK(cudaSetDevice(ix->img.device)); QBatch B; upload_queries(ix, seqs, off, n, B); sketch_tables(ix, B); DBuf<Capture> cap; DBuf<u32> owner; sketch_capture(ix, B, cap, owner);
    const int m = ix->img.m, k = ix->img.k; auto hc = cap.to_host(); auto hv = B.qvals.to_host(); auto ho = owner.to_host(); u64 ns = 0; std::vector<std::array<u64, 4>> trip;
    for (int q = 0; q < n; q++) { trip.clear();
      for (int i = 0; i < m; i++) { const Capture& c = hc[(u64)q * m + i]; u64 o = (u64)q * m + i; kmers[o] = c.kmer; nlocs[o] = c.kmer ? c.n : 0; u32 mn = 0xffffffffu; if (c.kmer) for (u32 t = 0; t < c.n; t++) mn = std::min(mn, hv[B.h_koff[q] + c.lo + t] & 0x7fffffffu); minloc[o] = c.kmer ? mn : 0;
        if (c.kmer && ho[B.h_koff[q] + c.lo] == (u32)i) trip.push_back({(u64)q, (u64)c.smask, (u64)i, kmer_reverse62(c.kmer, k)}); }
      std::sort(trip.begin(), trip.end()); for (auto& t : trip) { if (ns < suf_cap) memcpy(suf + 4 * ns, t.data(), 32); ns++; } }
    *n_suf = ns; return 0; } catch (std::exception& e) { g_err = e.what(); return -1; }
}

int lmg_anchor_batch(lmg_index* ix, const lmg_params* p, const uint8_t* seqs, const uint64_t* off, int32_t n, lmg_anchor** out, uint64_t* n_out) {
  try { std::lock_guard<std::mutex> lk(ix->mu); CUDA_CHECK(cudaSetDevice(ix->img.device)); QBatch B; upload_queries(ix, seqs, off, n, B); sketch_tables(ix, B); DBuf<Capture> cap; DBuf<u32> owner; sketch_capture(ix, B, cap, owner);
    Anchors A; seed_probe(ix, B, p, cap, owner, A, true); auto hi = A.hi.to_host(A.n), lo = A.lo.to_host(A.n); lmg_anchor* o = (lmg_anchor*)malloc(sizeof(lmg_anchor) * (A.n + 1));
    for (u64 i = 0; i < A.n; i++) { lmg_anchor& a = o[i]; u32 g = (u32)((hi[i] >> 2) & 0x3FFFFFFFFull); a.genome = ix->img.genome_bgi[g]; a.query = (u32)(hi[i] >> 36); a.qbegin = (i32)(lo[i] >> 36); a.len = (u8)(63 - ((lo[i] >> 30) & 63)); a.tbegin = (i32)((lo[i] >> 2) & 0x0FFFFFFF); a.qrc = (lo[i] >> 1) & 1; a.trc = lo[i] & 1; a.pad = 0; }
    *out = o; *n_out = A.n; return 0; } catch (std::exception& e) { g_err = e.what(); return -1; }
}


#define LMG_HAVE_CHAIN 1
int lmg_chain_batch(lmg_index* ix, const lmg_params* p, const uint8_t* seqs, const uint64_t* off, int32_t n, lmg_chain** out, uint64_t* n_out) {
  try { std::lock_guard<std::mutex> lk(ix->mu); CUDA_CHECK(cudaSetDevice(ix->img.device)); QBatch B; upload_queries(ix, seqs, off, n, B); sketch_tables(ix, B); DBuf<Capture> cap; DBuf<u32> owner; sketch_capture(ix, B, cap, owner);
    Anchors A; seed_probe(ix, B, p, cap, owner, A, false); Segments S; Chains C; chain_stage(ix, p, A, S, C);
    lmg_chain* o = (lmg_chain*)malloc(sizeof(lmg_chain) * (C.n + 1));
    for (u32 i = 0; i < C.n; i++) { const ChainRec& r = C.h[i]; lmg_chain& c = o[i]; u64 key = S.h_key[r.seg]; c.query = (u32)(key >> 36); c.genome = ix->img.genome_bgi[(u32)((key >> 2) & 0x3FFFFFFFFull)]; c.score = r.score; c.n_seeds = r.nseeds;
      c.q0 = r.q0; c.t0 = r.t0; c.len0 = r.len0; c.q1 = r.q1; c.t1 = r.t1; c.len1 = r.len1; bool qrc = (r.flags1 >> 1) & 1, trc = r.flags1 & 1; c.rc = (r.nseeds == 1) ? (qrc != trc) : (r.t0 > r.t1); }
    *out = o; *n_out = C.n; return 0; } catch (std::exception& e) { g_err = e.what(); return -1; }
}
}  // extern "C"

// ---- not yet implemented entry points (fail loudly)
extern "C" {
#ifndef LMG_HAVE_CHAIN
int lmg_chain_batch(lmg_index*, const lmg_params*, const uint8_t*, const uint64_t*, int32_t, lmg_chain**, uint64_t*) { g_err = "lmg_chain_batch: not implemented"; return -2; }
#endif
#ifndef LMG_HAVE_SEARCH
int lmg_search_batch(lmg_index*, const lmg_params*, const uint8_t*, const uint64_t*, int32_t, lmg_results**) { g_err = "lmg_search_batch: not implemented"; return -2; }
int lmg_results_rows(const lmg_results*, const lmg_hsp**, uint64_t*, const char**, uint64_t*) { return -2; }
int lmg_results_seq_id(const lmg_results*, uint64_t, const char**) { return -2; }
void lmg_results_free(lmg_results*) {}
#endif
#ifndef LMG_HAVE_WFA
int lmg_wfa_batch(int, const uint8_t*, const uint64_t*, int32_t, char**, uint64_t*) { g_err = "lmg_wfa_batch: not implemented"; return -2; }
#endif
}
