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
#include <atomic>
#include <thread>
#include <condition_variable>
#include <functional>
#include <exception>
#include <set>
#include <array>
#include <memory>
#include <sched.h>

// =====================================================================================================
// small device utilities
// =====================================================================================================
#define FULLMASK 0xffffffffu
static std::atomic<unsigned long long> g_launches{0};   // kernels of THIS library launched (CUB's internal kernels are not counted)
#undef KERNEL_CHECK
#define KERNEL_CHECK() do { g_launches++; CUDA_CHECK(cudaGetLastError()); } while (0)
#define CUB_CHECK() CUDA_CHECK(cudaGetLastError())
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
__device__ __forceinline__ u32 is_acgt(u8 c) { c &= 0xDF; return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }
// amask: 1 bit per base (bit i&7 of byte i>>3) = base is not A/C/G/T; base-level alignment compares raw bytes in the reference, so such bases never match
__global__ void k_pack_queries(const u8* __restrict__ ascii, const u64* __restrict__ off, const u64* __restrict__ boff, u8* __restrict__ packed, u8* __restrict__ amask, int nq) {
  int q = blockIdx.y; if (q >= nq) return; u64 a0 = off[q], L = off[q + 1] - a0; u64 nb = (L + 3) >> 2; u8* out = packed + boff[q]; u8* am = amask + boff[q];
  for (u64 b = blockIdx.x * (u64)blockDim.x + threadIdx.x; b < ((L + 7) >> 3); b += (u64)gridDim.x * blockDim.x) { u32 v = 0; for (int j = 0; j < 8; j++) { u64 i = b * 8 + j; if (i < L && !is_acgt(ascii[a0 + i])) v |= 1u << j; } am[b] = (u8)v; }
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
struct CapSoA { u64* kmer; u32 *lo, *n, *smask; };   // structure-of-arrays storage: the probe kernel reads 8 B (prefix probe) or 16 B (suffix probe) per slot
__global__ void __launch_bounds__(256) k_capture(const u64* __restrict__ qkeys, const u64* __restrict__ koff, const u64* __restrict__ masks, int m, int k, int slices,
                                                 CapSoA cap, u32* __restrict__ owner, u32 smem_cap_entries, int use_tma, const u32* __restrict__ mask_pstart, int mask_pbits) {
  extern __shared__ __align__(128) u8 smem_raw[]; u64* stab = (u64*)smem_raw; __shared__ __align__(8) u64 mbar; __shared__ u32 pst[1025], pen[1024];
  int q = blockIdx.x / slices, sl = blockIdx.x % slices; u64 o = koff[q]; u32 n = (u32)(koff[q + 1] - o);
  int per = (m + slices - 1) / slices, i0 = sl * per, i1 = min(m, i0 + per);
  if (n == 0) { for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) { u64 w = (u64)q * m + i; cap.kmer[w] = 0; cap.lo[w] = 0; cap.n[w] = 0; cap.smask[w] = 0; } return; }
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
  // prefix directory over the sorted table: the descent for a mask starts inside the bucket of k-mers sharing its leading pb bits
  // (the XOR-argmin always shares the longest available prefix), which skips the widest binary searches
  int pb = 31 - __clz(max(n, 16u)) - 3; pb = max(4, min(pb, 10)); const int psh = 2 * k - pb; const u32 NP = 1u << pb;
  for (u32 t = threadIdx.x; t < NP; t += blockDim.x) { pst[t] = 0xFFFFFFFFu; pen[t] = 0; } __syncthreads();
  for (u32 t = threadIdx.x; t < n; t += blockDim.x) { u32 p = (u32)(tab[t] >> psh); if (t == 0 || (u32)(tab[t - 1] >> psh) != p) pst[p] = t; if (t + 1 == n || (u32)(tab[t + 1] >> psh) != p) pen[p] = t + 1; } __syncthreads();
  const int msh = 2 * k - mask_pbits;
  for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
    u64 mk = masks[i]; u32 p = (u32)(mk >> psh); u32 lo = pst[p], hi = pen[p]; if (lo == 0xFFFFFFFFu) { lo = 0; hi = n; } xor_argmin_range(tab, lo, hi, mk);
    u64 km = tab[lo]; Capture c; c.kmer = km; c.lo = lo; c.n = hi - lo; c.smask = 0;
    if (kmer_low_complexity(km, k)) c.kmer = 0;   // km==0 is DUST-low-complexity too, as in the reference
    else { u64 rv = kmer_reverse62(km, k); u32 mp = (u32)(rv >> msh); u32 a = mask_pstart[mp], b = mask_pstart[mp + 1]; if (a == b) { a = 0; b = (u32)m; } xor_argmin_range(masks, a, b, rv); c.smask = a; atomicMin(&owner[o + lo], (u32)i); }
    u64 w = (u64)q * m + i; cap.kmer[w] = c.kmer; cap.lo[w] = c.lo; cap.n[w] = c.n; cap.smask[w] = c.smask;
  }
}

// =====================================================================================================
// K2: seed probe (kv.Searcher.Search / Search2, kv/kv-searcher.go:190-1088) + anchors (lib-index-search.go:1357-1569)
// =====================================================================================================
struct ProbeHit { u32 q, mask_dir; u64 e0; u32 ne; u32 lo, n; u64 kmer; u32 nanch; u32 bucket; };  // mask_dir = capturing mask<<1 | dir ; [lo,lo+n) = query table rows (locs); bucket = mask bucket searched; [e0, e0+ne) = matched index entries
struct ProbeParams { const u64 *bucket_off, *bucket_voff, *vals; const SeedEntry* entries; const u32 *anchor_cbase, *anchor_cstart, *anchor_bits, *pbloom; const u16 *anchor_cum, *anchor_cstart16; u32* bcnt; u32 pbmask; int m, k, NA, mask_prefix, anchor_prefix, p, hints; };   // bcnt: per-call histogram of the surviving probes over the buckets (filled by the kernels that emit them; may be null)
// L2 cache policies for the lookup kernel: the compact anchor table (tens of MB, re-read by every probe) must stay in L2 while the random sectors of the
// 16-byte entries (hundreds of MB to GB, each touched about once per launch) stream through it. Without hints the entry sectors evict the table and every
// probe pays DRAM bursts for its bitmap word, rank and start on top of the entry itself (ncu, round 2: L2 hit rate 30 %, 383 B of DRAM traffic per probe).
__device__ __forceinline__ u64 l2_policy(int kind) { u64 p; if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); else if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ ulonglong2 ld_hint_v2(const void* a, u64 pol) { ulonglong2 r; asm volatile("ld.global.nc.L2::cache_hint.v2.u64 {%0, %1}, [%2], %3;" : "=l"(r.x), "=l"(r.y) : "l"(a), "l"(pol)); return r; }
__device__ __forceinline__ u64 ld_hint_u64(const void* a, u64 pol) { u64 r; asm volatile("ld.global.nc.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(r) : "l"(a), "l"(pol)); return r; }
__device__ __forceinline__ u32 ld_hint_u32(const void* a, u64 pol) { u32 r; asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(a), "l"(pol)); return r; }
__device__ __forceinline__ u32 ld_hint_u16(const void* a, u64 pol) { unsigned short r; asm volatile("ld.global.nc.L2::cache_hint.u16 %0, [%1], %2;" : "=h"(r) : "l"(a), "l"(pol)); return r; }
// bucket-relative index of the first key of a present anchor, from its rank among the bucket's present anchors (the rank is computed where the bitmap word is already loaded: probe_may_hit)
__device__ __forceinline__ u32 anchor_start_of(const ProbeParams& P, u32 bucket, u32 rank, u64 pol) { const u64 i = (u64)ld_hint_u32(P.anchor_cbase + bucket, pol) + rank; return P.anchor_cstart16 ? ld_hint_u16(P.anchor_cstart16 + i, pol) : ld_hint_u32(P.anchor_cstart + i, pol); }
// a probe survives when its anchor exists in the bucket's anchor table (the reference's own test) AND some key of the bucket starts with the probe's first maskPrefix+anchorPrefix bases (prefix Bloom filter, image.cuh)
__device__ __forceinline__ bool probe_may_hit(const ProbeParams& P, u32 bucket, u64 left, int ash, u32& arank) { const u32 pre = (u32)(left >> ash); const u64 aslot = (u64)bucket * P.NA + (pre & (u32)(P.NA - 1)); const u32 bits = P.anchor_bits[aslot >> 5]; if (!((bits >> (aslot & 31)) & 1)) return false;
  const u32 hb = pb_hash(bucket, pre) & P.pbmask; if (!((P.pbloom[hb >> 5] >> (hb & 31)) & 1)) return false;
  arank = (u32)P.anchor_cum[aslot >> 5] + __popc(bits & ((1u << (aslot & 31)) - 1)); return true; }   /* rank of the anchor among the bucket's present anchors */

// Probe semantics (kv.Searcher.Search / Search2): dir 0 = captured k-mer against its own mask bucket, values must have reverse flag 0;
// dir 1 = base-reversed k-mer against bucket `smask`, reverse flag 1 (decided by the FIRST value of each key: on-disk searcher
// semantics, kv-searcher.go:466-488). Range = keys in [kmer & ~low, kmer | low] at/after the anchor start (:282-304, :349-355).
// ---- two-phase probe. Phase A (k_probe_filter, one thread per (query, mask)): everything that can be decided from coalesced / L2-resident
// data — captured k-mer present, first-owner test of the reversed k-mer, anchor presence bit — and compaction of the surviving probes.
// Phase B (k_probe_find2, one thread per survivor): the dependent random HBM accesses (anchor start, key search, value flags) with all
// lanes of a warp on the long path instead of ~1 in 3.
struct __align__(16) Surv { u64 kmer; u32 qi; u32 bucket_dir; u32 lo, n; u32 arank, pad; };   // one 32-byte sector. bucket_dir = bucket | dir << 31; arank = rank of the probe's anchor among the bucket's present anchors; [lo, lo+n) = rows of the query table holding the captured k-mer
__global__ void __launch_bounds__(256) k_probe_filter(ProbeParams P, CapSoA cap, const u32* __restrict__ owner, const u64* __restrict__ koff, u64 nslot, Surv* __restrict__ surv, u32* __restrict__ nsurv, u32 cap_surv, u64* __restrict__ stats) {
  u64 qi = blockIdx.x * (u64)blockDim.x + threadIdx.x; bool s0 = false, s1 = false; Surv a, b; u32 issued = 0;
  if (qi < nslot) { u64 kmer = cap.kmer[qi];
    if (kmer != 0) { u32 q = (u32)(qi / P.m); int i = (int)(qi % P.m); const int s2 = (P.k - P.p) << 1; const u64 low = (P.p < P.k) ? ((1ull << s2) - 1) : 0; const int ash = (P.k - P.mask_prefix - P.anchor_prefix) << 1;
      { u64 left = kmer & ~low; issued++; if (probe_may_hit(P, (u32)i, left, ash, a.arank)) { s0 = true; a.kmer = kmer; a.qi = (u32)qi; a.bucket_dir = (u32)i; a.pad = 0; a.lo = cap.lo[qi]; a.n = cap.n[qi]; } }
      u32 lo = cap.lo[qi]; if (owner[koff[q] + lo] == (u32)i) { u64 rv = kmer_reverse62(kmer, P.k); u32 sm = cap.smask[qi]; u64 left = rv & ~low; issued++;
        if (probe_may_hit(P, sm, left, ash, b.arank)) { s1 = true; b.kmer = rv; b.qi = (u32)qi; b.bucket_dir = sm | 0x80000000u; b.pad = 0; b.lo = lo; b.n = cap.n[qi]; } } } }
  int lane = threadIdx.x & 31; u32 b0 = __ballot_sync(FULLMASK, s0), b1 = __ballot_sync(FULLMASK, s1); u32 tot = __popc(b0) + __popc(b1);
  if (tot) { u32 base = 0; if (lane == 0) base = atomicAdd(nsurv, tot); base = __shfl_sync(FULLMASK, base, 0);
    if (s0) { u32 w = base + __popc(b0 & ((1u << lane) - 1)); if (w < cap_surv) surv[w] = a; if (P.bcnt) atomicAdd(&P.bcnt[a.bucket_dir & 0x7FFFFFFFu], 1u); } if (s1) { u32 w = base + __popc(b0) + __popc(b1 & ((1u << lane) - 1)); if (w < cap_surv) surv[w] = b; if (P.bcnt) atomicAdd(&P.bcnt[b.bucket_dir & 0x7FFFFFFFu], 1u); } }
  if (stats) { for (int o = 16; o; o >>= 1) issued += __shfl_xor_sync(FULLMASK, issued, o); if (lane == 0 && issued) atomicAdd((unsigned long long*)&stats[0], (unsigned long long)issued); }
}
// K1b + K2 phase A fused (queries whose sorted k-mer table fits in shared memory): one CTA per query captures every mask (pass 1, as
// k_capture), tests the anchor-presence bit of the prefix probe at once, and records the first mask that captured each table row in shared
// memory; pass 2 walks the table ROWS: only a row's first owner searches for the suffix mask of the base-reversed k-mer (about one mask
// in ten), tests its anchor bit and emits the suffix probe. Nothing per (query, mask) slot goes to HBM — only the surviving probes do.
// DUMP additionally writes the per-slot capture arrays for lmg_mask_batch.
template <bool DUMP>
__global__ void __launch_bounds__(1024) k_capture2(const u64* __restrict__ qkeys, const u64* __restrict__ koff, const u64* __restrict__ masks, ProbeParams P, CapSoA cap, u32* __restrict__ owner_g, u32 max_n, int use_tma,
                                                  const u32* __restrict__ mask_pstart, int mask_pbits, Surv* __restrict__ surv, u32* __restrict__ nsurv, u32 cap_surv, u64* __restrict__ stats) {
  extern __shared__ __align__(128) u8 smem_raw[]; u64* stab = (u64*)smem_raw; u32* own = (u32*)(stab + max_n); __shared__ __align__(8) u64 mbar; __shared__ u32 pst[1025], pen[1024]; __shared__ u32 lcb[512];   // lcb: low-complexity flag per table row (rows <= 16384)
  const int q = blockIdx.x, m = P.m, k = P.k, lane = threadIdx.x & 31; const u64 o = koff[q]; const u32 n = (u32)(koff[q + 1] - o);
  if (n == 0) { if (DUMP) for (int i = threadIdx.x; i < m; i += blockDim.x) { u64 w = (u64)q * m + i; cap.kmer[w] = 0; cap.lo[w] = 0; cap.n[w] = 0; cap.smask[w] = 0; } return; }
  if (use_tma) { if (threadIdx.x == 0) mbar_init(&mbar, 1); __syncthreads(); if (threadIdx.x == 0) { u32 bytes = ((n * 8u) + 15u) & ~15u; tma_load_1d(stab, qkeys + o, bytes, &mbar); } for (u32 t = threadIdx.x; t < n; t += blockDim.x) own[t] = 0xFFFFFFFFu; mbar_wait(&mbar, 0); }
  else { for (u32 t = threadIdx.x; t < n; t += blockDim.x) { stab[t] = qkeys[o + t]; own[t] = 0xFFFFFFFFu; } }
  const u64* tab = stab;
  int pb = 31 - __clz(max(n, 16u)) - 3; pb = max(4, min(pb, 10)); const int psh = 2 * k - pb; const u32 NP = 1u << pb;
  for (u32 t = threadIdx.x; t < NP; t += blockDim.x) { pst[t] = 0xFFFFFFFFu; pen[t] = 0; } __syncthreads();
  for (u32 t = threadIdx.x; t < n; t += blockDim.x) { u32 p = (u32)(tab[t] >> psh); if (t == 0 || (u32)(tab[t - 1] >> psh) != p) pst[p] = t; if (t + 1 == n || (u32)(tab[t + 1] >> psh) != p) pen[p] = t + 1; }
  // the DUST score is a property of the k-mer, not of the mask: once per table row (about 2,000) instead of once per mask (20,000); it was 65 % of this kernel's instructions
  for (u32 tb = 0; tb < n; tb += blockDim.x) { u32 t = tb + threadIdx.x; bool lc = (t < n) && kmer_low_complexity(tab[t], k); u32 bal = __ballot_sync(FULLMASK, lc); if (lane == 0 && (tb + threadIdx.x) < ((n + 31) & ~31u)) lcb[(tb + threadIdx.x) >> 5] = bal; }
  __syncthreads();
  const int s2 = (k - P.p) << 1; const u64 low = (P.p < k) ? ((1ull << s2) - 1) : 0; const int ash = (k - P.mask_prefix - P.anchor_prefix) << 1; const int msh = 2 * k - mask_pbits; u32 issued = 0;
  // survivors are appended with ONE global atomic per CTA and loop iteration (ballot -> per-warp counts in shared memory -> thread 0), double-buffered by iteration parity:
  // one atomic per warp and iteration meant ~5 million same-address atomics per batch
  __shared__ u32 e_cnt[2][32], e_base[2]; int ep = 0; const int wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  auto emit = [&](bool have, const Surv& r) { const u32 bal = __ballot_sync(FULLMASK, have); if (lane == 0) e_cnt[ep][wid] = __popc(bal); __syncthreads();
    if (threadIdx.x == 0) { u32 tot = 0; for (int i = 0; i < nwarp; i++) { const u32 c = e_cnt[ep][i]; e_cnt[ep][i] = tot; tot += c; } e_base[ep] = tot ? atomicAdd(nsurv, tot) : 0u; } __syncthreads();
    if (have) { const u32 w = e_base[ep] + e_cnt[ep][wid] + __popc(bal & ((1u << lane) - 1)); if (w < cap_surv) surv[w] = r; if (P.bcnt) atomicAdd(&P.bcnt[r.bucket_dir & 0x7FFFFFFFu], 1u); } ep ^= 1; };
  // pass 1: masks
  for (int ib = 0; ib < m; ib += blockDim.x) { int i = ib + threadIdx.x; bool s0 = false; Surv r;
    if (i < m) { u64 mk = masks[i]; u32 p = (u32)(mk >> psh); u32 lo = pst[p], hi = pen[p]; if (lo == 0xFFFFFFFFu) { lo = 0; hi = n; } xor_argmin_range(tab, lo, hi, mk);
      u64 km = tab[lo]; bool lc = (lcb[lo >> 5] >> (lo & 31)) & 1;   // km==0 is DUST-low-complexity too, as in the reference
      if (!lc) { atomicMin(&own[lo], (u32)i); u64 left = km & ~low; issued++;
        if (probe_may_hit(P, (u32)i, left, ash, r.arank)) { s0 = true; r.kmer = km; r.qi = (u32)((u64)q * m + i); r.bucket_dir = (u32)i; r.pad = 0; r.lo = lo; r.n = hi - lo; } }
      if (DUMP) { u64 w = (u64)q * m + i; cap.kmer[w] = lc ? 0 : km; cap.lo[w] = lo; cap.n[w] = hi - lo; cap.smask[w] = 0; } }
    emit(s0, r); }
  __syncthreads();
  // pass 2: table rows; own[r] is set only at the first row of a run of equal k-mers that some mask captured
  for (u32 rb = 0; rb < n; rb += blockDim.x) { u32 rr = rb + threadIdx.x; bool s1 = false; Surv r;
    if (rr < n) { u32 i = own[rr];
      if (i != 0xFFFFFFFFu) { u64 km = tab[rr]; u32 c = 1; while (rr + c < n && tab[rr + c] == km) c++;
        u64 rv = kmer_reverse62(km, k); u32 mp = (u32)(rv >> msh); u32 a = mask_pstart[mp], b = mask_pstart[mp + 1]; if (a == b) { a = 0; b = (u32)m; } xor_argmin_range(masks, a, b, rv);
        u64 left = rv & ~low; issued++;
        if (probe_may_hit(P, a, left, ash, r.arank)) { s1 = true; r.kmer = rv; r.qi = (u32)((u64)q * m + i); r.bucket_dir = a | 0x80000000u; r.pad = 0; r.lo = rr; r.n = c; }
        if (DUMP) { cap.smask[(u64)q * m + i] = a; owner_g[o + rr] = i; } } }
    emit(s1, r); }
  if (stats) { for (int of = 16; of; of >>= 1) issued += __shfl_xor_sync(FULLMASK, issued, of); if (lane == 0 && issued) atomicAdd((unsigned long long*)&stats[0], (unsigned long long)issued); }
}

// K2 phase A': the surviving probes regrouped by mask bucket (counting sort on the bucket id: histogram, one-CTA scan, scatter). The probes leave the
// capture kernel in query order, i.e. scattered over all m buckets: every lookup then costs three to four random 64-byte DRAM bursts (start of the anchor,
// first key, next key) and the kernel sits at the DRAM burst rate (~79 G/s measured) while moving four times the bytes it needs. Grouped by bucket, a warp's
// 32 probes search the SAME bucket: its entries (tens of KB) and anchor starts are fetched from DRAM once and then hit in L2, and the scatter itself
// writes m sequential streams that L2 merges into full lines.
__global__ void __launch_bounds__(256) k_surv_hist(const Surv* __restrict__ s, u32 ns, u32* __restrict__ cnt) { const u32 t = blockIdx.x * blockDim.x + threadIdx.x; if (t < ns) atomicAdd(&cnt[s[t].bucket_dir & 0x7FFFFFFFu], 1u); }
__global__ void __launch_bounds__(1024) k_bucket_scan(u32* __restrict__ cnt, int m) {   /* in-place exclusive prefix sum over m counters, one CTA */
  __shared__ u32 s_part[1024]; const int per = (m + 1023) / 1024, b = threadIdx.x * per, e = min(m, b + per); u32 sum = 0; for (int i = b; i < e; i++) sum += cnt[i]; s_part[threadIdx.x] = sum; __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) { u32 v = threadIdx.x >= o ? s_part[threadIdx.x - o] : 0; __syncthreads(); s_part[threadIdx.x] += v; __syncthreads(); }
  u32 run = s_part[threadIdx.x] - sum; for (int i = b; i < e; i++) { const u32 c = cnt[i]; cnt[i] = run; run += c; } }
__global__ void __launch_bounds__(256) k_surv_scatter(const Surv* __restrict__ s, u32 ns, u32* __restrict__ cur, Surv* __restrict__ out) { const u32 t = blockIdx.x * blockDim.x + threadIdx.x; if (t >= ns) return;
  const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(s + t), b = *(reinterpret_cast<const ulonglong2*>(s + t) + 1); const u32 pos = atomicAdd(&cur[(u32)(a.y >> 32) & 0x7FFFFFFFu], 1u);
  ulonglong2* o = reinterpret_cast<ulonglong2*>(out + pos); o[0] = a; o[1] = b; }
// K2 phase B, the roofline kernel: one thread per surviving probe. Anchor start (one random 32-byte sector), galloping + binary lower bound
// over the bucket's 16-byte entries (two per sector), then the range walk over the same records: the entry that ends the search already
// holds the first-value flag and the value count, so no second or third array is touched (round 1 read keys, val_off and vals[v0]).
// STATS additionally counts, per probe, the size n_a of its anchor's run of entries (for the SURVEY.md §8d byte model) — untimed passes only.
template <bool STATS>
__global__ void __launch_bounds__(256) k_probe_find2(ProbeParams P, const Surv* __restrict__ surv, u32 ns, ProbeHit* __restrict__ hits, u32* __restrict__ nhits, u32 cap_hits, u64* __restrict__ stats) {
  u32 t = blockIdx.x * blockDim.x + threadIdx.x; bool have = false; ProbeHit h; u32 steps = 0, ne = 0, lg = 0, hsec = 0, nout = 0;
  if (t < ns) { const u64 pol_tab = l2_policy(P.hints ? 2 : 0), pol_str = l2_policy(P.hints ? 1 : 0);
    Surv sv; { const ulonglong2 a = ld_hint_v2(surv + t, pol_str), b = ld_hint_v2((const u8*)(surv + t) + 16, pol_str); sv.kmer = a.x; sv.qi = (u32)a.y; sv.bucket_dir = (u32)(a.y >> 32); sv.lo = (u32)b.x; sv.n = (u32)(b.x >> 32); sv.arank = (u32)b.y; }
    const u32 dir = sv.bucket_dir >> 31, bucket = sv.bucket_dir & 0x7FFFFFFFu; u64 kmer = sv.kmer;
    int s2 = (P.k - P.p) << 1; u64 low = (P.p < P.k) ? ((1ull << s2) - 1) : 0; u64 left = kmer & ~low, right = kmer | low; const u32 as = anchor_start_of(P, bucket, sv.arank, pol_tab);
    const u64 b0 = ld_hint_u64(P.bucket_off + bucket, pol_tab), b1 = ld_hint_u64(P.bucket_off + bucket + 1, pol_tab); const SeedEntry* __restrict__ E = P.entries; u64 lo = b0 + as, hi = b1;
    auto keyat = [&](u64 i) { return ld_hint_u64(&E[i].key, pol_str); };
    u64 l = lo, r = lo;   // lower bound of `left` in [lo, hi): the anchor's first key is tested first (after the Bloom filter most probes end right there: one sector), then gallop + binary search
    if (keyat(lo) < left) { u64 step = 1; while (l + step < hi && keyat(l + step) < left) { l += step; step <<= 1; steps++; } r = min(hi, l + step); l = l + 1;
      while (l < r) { u64 mid = (l + r) >> 1; if (keyat(mid) < left) l = mid + 1; else r = mid; steps++; } }
    u64 e0 = l; u32 na = 0;
    while (e0 + ne < hi) { const ulonglong2 raw = ld_hint_v2(E + e0 + ne, pol_str); if (raw.x > right) break; const u32 nf = (u32)(raw.y >> 32); if ((nf >> 31) == dir) na += nf & 0x7FFFFFFFu; ne++; }
    if (na) { have = true; h.q = sv.qi / (u32)P.m; h.mask_dir = (u32)((sv.qi % (u32)P.m) << 1 | dir); h.e0 = e0; h.ne = ne; h.lo = sv.lo; h.n = sv.n; h.kmer = kmer; h.nanch = na * sv.n; h.bucket = bucket; }
    if (STATS) { const int ash = (P.k - P.mask_prefix - P.anchor_prefix) << 1; const u64 an = E[lo].key >> ash; u32 n_a = 0; while (lo + n_a < hi && (E[lo + n_a].key >> ash) == an) n_a++; lg = 32 - __clz(n_a); hsec = (16 * ne + 31) / 32; nout = na; }   // ceil(log2(n_a + 1)) = bit length of n_a
  }
  // hit records are appended with ONE global atomic per CTA (warp ballots -> shared-memory warp counts -> thread 0): one atomic per warp put ~400,000 same-address atomics per launch on the critical path
  __shared__ u32 s_wcnt[8], s_base; const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5; const u32 bal = __ballot_sync(FULLMASK, have); if (lane == 0) s_wcnt[wid] = __popc(bal); __syncthreads();
  if (threadIdx.x == 0) { u32 tot = 0; for (int i = 0; i < 8; i++) { const u32 c = s_wcnt[i]; s_wcnt[i] = tot; tot += c; } s_base = tot ? atomicAdd(nhits, tot) : 0u; } __syncthreads();
  if (have) { const u32 w = s_base + s_wcnt[wid] + __popc(bal & ((1u << lane) - 1)); if (w < cap_hits) hits[w] = h; }
  if (STATS && stats) { for (int o = 16; o; o >>= 1) { steps += __shfl_xor_sync(FULLMASK, steps, o); ne += __shfl_xor_sync(FULLMASK, ne, o); lg += __shfl_xor_sync(FULLMASK, lg, o); hsec += __shfl_xor_sync(FULLMASK, hsec, o); nout += __shfl_xor_sync(FULLMASK, nout, o); }
    if (lane == 0) { atomicAdd((unsigned long long*)&stats[2], (unsigned long long)steps); atomicAdd((unsigned long long*)&stats[3], (unsigned long long)ne); atomicAdd((unsigned long long*)&stats[4], (unsigned long long)lg); atomicAdd((unsigned long long*)&stats[5], (unsigned long long)hsec); atomicAdd((unsigned long long*)&stats[6], (unsigned long long)nout); } }
}

// ---- seed-lookup microbenchmark (BASELINE.json configs[4]; SURVEY.md §8d "C5"): query 31-mers -> one prefix probe and one suffix probe each,
// exactly the records k_capture2 emits for real queries. Half of the queries are stored keys of a random bucket with their last 0-16 bases
// randomised (LCP with the stored key in [15, 31]); half are uniform random k-mers (bucket = the mask that would capture them).
// Only probes whose bucket lies in [mask_lo, mask_hi) are kept: with the index range-partitioned by mask every GPU keeps its own share.
__global__ void __launch_bounds__(256) k_c5_gen(ProbeParams P, const u64* __restrict__ masks, const u32* __restrict__ mask_pstart, int mask_pbits, u64 per, u64 iseed, u64 qseed, u64 nq, int mask_lo, int mask_hi,
                                                 Surv* __restrict__ surv, u32* __restrict__ nsurv, u32 cap, u64* __restrict__ stats) {
  const u64 q = blockIdx.x * (u64)blockDim.x + threadIdx.x; bool s0 = false, s1 = false; Surv a, b; u32 issued = 0; const int k = P.k; const int lane = threadIdx.x & 31;
  if (q < nq) { u64 r = mix64(qseed ^ (q * 0x9E3779B97F4A7C15ull)); u64 key; u32 bucket; const int msh = 2 * k - mask_pbits;
    auto argmin_mask = [&](u64 x) { u32 mp = (u32)(x >> msh); u32 lo = mask_pstart[mp], hi = mask_pstart[mp + 1]; if (lo == hi) { lo = 0; hi = (u32)P.m; } xor_argmin_range(masks, lo, hi, x); return lo; };
    if (r & 1) { bucket = (u32)((r >> 1) % (u64)P.m); const u64 j = mix64(r) % per; key = synth_key(masks[bucket], P.mask_prefix, k, per, j, iseed, bucket); const u64 r2 = mix64(r ^ 0x5bd1e995ull); const int t = (int)(r2 % 17); if (t) key ^= (r2 >> 8) & ((1ull << (2 * t)) - 1); }
    else { key = r >> 2; bucket = argmin_mask(key); }
    const int s2 = (k - P.p) << 1; const u64 low = (P.p < k) ? ((1ull << s2) - 1) : 0; const int ash = (k - P.mask_prefix - P.anchor_prefix) << 1;
    if ((int)bucket >= mask_lo && (int)bucket < mask_hi) { issued++; const u64 left = key & ~low;
      if (probe_may_hit(P, bucket, left, ash, a.arank)) { s0 = true; a.kmer = key; a.qi = (u32)q; a.bucket_dir = bucket; a.pad = 0; a.lo = 0; a.n = 1; } }
    const u64 rv = kmer_reverse62(key, k); const u32 b2 = argmin_mask(rv);
    if ((int)b2 >= mask_lo && (int)b2 < mask_hi) { issued++; const u64 left = rv & ~low;
      if (probe_may_hit(P, b2, left, ash, b.arank)) { s1 = true; b.kmer = rv; b.qi = (u32)q; b.bucket_dir = b2 | 0x80000000u; b.pad = 0; b.lo = 0; b.n = 1; } } }
  const u32 b0 = __ballot_sync(FULLMASK, s0), b1 = __ballot_sync(FULLMASK, s1), tot = __popc(b0) + __popc(b1);
  if (tot) { u32 base = 0; if (lane == 0) base = atomicAdd(nsurv, tot); base = __shfl_sync(FULLMASK, base, 0);
    if (s0) { u32 w = base + __popc(b0 & ((1u << lane) - 1)); if (w < cap) surv[w] = a; if (P.bcnt) atomicAdd(&P.bcnt[a.bucket_dir & 0x7FFFFFFFu], 1u); } if (s1) { u32 w = base + __popc(b0) + __popc(b1 & ((1u << lane) - 1)); if (w < cap) surv[w] = b; if (P.bcnt) atomicAdd(&P.bcnt[b.bucket_dir & 0x7FFFFFFFu], 1u); } }
  for (int o = 16; o; o >>= 1) issued += __shfl_xor_sync(FULLMASK, issued, o); if (lane == 0 && issued) atomicAdd((unsigned long long*)&stats[0], (unsigned long long)issued);
}

__global__ void k_hit_counts(const ProbeHit* __restrict__ h, u32 n, u64* __restrict__ c) { u32 t = blockIdx.x * blockDim.x + threadIdx.x; if (t <= n) c[t] = (t < n) ? h[t].nanch : 0; }

// anchor keys: hi = query<<36 | genome(dense)<<2 | (sorted only on bits >= 2) ; lo = QBegin<<36 | (63-Len)<<30 | TBegin<<2 | qrc<<1 | trc
__device__ __forceinline__ u64 pack_lo(i32 qb, u32 len, i32 tb, u32 qrc, u32 trc) { return ((u64)(u32)qb << 36) | ((u64)(63 - len) << 30) | ((u64)((u32)tb & 0x0FFFFFFFu) << 2) | (qrc << 1) | trc; }
__device__ __forceinline__ u64 pack_hi(u32 q, u32 g) { return ((u64)q << 36) | ((u64)g << 2); }

// one thread per hit: (matched key) x (query locations) x (values) -> anchors (lib-index-search.go:1398-1557)
__global__ void k_probe_emit(ProbeParams P, const ProbeHit* __restrict__ hits, const u64* __restrict__ hoff, u32 nh, const u32* __restrict__ qvals, const u64* __restrict__ koff,
                             const u32* __restrict__ batch_base, u64* __restrict__ a_hi, u64* __restrict__ a_lo) {
  u32 t = blockIdx.x * blockDim.x + threadIdx.x; if (t >= nh) return; ProbeHit h = hits[t]; u64 w = hoff[t]; const int K = P.k; const u32 want = h.mask_dir & 1; const u32* locs = qvals + koff[h.q] + h.lo; const u64* V = P.vals + P.bucket_voff[h.bucket];
  for (u32 e = 0; e < h.ne; e++) { const SeedEntry se = P.entries[h.e0 + e]; const u32 nv = se.nflag & 0x7FFFFFFFu; if (nv == 0 || (se.nflag >> 31) != want) continue; const u64 key = se.key;
    int len = (__clzll(h.kmer ^ key) >> 1) + K - 32; if (h.kmer == key) len = K;   // Len = LZ(q^kmer)/2 + k - 32 (kv-searcher.go:480)
    for (u32 li = 0; li < h.n; li++) { u32 loc = locs[li] & 0x7fffffffu; u32 rcQ = loc & 1; i32 posQ = (i32)(loc >> 1);
      for (u32 vi = 0; vi < nv; vi++) { u64 rp = V[se.vrel + vi]; u64 bgi = rp >> 30; u32 g = batch_base[bgi >> 17] + (u32)(bgi & 0x1ffff); i32 posT = (i32)((rp << 34) >> 36); u32 rv = rp & 1, rcT = (rp >> 1) & 1; i32 bq, bt;
        if (!rv) { bq = rcQ ? posQ + K - len : posQ; bt = rcT ? posT + K - len : posT; } else { bq = rcQ ? posQ : posQ + K - len; bt = rcT ? posT : posT + K - len; }
        a_hi[w] = pack_hi(h.q, g); a_lo[w] = pack_lo(bq, (u32)len, bt, rcQ, rcT); w++; } } }
}

// =====================================================================================================
// engine object
// =====================================================================================================
struct QBatch {  // device-side query batch
  int nq = 0; u64 total_bases = 0, total_k = 0;
  DBuf<u8> ascii, packed, amask; DBuf<u64> off, boff, koff; std::vector<u64> h_off, h_boff, h_koff; std::vector<u8> h_ascii;   // h_ascii: host copy of the query bytes (alignment text of -a output)
  DBuf<u64> qkeys; DBuf<u32> qvals;    // per-query sorted (k-mer, loc) tables
};

// Host worker pool of one search context. The host phases between kernels (window geometry, contig mapping, scoring, row building) are
// split into static chunks; workers sleep on a condition variable between phases — OpenMP teams spin at their barriers, and with several
// lanes per call the spinning teams oversubscribed the cores and produced 30-50 ms stalls.
struct HostPool {
  std::vector<std::thread> th; std::mutex mu; std::condition_variable cv_go, cv_done; std::function<void(int)> fn; int nchunks = 0, next = 0, pending = 0; u64 gen = 0; bool quit = false; std::exception_ptr err;
  void start(int n) { for (int i = 0; i < n; i++) th.emplace_back([this] { u64 seen = 0; for (;;) { std::unique_lock<std::mutex> lk(mu); cv_go.wait(lk, [&] { return quit || gen != seen; }); if (quit) return; seen = gen; work(lk); } }); }
  void work(std::unique_lock<std::mutex>& lk) { while (next < nchunks) { int c = next++; lk.unlock(); try { fn(c); } catch (...) { lk.lock(); if (!err) err = std::current_exception(); lk.unlock(); } lk.lock(); if (--pending == 0) cv_done.notify_all(); } }
  // runs f(0..n-1), the caller takes part; returns when every chunk is done
  void run(int n, int max_workers, const std::function<void(int)>& f) { if (n <= 0) return; if (n == 1 || max_workers <= 1) { for (int c = 0; c < n; c++) f(c); return; }
    if ((int)th.size() < max_workers - 1) start(max_workers - 1 - (int)th.size());
    std::unique_lock<std::mutex> lk(mu); fn = f; nchunks = n; next = 0; pending = n; err = nullptr; gen++; cv_go.notify_all(); work(lk); cv_done.wait(lk, [&] { return pending == 0; }); nchunks = 0; std::exception_ptr e = err; lk.unlock(); if (e) std::rethrow_exception(e); }
  ~HostPool() { { std::lock_guard<std::mutex> lk(mu); quit = true; } cv_go.notify_all(); for (auto& t : th) t.join(); }
};

// One search context: the shared HBM image plus this context's stream, arena, scratch and timers. The handle returned by lmg_index_open
// is lane 0 and owns the image; further lanes (same image, own stream/arena) are created on demand so that big batches can be split into
// sub-batches whose host phases (window geometry, contig mapping, scoring) overlap the other sub-batch's kernels.
struct lmg_index {
  Image* imgp; Image& img; bool owner; cudaStream_t st = 0; cudaStream_t st_hi = 0; cudaEvent_t ev_hi[2] = {nullptr, nullptr};   /* st_hi: high-priority stream of this lane for the memory-bound index lookup */ CubTemp tmp; int sm_count = 148; u32 smem_optin = 0; int use_tma = 1;
  double ms[16] = {0}; u64 counters[16] = {0}; u64 pstat[4] = {0, 0, 0, 0};   /* last statistics pass of the seed lookup: sum ceil(log2(n_a+1)), sum of 32-byte sectors of matched entries, sum of values of matched entries */ std::mutex mu; cudaEvent_t kev[3] = {nullptr, nullptr, nullptr}; Arena arena; std::vector<lmg_index*> lanes; int lane_id = 0, active_lanes = 1; size_t total_mem = 0; HostPool pool;
  // workers per lane: this process's usable cores over the active lanes. LMG_HOST_CORES (or OMP_NUM_THREADS, which launchers such as torchrun set per
  // rank) tells how many cores the process may use when several ranks share a node; otherwise the affinity mask capped by the cgroup CPU quota
  // (a GPU lease of a big box shows all its cores in hardware_concurrency() but schedules only the quota: oversubscribed pools stall every lane).
  static int usable_cores() { int v = 0; cpu_set_t cs; CPU_ZERO(&cs); if (sched_getaffinity(0, sizeof cs, &cs) == 0) v = CPU_COUNT(&cs); if (v <= 0) v = (int)std::thread::hardware_concurrency();
    if (FILE* f = fopen("/sys/fs/cgroup/cpu.max", "r")) { char a[64]; double per = 0; if (fscanf(f, "%63s %lf", a, &per) == 2 && strcmp(a, "max") != 0 && per > 0) { int q = (int)(atof(a) / per + 0.5); if (q > 0) v = std::min(v, q); } fclose(f); }
    return v > 0 ? v : 8; }
  int host_threads() const { static const int hc = [] { int v = 0; if (const char* e = getenv("LMG_HOST_CORES")) v = atoi(e); if (v <= 0) if (const char* e = getenv("OMP_NUM_THREADS")) v = atoi(e); if (v <= 0) v = usable_cores(); return v > 0 ? v : 8; }();
    return std::max(1, std::min(32, hc / std::max(1, active_lanes))); }
  lmg_index(Image* p, bool own) : imgp(p), img(*p), owner(own) {}
};

static thread_local std::string g_err;
// a sub-batch whose intermediate lists exceed a 2^31 index space or the device arena: the caller halves it and retries
struct BatchTooLarge : std::runtime_error { using std::runtime_error::runtime_error; };
// host-side lap timer (LMG_DEBUG_TIMING=1): prints the wall time since the previous lap of the calling thread
struct LapTimer { std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now(); bool on = getenv("LMG_DEBUG_TIMING") != nullptr; int lane = 0;
  void operator()(const char* what) { if (!on) return; auto n = std::chrono::steady_clock::now(); fprintf(stderr, "[lmg host L%d] %-18s %.2f ms\n", lane, what, std::chrono::duration<double, std::milli>(n - t0).count()); t0 = n; } };
static thread_local LapTimer* g_lap = nullptr;
struct KTimer { cudaEvent_t a, b; cudaStream_t st; double* dst; KTimer(cudaStream_t s, double* d) : st(s), dst(d) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, st); } ~KTimer() { cudaEventRecord(b, st); cudaEventSynchronize(b); float f = 0; cudaEventElapsedTime(&f, a, b); *dst += f; cudaEventDestroy(a); cudaEventDestroy(b); } };
// activates the index's arena for the calling thread and rewinds it when the batch is done (all DBufs of the batch are dead by then)
struct ArenaReset { lmg_index* ix; ArenaScope sc; ArenaReset(lmg_index* i) : ix(i), sc(&i->arena) {} ~ArenaReset() { cudaStreamSynchronize(ix->st); ix->arena.reset(); } };

static void upload_queries(lmg_index* ix, const u8* seqs, const u64* off, int nq, QBatch& B) {
  cudaStream_t st = ix->st; const int k = ix->img.k; B.nq = nq; B.h_off.assign(off, off + nq + 1); B.total_bases = off[nq] - off[0];
  if (off[0] != 0) for (auto& x : B.h_off) x -= off[0];
  B.h_boff.resize(nq + 1); B.h_koff.resize(nq + 1); u64 b = 0, kk = 0;
  for (int q = 0; q < nq; q++) { u64 L = B.h_off[q + 1] - B.h_off[q]; if (L >= (1ull << 27)) throw std::runtime_error("query longer than 2^27 bases is not supported");
    B.h_boff[q] = b; b += (((L + 3) >> 2) + 16 + 15) & ~15ull; B.h_koff[q] = kk; kk += (L >= (u64)k) ? 2 * (L - k + 1) : 0; }
  B.h_boff[nq] = b; B.h_koff[nq] = kk; B.total_k = kk;
  // 32-bit index spaces of one sub-batch: (query, mask) slots and k-mer table rows. Larger batches are halved by the caller (search_range).
  if ((u64)nq * (u64)ix->img.m >= (1ull << 32) || kk >= (1ull << 31)) throw BatchTooLarge("sub-batch exceeds the 2^32 (query, mask) slots or 2^31 k-mer rows of one lane");
  B.h_ascii.assign(seqs + off[0], seqs + off[0] + B.total_bases); B.ascii.alloc(B.total_bases + 16, st); B.ascii.from_host(B.h_ascii.data(), B.total_bases); B.off.alloc(nq + 1, st); B.off.from_host(B.h_off.data(), nq + 1);
  B.boff.alloc(nq + 1, st); B.boff.from_host(B.h_boff.data(), nq + 1); B.koff.alloc(nq + 1, st); B.koff.from_host(B.h_koff.data(), nq + 1);
  B.packed.alloc(b + 64, st); B.packed.zero(); B.amask.alloc(b + 64, st); B.amask.zero();
}

// K1a: pack + k-mers + per-query stable sort by k-mer
static void sketch_tables(lmg_index* ix, QBatch& B) {
  cudaStream_t st = ix->st; const int k = ix->img.k; int nq = B.nq; u64 maxL = 0; for (int q = 0; q < nq; q++) maxL = std::max(maxL, B.h_off[q + 1] - B.h_off[q]);
  dim3 g1((unsigned)std::max(1, std::min(64, cdiv((i64)(maxL + 3) / 4, 256))), nq); k_pack_queries<<<g1, 256, 0, st>>>(B.ascii.p, B.off.p, B.boff.p, B.packed.p, B.amask.p, nq); KERNEL_CHECK();
  DBuf<u64> keys_in(B.total_k + 2, st); DBuf<u32> vals_in(B.total_k + 2, st); B.qkeys.alloc(B.total_k + 2, st); B.qvals.alloc(B.total_k + 2, st);
  if (B.total_k == 0) return;
  dim3 g2((unsigned)std::max(1, std::min(64, cdiv((i64)maxL, 128))), nq); k_gen_kmers<<<g2, 128, 0, st>>>(B.packed.p, B.boff.p, B.off.p, B.koff.p, keys_in.p, vals_in.p, nq, k); KERNEL_CHECK();
  size_t tb = 0; cub::DeviceSegmentedSort::StableSortPairs(nullptr, tb, keys_in.p, B.qkeys.p, vals_in.p, B.qvals.p, (int)B.total_k, nq, B.koff.p, B.koff.p + 1, st);
  cub::DeviceSegmentedSort::StableSortPairs(ix->tmp.get(tb), tb, keys_in.p, B.qkeys.p, vals_in.p, B.qvals.p, (int)B.total_k, nq, B.koff.p, B.koff.p + 1, st); CUB_CHECK();
}

// K1b: capture
struct CapBufs { DBuf<u64> kmer; DBuf<u32> lo, n, smask; CapSoA soa() { CapSoA c; c.kmer = kmer.p; c.lo = lo.p; c.n = n.p; c.smask = smask.p; return c; } void free() { kmer.free(); lo.free(); n.free(); smask.free(); } };
struct Anchors { u64 n = 0; DBuf<u64> hi, lo; };

static ProbeParams probe_params(const Image& I, int p) { ProbeParams P; P.bucket_off = I.d_bucket_off; P.bucket_voff = I.d_bucket_voff; P.entries = I.d_entries; P.vals = I.d_vals; P.anchor_cbase = I.d_anchor_cbase; P.anchor_cstart = I.d_anchor_cstart; P.anchor_cum = I.d_anchor_cum; P.anchor_bits = I.d_anchor_bits; P.pbloom = I.d_pbloom; P.pbmask = I.pbmask; P.m = I.m; P.k = I.k; P.NA = I.NA; P.mask_prefix = I.mask_prefix; P.anchor_prefix = I.anchor_prefix; P.p = p; P.anchor_cstart16 = I.d_anchor_cstart16; P.bcnt = nullptr; static const int hints = getenv("LMG_L2_HINTS") ? 1 : 0; P.hints = hints; return P; }   /* measured on C2: evict-first entries + evict-last tables lower the L2 hit rate (27 % -> 19 %) and cost 6 %: off unless asked for */

template <class K, class V> static void radix_sort_pairs(lmg_index* ix, DBuf<K>& k_in, DBuf<K>& k_out, DBuf<V>& v_in, DBuf<V>& v_out, u64 n, int begin_bit, int end_bit) {
  size_t tb = 0; cub::DeviceRadixSort::SortPairs(nullptr, tb, k_in.p, k_out.p, v_in.p, v_out.p, (i64)n, begin_bit, end_bit, ix->st);
  cub::DeviceRadixSort::SortPairs(ix->tmp.get(tb), tb, k_in.p, k_out.p, v_in.p, v_out.p, (i64)n, begin_bit, end_bit, ix->st); CUB_CHECK();
}

static int bits_for(u64 v) { int b = 1; while ((v >> b) && b < 64) b++; return b; }

// K1b + K2 phase A: surviving probes of the batch. Fused kernel when every query table fits in shared memory next to its owner array,
// otherwise (long queries) capture into HBM arrays with several CTAs per query, then the separate filter kernel.
struct Survivors { DBuf<Surv> d; DBuf<u32> bcnt; u32 n = 0; };   // bcnt: survivors per bucket, counted while they are emitted
static void probe_survivors(lmg_index* ix, QBatch& B, const lmg_params* prm, Survivors& SV, CapBufs* dump_cap, DBuf<u32>* dump_owner) {
  cudaStream_t st = ix->st; const Image& I = ix->img; if (prm->min_prefix < I.mask_prefix + I.anchor_prefix || prm->min_prefix > I.k) throw std::runtime_error("the minimum prefix length should be in the range of [maskPrefix+anchorPrefix, k]");  // kv-searcher.go:202
  ProbeParams P = probe_params(I, prm->min_prefix); const u64 nslot = (u64)B.nq * I.m, nprobe = nslot * 2; DBuf<u32> nsv(1, st); DBuf<u64> dstats(8, st); SV.bcnt.alloc((u64)I.m + 1, st); P.bcnt = SV.bcnt.p;
  if (!ix->kev[0]) { cudaEventCreate(&ix->kev[0]); cudaEventCreate(&ix->kev[1]); cudaEventCreate(&ix->kev[2]); }
  u64 maxn = 0; for (int q = 0; q < B.nq; q++) maxn = std::max(maxn, B.h_koff[q + 1] - B.h_koff[q]);
  const bool fused = maxn <= 16384 && maxn * 12 + 16384 <= ix->smem_optin && !getenv("LMG_NO_FUSED_CAPTURE"); ix->ms[8] = 0;
  const int cthreads = maxn <= 4096 ? 256 : 1024;   // big tables leave room for one CTA per SM only: make it a full one
  if (dump_cap) { dump_cap->kmer.alloc(nslot, st); dump_cap->lo.alloc(nslot, st); dump_cap->n.alloc(nslot, st); dump_cap->smask.alloc(nslot, st); dump_owner->alloc(B.total_k + 2, st); dump_owner->fill_ff(); }
  CapBufs capl; DBuf<u32> ownl; CapBufs* cap = dump_cap ? dump_cap : &capl; DBuf<u32>* owner = dump_owner ? dump_owner : &ownl;
  if (!fused) { if (!dump_cap) { capl.kmer.alloc(nslot, st); capl.lo.alloc(nslot, st); capl.n.alloc(nslot, st); capl.smask.alloc(nslot, st); ownl.alloc(B.total_k + 2, st); ownl.fill_ff(); }
    u32 smem_cap = (u32)std::min<u64>((ix->smem_optin - 1024) / 8, 24576); u32 need = (u32)std::min<u64>(maxn, smem_cap); size_t smem = ((size_t)need * 8 + 15) & ~15ull;
    // slices: enough CTAs to fill 148 SMs a few times over, but keep >= 1024 masks per CTA so the staged table is reused
    int slices = std::max(1, std::min(I.m / 1024, cdiv(ix->sm_count * 8, std::max(1, B.nq))));
    k_capture<<<B.nq * slices, 256, smem, st>>>(B.qkeys.p, B.koff.p, I.d_masks, I.m, I.k, slices, cap->soa(), owner->p, need, ix->use_tma, I.d_mask_pstart, I.mask_pbits); KERNEL_CHECK(); }
  // capacity: a quarter of all probes first (about 10 % survive on the bench workload), everything on overflow
  u64 capS = std::max<u64>(1u << 20, nprobe / 4);
  for (int attempt = 0; attempt < 2; attempt++) { SV.d.alloc(capS, st); nsv.zero(); dstats.zero(); SV.bcnt.zero();
    if (fused) { u32 mx = (u32)((maxn + 1) & ~1ull); size_t smem = (size_t)mx * 12 + 16;
      if (dump_cap) { k_capture2<true><<<B.nq, cthreads, smem, st>>>(B.qkeys.p, B.koff.p, I.d_masks, P, cap->soa(), owner->p, mx, ix->use_tma, I.d_mask_pstart, I.mask_pbits, SV.d.p, nsv.p, (u32)std::min<u64>(capS, 0xFFFFFFFFu), dstats.p); KERNEL_CHECK(); }
      else { k_capture2<false><<<B.nq, cthreads, smem, st>>>(B.qkeys.p, B.koff.p, I.d_masks, P, cap->soa(), nullptr, mx, ix->use_tma, I.d_mask_pstart, I.mask_pbits, SV.d.p, nsv.p, (u32)std::min<u64>(capS, 0xFFFFFFFFu), dstats.p); KERNEL_CHECK(); } }
    else { cudaEventRecord(ix->kev[0], st); k_probe_filter<<<cdiv((i64)nslot, 256), 256, 0, st>>>(P, cap->soa(), owner->p, B.koff.p, nslot, SV.d.p, nsv.p, (u32)std::min<u64>(capS, 0xFFFFFFFFu), dstats.p); KERNEL_CHECK(); cudaEventRecord(ix->kev[1], st); }
    SV.n = nsv.to_host()[0]; if (SV.n <= capS) break; capS = nprobe; }
  if (!fused) { float fa = 0; cudaEventSynchronize(ix->kev[1]); cudaEventElapsedTime(&fa, ix->kev[0], ix->kev[1]); ix->ms[8] = fa; ix->counters[12] = (u64)(fa * 1000); } else ix->counters[12] = 0;
  ix->counters[0] = dstats.to_host()[0]; ix->counters[1] = SV.n; ix->counters[8] = nprobe;
}

// regroup `in` by bucket into `out` on stream sl (same capacity); LMG_NO_REGROUP keeps the capture order
static bool regroup_survivors(cudaStream_t sl, const Surv* in, u32 ns, int m, Surv* out, u32* bcnt, bool have_hist) {   /* have_hist: the kernel that emitted the probes already counted them per bucket (ProbeParams::bcnt); bcnt is consumed */
  static const bool off = getenv("LMG_NO_REGROUP") != nullptr; if (off || ns == 0) return false;
  if (!have_hist) { CUDA_CHECK(cudaMemsetAsync(bcnt, 0, ((size_t)m + 1) * 4, sl)); k_surv_hist<<<cdiv(ns, 256), 256, 0, sl>>>(in, ns, bcnt); KERNEL_CHECK(); }
  k_bucket_scan<<<1, 1024, 0, sl>>>(bcnt, m); KERNEL_CHECK();
  k_surv_scatter<<<cdiv(ns, 256), 256, 0, sl>>>(in, ns, bcnt, out); KERNEL_CHECK(); return true; }
// K2 phase B: index lookup on the surviving probes -> anchors sorted by (query, genome, QBegin, QEnd desc, TBegin, qrc, trc)
static void seed_probe(lmg_index* ix, QBatch& B, const lmg_params* prm, Survivors& SV, Anchors& A, bool stats) {
  cudaStream_t st = ix->st; const Image& I = ix->img; ProbeParams P = probe_params(I, prm->min_prefix); DBuf<u32> nh(1, st); DBuf<u64> dstats(8, st); dstats.zero(); const u32 ns = SV.n; DBuf<Surv>& surv = SV.d;
  // phase B: index lookup on the survivors
  DBuf<ProbeHit> hits; u64 capH = std::max<u64>(1u << 18, (u64)ns / 2 + 1024); u32 nhit = 0;
  cudaStream_t sl = getenv("LMG_NO_PRIO_LOOKUP") ? st : ix->st_hi; DBuf<Surv> grouped; DBuf<u32> bcnt; const Surv* sp = surv.p; float fr = 0; bool fr_valid = ns != 0;
  if (ns) { grouped.alloc(ns, st); const bool hh = SV.bcnt.p != nullptr; if (!hh) bcnt.alloc((u64)I.m + 1, st); if (sl != st) { cudaEventRecord(ix->ev_hi[0], st); cudaStreamWaitEvent(sl, ix->ev_hi[0], 0); }
    cudaEventRecord(ix->kev[0], sl); if (regroup_survivors(sl, surv.p, ns, I.m, grouped.p, hh ? SV.bcnt.p : bcnt.p, hh)) sp = grouped.p; }
  for (int attempt = 0; attempt < 2 && ns; attempt++) { hits.alloc(capH, st); nh.zero(); if (attempt) CUDA_CHECK(cudaMemsetAsync(dstats.p + 2, 0, 40, st));
    // The lookup is the one memory-bound kernel of the path: it runs on the lane's HIGH-PRIORITY stream so that its CTAs are scheduled ahead of the issue-bound
    // kernels of the other lanes that share the GPU (they would otherwise stretch it 2x without gaining anything themselves).
    if (attempt) fr_valid = false;   /* kev[1] is recorded again: the regroup time of this (rare, hit-buffer overflow) call is not reported */
    if (sl != st) { cudaEventRecord(ix->ev_hi[0], st); cudaStreamWaitEvent(sl, ix->ev_hi[0], 0); }   /* hits / counters were (re)allocated and cleared on st */
    cudaEventRecord(ix->kev[1], sl); if (stats) k_probe_find2<true><<<cdiv(ns, 256), 256, 0, sl>>>(P, sp, ns, hits.p, nh.p, (u32)std::min<u64>(capH, 0xFFFFFFFFu), dstats.p); else k_probe_find2<false><<<cdiv(ns, 256), 256, 0, sl>>>(P, sp, ns, hits.p, nh.p, (u32)std::min<u64>(capH, 0xFFFFFFFFu), nullptr); KERNEL_CHECK(); cudaEventRecord(ix->kev[2], sl);
    if (sl != st) { cudaEventRecord(ix->ev_hi[1], sl); cudaStreamWaitEvent(st, ix->ev_hi[1], 0); }
    nhit = nh.to_host()[0]; if (nhit <= capH) break; capH = (u64)ns + 1024; }
  if (!hits.p) hits.alloc(16, st);
  { float fb = 0; if (ns) { cudaEventSynchronize(ix->kev[2]); cudaEventElapsedTime(&fb, ix->kev[1], ix->kev[2]); } ix->ms[8] += fb; ix->counters[13] = (u64)(fb * 1000); if (fr_valid) { cudaEventElapsedTime(&fr, ix->kev[0], ix->kev[1]); ix->counters[12] += (u64)(fr * 1000); } }   /* [12]: filter kernel of the non-fused capture + the regrouping pass */
  { auto sdt = dstats.to_host(); if (stats) { ix->counters[2] = sdt[2]; ix->counters[3] = sdt[3]; ix->pstat[0] = sdt[4]; ix->pstat[1] = sdt[5]; ix->pstat[2] = sdt[6]; } ix->counters[4] = nhit; }
  A.n = 0; if (nhit == 0) return;
  DBuf<u64> hoff(nhit + 1, st);
  { DBuf<u64> cnt(nhit + 1, st); k_hit_counts<<<cdiv(nhit + 1, 256), 256, 0, st>>>(hits.p, nhit, cnt.p); KERNEL_CHECK();
    size_t tb = 0; cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt.p, hoff.p, (int)(nhit + 1), st); cub::DeviceScan::ExclusiveSum(ix->tmp.get(tb), tb, cnt.p, hoff.p, (int)(nhit + 1), st); CUB_CHECK(); }
  u64 total; CUDA_CHECK(cudaMemcpyAsync(&total, hoff.p + nhit, 8, cudaMemcpyDeviceToHost, st)); CUDA_CHECK(cudaStreamSynchronize(st));
  A.n = total; if (stats) ix->counters[5] = total; if (total == 0) return; if (total >= (1ull << 31)) throw BatchTooLarge("more than 2^31 anchors in one batch; use smaller batches");
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
#define BT_SMALL 96   // segments up to this many anchors walk their (score,index) keys by repeated selection; larger ones use the CUB-sorted copy
__global__ void k_seg_end(const u64* __restrict__ seg_off, const u32* __restrict__ c_n, u32 nseg, u64* __restrict__ seg_end) { u32 s = blockIdx.x * blockDim.x + threadIdx.x; if (s < nseg) seg_end[s] = seg_off[s] + (c_n[s] > BT_SMALL ? c_n[s] : 0); }
__global__ void k_widen(const u32* __restrict__ a, u32 n, u64* __restrict__ o) { u32 i = blockIdx.x * blockDim.x + threadIdx.x; if (i <= n) o[i] = (i < n) ? a[i] : 0; }
__global__ void k_backtrack(const u64* __restrict__ seg_off, const u32* __restrict__ c_n, u32 nseg, const u64* __restrict__ c_lo, const u32* __restrict__ pred, const signed char* __restrict__ dirs, const u64* __restrict__ s2i_unsorted,
                            const u64* __restrict__ s2i_sorted, u8* __restrict__ visited, ChainParams P, ChainRec* __restrict__ out, u32* __restrict__ nout, u32 cap, float* __restrict__ seg_score) {
  u32 seg = blockIdx.x * blockDim.x + threadIdx.x; if (seg >= nseg) return; u64 b = seg_off[seg]; i32 n = (i32)c_n[seg]; const u64* C = c_lo + b; const u32* Pd = pred + b; const signed char* D = dirs + b; const u64* K2 = s2i_sorted + b; const u64* KU = s2i_unsorted + b; u8* V = visited + b; const bool small = n <= BT_SMALL;
  u32 ord = 0;
  auto emit = [&](i32 first, i32 last, i32 cnt, float sc) { u32 w = atomicAdd(nout, 1u); if (w < cap) { ChainRec r; r.seg = seg; r.ord = ord; u64 f = C[first], l = C[last]; r.q0 = a_q(f); r.t0 = a_t(f); r.len0 = a_len(f); r.q1 = a_q(l); r.t1 = a_t(l); r.len1 = a_len(l); r.flags1 = (u32)(l & 3); r.nseeds = cnt; r.score = sc; r.pad = 0; out[w] = r; } ord++; };
  if (n == 1) { float w = __uint_as_float((u32)(KU[0] >> 32)); seg_score[seg] = w; if (w >= P.min_score) emit(0, 0, 1, w); return; }
  i32 iMax = n - 1; float maxScore = 0; bool first = true; int nChecked = 0; u64 last = ~0ull; bool exhausted = false;
  for (;;) {
    nChecked++; if (P.top_chains > 0 && nChecked > P.top_chains) break;
    float M = 0; u32 Mi = 0;
    if (!small) { while (iMax >= 0) { u64 e = K2[iMax]; M = __uint_as_float((u32)(e >> 32)); Mi = (u32)e; if (!V[Mi]) { iMax--; break; } iMax--; } }
    else { while (!exhausted) { u64 best = 0; bool f = false; for (i32 x = 0; x < n; x++) { u64 e = KU[x]; if (e < last && (!f || e > best)) { best = e; f = true; } } if (!f) { exhausted = true; break; }
        last = best; M = __uint_as_float((u32)(best >> 32)); Mi = (u32)best; if (!V[Mi]) break; } }   // keys are distinct (index in the low word): descending walk == the sorted order
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

// counting sort by a dense integer key, then tiny per-bucket sorts (host lists of chain / HSP records are grouped by segment / window)
template <class T, class KeyFn, class Less> static void bucket_sort(std::vector<T>& v, u32 nkeys, KeyFn key, Less less) {
  if (v.size() < 2) return; std::vector<u32> off(nkeys + 1, 0); for (const T& x : v) off[key(x) + 1]++; for (u32 i = 0; i < nkeys; i++) off[i + 1] += off[i];
  std::vector<T> out(v.size()); std::vector<u32> cur(off.begin(), off.end() - 1); for (const T& x : v) out[cur[key(x)]++] = x;
  for (u32 i = 0; i < nkeys; i++) if (off[i + 1] - off[i] > 1) std::sort(out.begin() + off[i], out.begin() + off[i + 1], less); v.swap(out);
}

struct Segments { u32 nseg = 0; DBuf<u64> key, off; DBuf<u32> cn; DBuf<u64> c_lo; DBuf<float> score; std::vector<u64> h_key, h_off; };
struct Chains { u32 n = 0; DBuf<ChainRec> rec; std::vector<ChainRec> h; std::vector<float> seg_score; std::vector<char> topn_sorted; };   // topn_sorted[q]: the top-N genome sort ran for query q (it fixes the order of equal genome indexes from different batches)

static void chain_stage(lmg_index* ix, const lmg_params* prm, Anchors& A, Segments& S, Chains& Cn) {
  cudaStream_t st = ix->st; u64 N = A.n; S.nseg = 0; Cn.n = 0; if (N == 0) return;
  // segments = runs of equal (query, genome)
  DBuf<u64> ukey(N, st); DBuf<u32> cnt(N + 1, st); DBuf<u32> nruns(1, st);
  { size_t tb = 0; cub::DeviceRunLengthEncode::Encode(nullptr, tb, A.hi.p, ukey.p, cnt.p, nruns.p, (int)N, st); cub::DeviceRunLengthEncode::Encode(ix->tmp.get(tb), tb, A.hi.p, ukey.p, cnt.p, nruns.p, (int)N, st); CUB_CHECK(); }
  u32 nseg = nruns.to_host()[0]; S.nseg = nseg; S.off.alloc(nseg + 1, st);
  { DBuf<u64> c64(nseg + 1, st); k_widen<<<cdiv(nseg + 1, 256), 256, 0, st>>>(cnt.p, nseg, c64.p); KERNEL_CHECK(); size_t tb = 0; cub::DeviceScan::ExclusiveSum(nullptr, tb, c64.p, S.off.p, (int)(nseg + 1), st); cub::DeviceScan::ExclusiveSum(ix->tmp.get(tb), tb, c64.p, S.off.p, (int)(nseg + 1), st); CUB_CHECK(); }
  S.h_key = ukey.to_host(nseg); S.key = std::move(ukey);
  // gap-score table on the host (gapScore lib-chaining.go:662: 0.1*g + 0.5*float32(log2(float64(g))), g integer <= max_gap)
  int gt = std::max(2, (int)std::floor(prm->max_gap) + 2); std::vector<float> gtab(gt, 0.0f);
  for (int g = 1; g < gt; g++) { volatile float a = 0.1f * (float)g; volatile float bb = 0.5f * (float)std::log2((double)g); volatile float c = a + bb; gtab[g] = c; }
  DBuf<float> dgt(gt, st); dgt.from_host(gtab.data(), gt);
  ChainParams P; P.max_gap = prm->max_gap; { volatile float w = 0.1f * (float)prm->min_single_prefix; volatile float w2 = w * (float)prm->min_single_prefix; P.min_score = w2; } P.max_distance = prm->max_distance; P.top_chains = prm->top_n_chains; P.k = ix->img.k; P.gap_score = dgt.p; P.gap_tab = gt;
  S.c_lo.alloc(N, st); S.cn.alloc(nseg, st); S.score.alloc(nseg, st); DBuf<float> sc(N, st); DBuf<u32> pred(N, st); DBuf<signed char> dirs(N, st); DBuf<u64> s2i(N, st), s2i_s(N, st); DBuf<u8> visited(N, st); visited.zero();
  k_clear_chain<<<cdiv((i64)nseg * 32, 128), 128, 0, st>>>(A.lo.p, S.off.p, nseg, P, S.c_lo.p, S.cn.p, sc.p, pred.p, dirs.p, s2i.p); KERNEL_CHECK();
  // per-segment ascending sort of (score bits << 32 | index) over the compacted prefix of each segment
  DBuf<u64> seg_end(nseg, st); k_seg_end<<<cdiv(nseg, 256), 256, 0, st>>>(S.off.p, S.cn.p, nseg, seg_end.p); KERNEL_CHECK();
  { size_t tb = 0; cub::DeviceSegmentedSort::SortKeys(nullptr, tb, s2i.p, s2i_s.p, (int)N, (int)nseg, S.off.p, seg_end.p, st); cub::DeviceSegmentedSort::SortKeys(ix->tmp.get(tb), tb, s2i.p, s2i_s.p, (int)N, (int)nseg, S.off.p, seg_end.p, st); CUB_CHECK(); }
  u32 cap = (u32)std::min<u64>(N + nseg, 0x7fffffffu); Cn.rec.alloc(cap, st); DBuf<u32> nout(1, st); nout.zero();
  k_backtrack<<<cdiv(nseg, 128), 128, 0, st>>>(S.off.p, S.cn.p, nseg, S.c_lo.p, pred.p, dirs.p, s2i.p, s2i_s.p, visited.p, P, Cn.rec.p, nout.p, cap, S.score.p); KERNEL_CHECK();
  u32 nc = nout.to_host()[0]; if (nc > cap) throw std::runtime_error("chain list overflow"); Cn.n = nc; Cn.seg_score = S.score.to_host(nseg);
  Cn.h = Cn.rec.to_host(nc);
  // drop genomes below min score (:1724), optional top-N genomes per query (:1780-1805), order chains by (segment, first TBegin, emission order) (:1967-1974)
  std::vector<char> keep(nseg, 1); for (u32 s = 0; s < nseg; s++) if (Cn.seg_score[s] < P.min_score) keep[s] = 0;
  if (prm->top_n_genomes > 0) { u32 s0 = 0; while (s0 < nseg) { u32 q = (u32)(S.h_key[s0] >> 36), s1 = s0; std::vector<u32> v; while (s1 < nseg && (u32)(S.h_key[s1] >> 36) == q) { if (keep[s1]) v.push_back(s1); s1++; }
      if ((int)v.size() > prm->top_n_genomes) { if (Cn.topn_sorted.size() <= q) Cn.topn_sorted.resize((size_t)q + 1, 0); Cn.topn_sorted[q] = 1; std::stable_sort(v.begin(), v.end(), [&](u32 a, u32 b) { return Cn.seg_score[a] > Cn.seg_score[b]; }); for (size_t t = prm->top_n_genomes; t < v.size(); t++) keep[v[t]] = 0; } s0 = s1; } }
  std::vector<ChainRec> kept; kept.reserve(nc); for (auto& r : Cn.h) if (keep[r.seg]) { r.score = Cn.seg_score[r.seg]; kept.push_back(r); }
  bucket_sort(kept, nseg, [](const ChainRec& a) { return a.seg; }, [](const ChainRec& a, const ChainRec& b) { if (a.t0 != b.t0) return a.t0 < b.t0; return a.ord < b.ord; });
  Cn.h.swap(kept); Cn.n = (u32)Cn.h.size();
}

// =====================================================================================================
// K4: pseudo-alignment (SeqComparator.Index/Compare lib-seq_compare.go:115-159,:335-522; tree.Search tree/tree.go:441-527;
//     ClearSubstrPairs; TrimSubStrPairs :553-621; Chainer2 lib-chaining2.go:152-658)
// =====================================================================================================
__global__ void k_tree_flags(const u32* __restrict__ vals, u64 n, u32* __restrict__ flags) { u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; if (i < n) flags[i] = (vals[i] >> 31) ? 0u : 1u; }
__global__ void k_tree_scatter(const u64* __restrict__ keys, const u32* __restrict__ vals, const u32* __restrict__ pos, u64 n, u64* __restrict__ tkeys, u32* __restrict__ tvals) {
  u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x; if (i < n && !(vals[i] >> 31)) { tkeys[pos[i]] = keys[i]; tvals[pos[i]] = vals[i]; } }
__global__ void k_tree_offsets(const u32* __restrict__ pos, const u64* __restrict__ koff, int nq, u64 total_k, u32 total_t, u32* __restrict__ toff) {
  int q = blockIdx.x * blockDim.x + threadIdx.x; if (q <= nq) toff[q] = (q == nq || koff[q] >= total_k) ? total_t : pos[koff[q]]; }

struct WinItem { u32 q, g; i32 tBegin, tEnd, W, qBegin, qEnd; u32 rc; i32 mp; u32 chain; };  // mp = min prefix for this window (lib-seq_compare.go:339-348)

__device__ __forceinline__ u32 win_base(const u8* __restrict__ g2, i32 tBegin, i32 tEnd, u32 rc, i32 i) { return rc ? 3u - get_base(g2, (u64)(tEnd - i)) : get_base(g2, (u64)(tBegin + i)); }
// 16 window bases [16x, 16x+16) of the oriented window as one word (first base in the top bits), zero beyond the window: five byte loads and a shift instead of
// sixteen single-base extractions; the minus strand reverses the 2-bit groups of the mirrored genome word and complements them
__device__ __forceinline__ u32 gword16(const u8* __restrict__ g2, i64 pos) {   // genome bases [pos, pos+16), pos >= 0; reads 5 bytes (the payload arrays are padded)
  const u8* p = g2 + (pos >> 2); const u64 v = ((u64)p[0] << 32) | ((u64)p[1] << 24) | ((u64)p[2] << 16) | ((u64)p[3] << 8) | (u64)p[4]; return (u32)(v >> (8 - 2 * (pos & 3))); }
__device__ __forceinline__ u32 win_word16(const u8* __restrict__ g2, i32 tBegin, i32 tEnd, u32 rc, i32 W, i32 x) {
  const i32 i0 = x * 16; if (i0 >= W) return 0u; const i32 n = min(16, W - i0); u32 v;
  if (!rc) v = gword16(g2, (i64)tBegin + i0);
  else { const i64 hi = (i64)tEnd - i0, lo = hi - 15;   // window base i0 + j = complement of genome base hi - j
    u32 g; if (lo >= 0) g = gword16(g2, lo); else g = gword16(g2, 0) >> (2 * (int)(-lo));   // genome bases [lo, hi] right-aligned; bases before the genome start read as 0 and fall outside n anyway
    u32 r = __brev(g); r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1); v = ~r; }
  return n == 16 ? v : (v & ~((1u << (2 * (16 - n))) - 1u));
}
__device__ __forceinline__ int lcp31(u64 a, u64 b) { return a == b ? 31 : min(31, (__clzll(a ^ b) >> 1) - 1); }   // K = 31: LZ/2 + K - 32

// tree.Search emulated on the sorted table, incl. the uint8 wrap in the mismatch branch (tree/tree.go:498-501)
__device__ bool tree_search_slow(const u64* __restrict__ a, u32 n, u64 key, int p, u32* rlo, u32* rhi) {
  const int K = 31; u32 lo = 0, hi = n; int depth = 0;
  while (depth < K) {
    int sh = 2 * (K - depth - 1); u64 qb = (key >> sh) & 3; u32 x = lo, y = hi;
    while (x < y) { u32 m = (x + y) >> 1; if (((a[m] >> sh) & 3) < qb) x = m + 1; else y = m; } u32 l = x; y = hi;
    while (x < y) { u32 m = (x + y) >> 1; if (((a[m] >> sh) & 3) <= qb) x = m + 1; else y = m; } u32 h = x;
    if (l == h) return false;
    int nodeEnd = lcp31(a[l], a[h - 1]), nk = nodeEnd - depth, mm = lcp31(key, a[l]);
    if (mm >= nodeEnd) { depth = nodeEnd; lo = l; hi = h; if (depth >= p) { *rlo = lo; *rhi = hi; return true; } continue; }
    int atleast = p - depth; bool hit; if (atleast <= nk) hit = mm >= depth + atleast; else { u64 nxt = (key >> (2 * (K - depth - atleast))) & ((1ull << (2 * atleast)) - 1); hit = (nxt == 0); }
    if (hit) { *rlo = l; *rhi = h; return true; } return false;
  }
  return false;
}
// keys sharing >= p leading bases with `key`: [lo,hi) in the sorted table (fast path = range search; the radix-tree quirk only
// adds results when the range is empty and bases [p-2,p) of the key are AA)
// The radix-tree quirk can only fire when the query shares d+1 >= 1 bases with some key (d = matched depth), d <= p-2, and its bases
// [d, p) are all A. With L = longest common prefix with the two neighbours of the insertion point (L < p here), that implies bases
// [L-1, p) all A: a cheap necessary test that keeps the emulation off the common path.
__device__ __forceinline__ bool quirk_possible(const u64* __restrict__ a, u32 n, u32 x, u64 key, int p) {
  const int K = 31; if (p < 2 || ((key >> (2 * (K - p))) & 0xF) != 0) return false; int L = 0; if (x > 0) L = max(L, lcp31(key, a[x - 1])); if (x < n) L = max(L, lcp31(key, a[x])); if (L < 1) return false;
  int from = min(L - 1, p - 2); u64 span = (key >> (2 * (K - p))) & ((1ull << (2 * (p - from))) - 1); return span == 0;
}
__device__ __forceinline__ bool tree_search_in(const u64* __restrict__ a, u32 n, u32 x, u32 y, u64 key, int p, u32* rlo, u32* rhi) {   // [x,y) must contain every key sharing >= p bases with `key`
  const int K = 31; u64 low = (1ull << (2 * (K - p))) - 1, left = key & ~low, right = key | low;
  while (x < y) { u32 m = (x + y) >> 1; if (a[m] < left) x = m + 1; else y = m; }
  u32 e = x; while (e < n && a[e] <= right) e++;
  if (e > x) { *rlo = x; *rhi = e; return true; }
  if (quirk_possible(a, n, x, key, p)) return tree_search_slow(a, n, key, p, rlo, rhi);
  return false;
}
__device__ __forceinline__ bool tree_search(const u64* __restrict__ a, u32 n, u64 key, int p, u32* rlo, u32* rhi) { return tree_search_in(a, n, 0, n, key, p, rlo, rhi); }


// ---- K4 v2: per-query hash index over the 11-base prefixes of the table (the default minimum prefix), window staged in shared memory
// slot = prefix22 << 42 | start24 << 18 | count18 ; empty = ~0
#define TH_EMPTY 0xFFFFFFFFFFFFFFFFull
__device__ __forceinline__ u32 th_hash(u32 prefix, u32 mask) { return (prefix * 2654435761u >> 7) & mask; }
__global__ void k_tree_hash_build(const u64* __restrict__ tkeys, const u32* __restrict__ toff, const u64* __restrict__ hoff, int nq, u64* __restrict__ table) {
  int q = blockIdx.y; if (q >= nq) return; u32 t0 = toff[q], n = toff[q + 1] - t0; u64 h0 = hoff[q]; u32 H = (u32)(hoff[q + 1] - h0); if (!n || !H) return; const u64* a = tkeys + t0; u64* T = table + h0;
  for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { u32 p = (u32)(a[i] >> 40); if (i > 0 && (u32)(a[i - 1] >> 40) == p) continue; u32 c = 1; while (i + c < n && (u32)(a[i + c] >> 40) == p) c++;
    if (c >= (1u << 18) || i >= (1u << 24)) c = 0;   // does not fit the slot: count 0 = "use binary search"
    u64 slot = ((u64)p << 42) | ((u64)(i & 0xFFFFFF) << 18) | c; u32 h = th_hash(p, H - 1); for (;;) { unsigned long long old = atomicCAS((unsigned long long*)&T[h], TH_EMPTY, slot); if (old == TH_EMPTY) break; h = (h + 1) & (H - 1); } }
}
__device__ __forceinline__ bool th_lookup(const u64* __restrict__ T, u32 H, u32 prefix, u32* start, u32* count) {   // false = prefix absent
  u32 h = th_hash(prefix, H - 1); for (;;) { u64 s = T[h]; if (s == TH_EMPTY) return false; if ((u32)(s >> 42) == prefix) { *start = (u32)(s >> 18) & 0xFFFFFF; *count = (u32)s & 0x3FFFF; return true; } h = (h + 1) & (H - 1); }
}
// keys sharing >= p bases: hash path when p == 11, else binary search; identical result set to tree_search()
__device__ __forceinline__ bool tree_search_h(const u64* __restrict__ a, u32 n, const u64* __restrict__ T, u32 H, u64 key, int p, u32* rlo, u32* rhi) {
  if (p == 11 && H) { u32 st, c; if (th_lookup(T, H, (u32)(key >> 40), &st, &c)) { if (c) { *rlo = st; *rhi = st + c; return true; } return tree_search(a, n, key, p, rlo, rhi); }
    if (((key >> 40) & 0xF) == 0) return tree_search(a, n, key, p, rlo, rhi); return false; }
  return tree_search(a, n, key, p, rlo, rhi);
}
// one CTA per window. The oriented window is packed into shared memory (16 bases per word) once; every thread extracts its 31-mer
// with three word loads. Anchors go to a per-window region of capacity cap[it]; counts[it] is exact even when it overflows.
__global__ void __launch_bounds__(128) k_pa_anchors2(const WinItem* __restrict__ items, const u32* __restrict__ item_ids, u32 nitems, const u8* __restrict__ g2bit, const u64* __restrict__ g_off, const u64* __restrict__ tkeys, const u32* __restrict__ tvals, const u32* __restrict__ toff,
                                                     const u64* __restrict__ htab, const u64* __restrict__ hoff, const u64* __restrict__ abeg, const u32* __restrict__ acap, u32* __restrict__ counts, u64* __restrict__ a_lo, u64 dbg_total) {
  __shared__ u32 s_base; extern __shared__ u32 sw[];
  u32 bi = blockIdx.x; if (bi >= nitems) return; u32 it = item_ids ? item_ids[bi] : bi; WinItem w = items[it]; const int K = 31; const u8* g2 = g2bit + g_off[w.g]; const u64* tk = tkeys + toff[w.q]; const u32* tv = tvals + toff[w.q]; u32 tn = toff[w.q + 1] - toff[w.q];
  const u64* HT = htab + hoff[w.q]; u32 H = (u32)(hoff[w.q + 1] - hoff[w.q]); const u64 base0 = abeg[it]; const u32 cap = acap[it];
  i32 nw = (w.W + 15) / 16 + 2;
  for (i32 x = threadIdx.x; x < nw; x += 128) { u32 v = 0; for (int j = 0; j < 16; j++) { i32 i = x * 16 + j; u32 b = (i < w.W) ? win_base(g2, w.tBegin, w.tEnd, w.rc, i) : 0; v = (v << 2) | b; } sw[x] = v; }
  if (threadIdx.x == 0) s_base = 0; __syncthreads();
  i32 np = w.W - K + 1; const u64 ccc = 0x1555555555555555ull, ggg = 0x2AAAAAAAAAAAAAAAull, ttt = 0x3FFFFFFFFFFFFFFFull; const u32 begin = (u32)w.qBegin, end = (u32)w.qEnd;
  for (i32 t0 = 0; t0 < np; t0 += 128) {
    i32 idx = t0 + (i32)threadIdx.x; u32 c = 0; u64 km = 0, kr = 0; bool ok = false; u32 l1 = 0, h1 = 0, l2 = 0, h2 = 0; bool f1 = false, f2 = false;
    if (idx < np) { u32 wi = (u32)idx >> 4, sh = ((u32)idx & 15) * 2; u64 hi64 = ((u64)sw[wi] << 32) | sw[wi + 1]; u64 v = sh ? ((hi64 << sh) | ((u64)sw[wi + 2] >> (32 - sh))) : hi64; km = v >> 2;
      kr = kmer_reverse62(~km & ttt, K); ok = !(km == 0 || km == ccc || km == ggg || km == ttt); }
    if (ok && tn) {
      f1 = tree_search_h(tk, tn, HT, H, km, w.mp, &l1, &h1);
      if (f1) for (u32 u = l1; u < h1; u++) { u32 v = tv[u]; int lp = lcp31(km, tk[u]); u32 p = v >> 1; if ((v & 1) == 1 || p < begin || p + (u32)lp > end) continue; c++; }
      f2 = tree_search_h(tk, tn, HT, H, kr, w.mp, &l2, &h2);
      if (f2) for (u32 u = l2; u < h2; u++) { u32 v = tv[u]; int lp = lcp31(kr, tk[u]); u32 p = (v >> 1) + (u32)K - (u32)lp; if ((v & 1) == 0 || p + (u32)lp < begin || p > end) continue; c++; }
    }
    u32 wpos = 0; if (c) wpos = atomicAdd(&s_base, c);   // order inside a window is irrelevant (sorted next); s_base ends as the exact count
    if (c && (u64)wpos + c <= (u64)cap) { u64* out = a_lo + base0 + wpos;
      {
      if (f1) for (u32 u = l1; u < h1; u++) { u32 v = tv[u]; int lp = lcp31(km, tk[u]); u32 p = v >> 1; if ((v & 1) == 1 || p < begin || p + (u32)lp > end) continue; *out++ = pack_lo((i32)p, (u32)lp, idx, 0, 0); }
      if (f2) for (u32 u = l2; u < h2; u++) { u32 v = tv[u]; int lp = lcp31(kr, tk[u]); u32 p = (v >> 1) + (u32)K - (u32)lp; if ((v & 1) == 0 || p + (u32)lp < begin || p > end) continue; *out++ = pack_lo((i32)p, (u32)lp, idx + K - lp, 1, 1); } } }
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[it] = s_base;
}

// helpers of k_pa_anchors3 (plain functions, no early returns around the barriers of the caller)
__device__ __forceinline__ void pa3_push(u32* qu, u32* cnt, bool have, u32 e, int lane) {   // warp-aggregated append to a shared-memory queue
  u32 bal = __ballot_sync(FULLMASK, have); u32 base = 0; int ldr = bal ? (__ffs(bal) - 1) : 0; if (bal && lane == ldr) base = atomicAdd(cnt, __popc(bal)); base = __shfl_sync(FULLMASK, base, ldr);
  if (have) qu[base + __popc(bal & ((1u << lane) - 1))] = e;
}
__device__ __forceinline__ u64 pa3_key(const u32* sw, u32 e, u64 ttt) {   // k-mer (strand e&1) at window position e>>1
  u32 idx = e >> 1; u32 wi = idx >> 4, sh = (idx & 15) * 2; u64 hi64 = ((u64)sw[wi] << 32) | sw[wi + 1]; u64 v = sh ? ((hi64 << sh) | ((u64)sw[wi + 2] >> (32 - sh))) : hi64; u64 km = v >> 2; return (e & 1) ? kmer_reverse62(~km & ttt, 31) : km;
}
// anchors of one candidate: table rows [l,h) share >= mp bases with `key`; keep those inside the query region, append to the window's slot range
__device__ __forceinline__ void pa3_emit(const u64* sk, const u32* sv, u32 e, u64 key, u32 l, u32 h, u32 begin, u32 end, u32* s_base, u32 cap, u64* out_base) {
  const int K = 31; const u32 idx = e >> 1; const bool rcs = e & 1; u32 c = 0;
  for (u32 u = l; u < h; u++) { u32 vv = sv[u]; int lp = lcp31(key, sk[u]); bool ok; if (!rcs) { u32 p = vv >> 1; ok = !((vv & 1) == 1 || p < begin || p + (u32)lp > end); } else { u32 p = (vv >> 1) + (u32)K - (u32)lp; ok = !((vv & 1) == 0 || p + (u32)lp < begin || p > end); } c += ok ? 1u : 0u; }
  if (c) { u32 wpos = atomicAdd(s_base, c);
    if ((u64)wpos + c <= (u64)cap) { u64* out = out_base + wpos;
      for (u32 u = l; u < h; u++) { u32 vv = sv[u]; int lp = lcp31(key, sk[u]);
        if (!rcs) { u32 p = vv >> 1; if (!((vv & 1) == 1 || p < begin || p + (u32)lp > end)) *out++ = pack_lo((i32)p, (u32)lp, (i32)idx, 0, 0); }
        else { u32 p = (vv >> 1) + (u32)K - (u32)lp; if (!((vv & 1) == 0 || p + (u32)lp < begin || p > end)) *out++ = pack_lo((i32)p, (u32)lp, (i32)idx + K - lp, 1, 1); } } } }
}

// ---- K4 v3: one CTA per QUERY. The query's sorted (k-mer, loc) table is staged in shared memory once and reused by all of the query's
// target windows (typically one per candidate genome), so every table probe is a shared-memory binary search instead of an L2 round trip.
// Windows are packed into shared memory one at a time. Queries whose table does not fit use k_pa_anchors2 (table in L2).
// Named-barrier helpers: a CTA of NG independent groups, each working on its own window (barrier ids 1..NG; id 0 = __syncthreads).
__device__ __forceinline__ void gbar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ bool gbar_or(int id, int n, bool p) { u32 r; asm volatile("{ .reg .pred q, qr; setp.ne.u32 q, %3, 0; bar.red.or.pred qr, %1, %2, q; selp.u32 %0, 1, 0, qr; }" : "=r"(r) : "r"(id), "r"(n), "r"((u32)p) : "memory"); return r != 0; }
template <int NT>   // threads per CTA: 256, or 1024 when the table leaves room for a single CTA per SM anyway (5-kb queries). The 1,024-thread CTA runs as TWO groups
                    // of 512 threads, each on its own window with its own named barrier: with one CTA per SM a whole-CTA barrier left nothing to issue (ncu: 24 warps
                    // waiting at the barrier per issue slot, issue-active 23 %); now one group scans while the other waits.
__global__ void __launch_bounds__(NT) k_pa_anchors3(const WinItem* __restrict__ items, const u32* __restrict__ qlist, const u32* __restrict__ qitem_beg, const u32* __restrict__ qitem_end, const u8* __restrict__ g2bit, const u64* __restrict__ g_off,
                                                     const u64* __restrict__ tkeys, const u32* __restrict__ tvals, const u32* __restrict__ toff, const u64* __restrict__ abeg, const u32* __restrict__ acap, u32* __restrict__ counts, u64* __restrict__ a_lo, u32 max_tn, u32 win_words) {
  constexpr int BW = NT >= 1024 ? 4096 : 1024, BSH = NT >= 1024 ? 15 : 17;   // Bloom filter over hashed 11-base prefixes: 32 kbit for tables of up to ~4,000 rows, 128 kbit for the big-table variant (a 10,000-row table filled 26 % of 32 kbit)
  constexpr int NG = NT >= 1024 ? 2 : 1, GT = NT / NG;   // groups per CTA, threads per group
  __shared__ u32 s_base_g[NG]; __shared__ u32 bloom[BW]; __shared__ u32 pdir[257];   // pdir: first table row of every 4-base prefix
  extern __shared__ __align__(16) u8 smem3[]; u64* sk = (u64*)smem3; u32* sv = (u32*)(sk + max_tn); const int grp = threadIdx.x / GT, gtid = threadIdx.x % GT, bid = NG > 1 ? 1 + grp : 0; u32* sw = sv + max_tn + (size_t)grp * win_words; u32* const s_base = &s_base_g[grp];
  const u32 q = qlist[blockIdx.x]; const u32 t0q = toff[q], tn = toff[q + 1] - t0q; const int K = 31;
  for (u32 i = threadIdx.x; i < (u32)BW; i += NT) bloom[i] = 0; for (u32 i = threadIdx.x; i < 257; i += NT) pdir[i] = tn;
  for (u32 i = threadIdx.x; i < tn; i += NT) { sk[i] = tkeys[t0q + i]; sv[i] = tvals[t0q + i]; }
  __syncthreads();
  for (u32 i = threadIdx.x; i < tn; i += NT) { u64 kk = sk[i]; u32 h = ((u32)(kk >> 40) * 2654435761u) >> BSH; atomicOr(&bloom[h >> 5], 1u << (h & 31)); u32 b = (u32)(kk >> 54); if (i == 0 || (u32)(sk[i - 1] >> 54) != b) pdir[b] = i; }
  __syncthreads();
  if (threadIdx.x == 0) { u32 nxt = tn; for (int b = 255; b >= 0; b--) { if (pdir[b] == tn) pdir[b] = nxt; else nxt = pdir[b]; } pdir[256] = tn; }   // empty buckets -> start of the next one
  const u64 ccc = 0x1555555555555555ull, ggg = 0x2AAAAAAAAAAAAAAAull, ttt = 0x3FFFFFFFFFFFFFFFull; const int lane = threadIdx.x & 31;
  // Work compaction: most window positions fail the Bloom / prefix pre-check, and of those that search the table only a few take the
  // radix-tree emulation. Doing everything in one pass ran at ~10 active lanes per instruction; instead positions that pass the pre-check
  // are queued (position << 1 | strand) and searched 256 at a time, and the rare slow-path searches are queued again.
  __shared__ u32 qfast_g[NG][4 * GT], qslow_g[NG][3 * GT]; __shared__ u32 nslow_g[NG]; __shared__ u32 s_wc_g[NG][2][32]; u32* const qfast = qfast_g[grp]; u32* const qslow = qslow_g[grp]; u32* const nslow = &nslow_g[grp]; u32 (*s_wc)[32] = s_wc_g[grp];
  __syncthreads();   // table, Bloom filter and directory complete before any group starts
  for (u32 it = qitem_beg[q] + grp; it < qitem_end[q]; it += NG) {
    WinItem w = items[it]; const u8* g2 = g2bit + g_off[w.g]; const u64 base0 = abeg[it]; const u32 cap = acap[it]; const u32 begin = (u32)w.qBegin, end = (u32)w.qEnd;
    gbar(bid, GT);   // previous window of this group fully consumed
    i32 nw = (w.W + 15) / 16 + 2;
    for (i32 x = gtid; x < nw; x += GT) sw[x] = win_word16(g2, w.tBegin, w.tEnd, w.rc, w.W, x);
    if (gtid == 0) { *s_base = 0; *nslow = 0; } gbar(bid, GT);
    const i32 np = w.W - K + 1; u64* const out_base = a_lo + base0;
    // The scan queues candidates, the drain searches them 1 CTA-width at a time. Barrier stalls were 24 waiting warps per issue slot in the 1,024-thread
    // variant (ncu), so the queue length lives in a REGISTER that every thread advances identically (nq; the per-slice total comes from the same
    // shared-memory warp counts in every warp), the warp counts are double-buffered, and the slow-path queue is only looked at when some thread
    // pushed to it (__syncthreads_or): a slice costs one barrier, a drain two.
    u32 nq = 0; int par = 0; bool any_slow_window = false;
    auto slow_drain = [&](u32 ns0) { for (u32 ns = ns0; ns > 0;) { u32 tk = min(ns, (u32)GT), st2 = ns - tk; if ((u32)gtid < tk) { u32 e2 = qslow[st2 + gtid]; u64 key = pa3_key(sw, e2, ttt); u32 l = 0, h = 0; if (tree_search_slow(sk, tn, key, w.mp, &l, &h)) pa3_emit(sk, sv, e2, key, l, h, begin, end, s_base, cap, out_base); } ns = st2; } };
    for (i32 base = 0;; base += GT) {
      const bool more = (base < np) && tn;   // uniform: another slice of window positions to scan
      if (more) { i32 idx = base + (i32)gtid; bool c1 = false, c2 = false;
        if (idx < np) { u32 wi = (u32)idx >> 4, sh = ((u32)idx & 15) * 2; u64 hi64 = ((u64)sw[wi] << 32) | sw[wi + 1]; u64 v = sh ? ((hi64 << sh) | ((u64)sw[wi + 2] >> (32 - sh))) : hi64; u64 km = v >> 2;
          if (!(km == 0 || km == ccc || km == ggg || km == ttt)) { u64 kr = kmer_reverse62(~km & ttt, K);
            u32 hb = ((u32)(km >> 40) * 2654435761u) >> BSH; c1 = (w.mp != 11) || ((bloom[hb >> 5] >> (hb & 31)) & 1) || (((km >> 40) & 0xF) == 0);
            hb = ((u32)(kr >> 40) * 2654435761u) >> BSH; c2 = (w.mp != 11) || ((bloom[hb >> 5] >> (hb & 31)) & 1) || (((kr >> 40) & 0xF) == 0); } }
        // queue the candidates of this slice with a block-wide scan (ballots -> per-warp counts in shared memory -> prefix over the warps)
        const u32 b1 = __ballot_sync(FULLMASK, c1), b2 = __ballot_sync(FULLMASK, c2); if (lane == 0) s_wc[par][gtid >> 5] = __popc(b1) + __popc(b2); gbar(bid, GT);
        const u32 wv = (lane < GT / 32) ? s_wc[par][lane] : 0u; const u32 before = __reduce_add_sync(FULLMASK, lane < (int)(gtid >> 5) ? wv : 0u), tot = __reduce_add_sync(FULLMASK, wv); const u32 wbase = nq + before;
        if (c1) qfast[wbase + __popc(b1 & ((1u << lane) - 1))] = (u32)idx << 1; if (c2) qfast[wbase + __popc(b1) + __popc(b2 & ((1u << lane) - 1))] = ((u32)idx << 1) | 1u; nq += tot; par ^= 1; }
      // drain: full chunks while scanning, everything at the end
      while (more ? (nq >= (u32)GT) : (nq > 0)) {
        gbar(bid, GT);   // the queue entries written above (by other warps) are visible
        const u32 take = min(nq, (u32)GT), start = nq - take; bool slow = false; u32 e = 0;
        if ((u32)gtid < take) { e = qfast[start + gtid]; u64 key = pa3_key(sw, e, ttt); const int p = w.mp; u64 low = (1ull << (2 * (K - p))) - 1, left = key & ~low, right = key | low; u32 b = (u32)(key >> 54), x = pdir[b], y = pdir[b + 1];
          while (x < y) { u32 m = (x + y) >> 1; if (sk[m] < left) x = m + 1; else y = m; } u32 en = x; while (en < tn && sk[en] <= right) en++;
          if (en > x) pa3_emit(sk, sv, e, key, x, en, begin, end, s_base, cap, out_base); else slow = quirk_possible(sk, tn, x, key, p); }
        nq = start;
        if (gbar_or(bid, GT, slow)) {   // (also the barrier between reading this chunk and the next slice overwriting it) some radix-tree emulations to queue: rare
          any_slow_window = true; pa3_push(qslow, nslow, slow, e, lane); gbar(bid, GT); const u32 ns1 = *nslow; gbar(bid, GT);
          if (ns1 >= (u32)GT) { slow_drain(ns1); gbar(bid, GT); if (gtid == 0) *nslow = 0; gbar(bid, GT); } }
      }
      if (!more) break;
    }
    if (any_slow_window) { gbar(bid, GT); const u32 ns0 = *nslow; gbar(bid, GT); if (ns0 > 0) slow_drain(ns0); }
    gbar(bid, GT); if (gtid == 0) counts[it] = *s_base;
  }
}

struct C2Rec { u32 item, ord; i32 qb, qe, tb, te, aligned_q, aligned_t, matched, n_anchors; };
struct Chain2Params { int max_gap, min_score, min_align_len, band_count, band_base, k; };

// Per-window anchor sort (the order ClearSubstrPairs / Chainer2 need: ascending packed key). Windows are binned by anchor count on the host;
// each bin runs a shared-memory bitonic network sized to the bin (TPW threads per window), only the rare huge window goes through CUB.
template <int CAP, int TPW>
__global__ void __launch_bounds__(TPW < 128 ? 128 : TPW) k_pa_sort(const u32* __restrict__ list, u32 nlist, const u64* __restrict__ abeg, const u64* __restrict__ cbeg, const u32* __restrict__ counts, const u64* __restrict__ in, u64* __restrict__ out) {
  extern __shared__ u64 ssort[];
  const u32 wpc = blockDim.x / TPW, sub = threadIdx.x / TPW, t0 = threadIdx.x % TPW; u32 li = blockIdx.x * wpc + sub; bool live = li < nlist;
  u64* sm = ssort + (size_t)sub * CAP; u32 it = live ? list[li] : 0; u32 n = live ? counts[it] : 0; u64 b = live ? abeg[it] : 0, cb = live ? cbeg[it] : 0; u32 Pn = 2; while (Pn < n) Pn <<= 1;
  auto sync = [&]() { if (TPW == 32) __syncwarp(); else __syncthreads(); };
  for (u32 i = t0; i < Pn; i += TPW) sm[i] = (i < n) ? in[b + i] : ~0ull; sync();
  for (u32 kk = 2; kk <= Pn; kk <<= 1) for (u32 j = kk >> 1; j > 0; j >>= 1) { for (u32 t = t0; t < (Pn >> 1); t += TPW) { u32 i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), l = i | j; bool up = (i & kk) == 0; u64 x = sm[i], y = sm[l]; if ((x > y) == up) { sm[i] = y; sm[l] = x; } } sync(); }
  for (u32 i = t0; i < n; i += TPW) out[cb + i] = sm[i];
}
// windows too large for the shared-memory sort: copy their anchors to the compact layout, CUB sorts them there
__global__ void k_pa_gather(const u32* __restrict__ list, u32 nlist, const u64* __restrict__ abeg, const u64* __restrict__ cbeg, const u32* __restrict__ counts, const u64* __restrict__ in, u64* __restrict__ out) {
  if (blockIdx.x >= nlist) return; u32 it = list[blockIdx.x]; u32 n = counts[it]; u64 b = abeg[it], cb = cbeg[it]; for (u32 i = threadIdx.x; i < n; i += blockDim.x) out[cb + i] = in[b + i];
}

// one warp per window: nested-anchor removal, trimming, banded chaining DP, region splitting. Scalar control flow is executed
// redundantly by all lanes (uniform); the DP inner loop and arg-max scans are lane-parallel. lo_sorted holds each window's anchors sorted;
// c_lo (the unsorted copy) is reused as the compacted output.
__global__ void __launch_bounds__(128) k_pa_chain(const u64* __restrict__ lo_sorted, const u64* __restrict__ abeg, const u64* __restrict__ aend, const u64* __restrict__ cbeg, u32 nitems, Chain2Params P, u64* __restrict__ c_lo, i32* __restrict__ score, u32* __restrict__ pred, u64* __restrict__ stack,
                                                  C2Rec* __restrict__ out, u32* __restrict__ nout, u32 cap) {
  u32 it = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; int lane = threadIdx.x & 31; if (it >= nitems) return;
  u64 b = abeg[it]; u32 n = (u32)(aend[it] - b); if (n == 0) return; u64* C = c_lo + b; b = cbeg[it]; const u64* A = lo_sorted + b; const int k = P.k;   // from here on b indexes the compact arrays
  // ---- ClearSubstrPairs
  if (n > 1) { u32 kept = 0;
    for (u32 base = 0; base < n; base += 32) { u32 i = base + lane; bool keep = false;
      if (i < n) { keep = true; if (i > 0) { u64 v = A[i]; i32 vq = a_q(v), vl = a_len(v), vt = a_t(v); i32 vQEnd = vq + vl, up = max(vQEnd - k, 0), vTEnd = vt + vl;
          for (i32 j = (i32)i - 1; j >= 0; j--) { u64 p = A[j]; i32 pq = a_q(p); if (pq < up) break; i32 pl = a_len(p), pt = a_t(p); if (vQEnd <= pq + pl && vt >= pt && vTEnd <= pt + pl) { keep = false; break; } } } }
      u32 bal = __ballot_sync(FULLMASK, keep); if (keep) C[kept + __popc(bal & ((1u << lane) - 1))] = A[i]; kept += __popc(bal); }
    n = kept; } else { if (lane == 0) C[0] = A[0]; }
  __syncwarp();
  // ---- TrimSubStrPairs(subs, k, 100)
  if (n >= 2) { i32 last = (i32)n - 1; u64 _p = C[0]; i32 start = 0;
    for (i32 i = 0; i < last; i++) { u64 p = C[i + 1]; i32 pq = a_q(p), pt = a_t(p), _q = a_q(_p), _t = a_t(_p), _l = a_len(_p); i32 dist = max(abs(pq - _q), abs(pt - _t)); bool c1 = (pq == _q || pt == _t);
      bool take = false; if ((float)dist < 100.0f) { if (c1) take = true; else { i32 g2 = abs(abs(_q - pq) - abs(_t - pt)); i32 qo = 0, to = 0; if (pq >= _q && pq <= _q + _l) qo = _q + _l - pq + 1; if (pt >= _t && pt <= _t + _l) to = _t + _l - pt + 1;   // overlap(_p, p)
          take = (g2 > 11 && (double)max(qo, to) / (double)_l > 0.8); } }
      if (take) { start = i; _p = p; continue; } break; }
    _p = C[last]; i32 end = last;
    for (i32 i = (i32)n - 2; i >= 0; i--) { u64 p = C[i]; i32 pq = a_q(p), pt = a_t(p), pl = a_len(p), _q = a_q(_p), _t = a_t(_p), _l = a_len(_p); i32 dist = max(abs(pq - _q), abs(pt - _t)); bool c1 = (pq == _q || pt == _t);
      bool take = false; if ((float)dist < 100.0f) { if (c1) take = true; else { i32 g2 = abs(abs(pq - _q) - abs(pt - _t)); i32 qo = 0, to = 0; if (_q >= pq && _q <= pq + pl) qo = pq + pl - _q + 1; if (_t >= pt && _t <= pt + pl) to = pt + pl - _t + 1;   // overlap(p, _p)
          take = (g2 > 11 && (double)max(qo, to) / (double)_l > 0.8); } }
      if (take) { end = i; _p = p; continue; } break; }
    if (start >= end) return;   // all discarded
    C += start; n = (u32)(end - start + 1); b += start; }
  i32* S = score + b; u32* Pd = pred + b; u64* ST = stack + b; u32 ord = 0;
  auto emit = [&](i32 qb, i32 qe, i32 tb, i32 te, i32 aq, i32 at, i32 matched, i32 na) { if (lane == 0) { u32 wv = atomicAdd(nout, 1u); if (wv < cap) { C2Rec r; r.item = it; r.ord = ord; r.qb = qb; r.qe = qe; r.tb = tb; r.te = te; r.aligned_q = aq; r.aligned_t = at; r.matched = matched; r.n_anchors = na; out[wv] = r; } } ord++; };
  // ---- Chainer2.Chain
  if (n == 1) { u64 v = C[0]; i32 sl = a_len(v); if (sl >= P.min_score && sl >= P.min_align_len) emit(a_q(v), a_q(v) + sl - 1, a_t(v), a_t(v) + sl - 1, sl, 0, sl, 1); return; }
  if (lane == 0) { S[0] = a_len(C[0]); Pd[0] = 0; } __syncwarp();
  i32 M = 0; u32 Mi = 0;
  for (u32 i = 1; i < n; i++) {
    u64 av = C[i]; i32 aq = a_q(av), al = a_len(av), at = a_t(av); i32 bs = -1, bj = -1; i32 cnt = 0; bool stop = false;
    for (i32 top = (i32)i - 1; top >= 0 && !stop; top -= 32) {
      i32 j = top - lane; bool nonskip = false; i32 bq = 0, bl = 0, bt = 0;
      if (j >= 0) { u64 bv = C[j]; bq = a_q(bv); bl = a_len(bv); bt = a_t(bv); nonskip = !(bq == aq || bt > at); }
      u32 bal = __ballot_sync(FULLMASK, nonskip); i32 mycnt = cnt + __popc(bal & ((2u << lane) - 1));   // inclusive count in descending-j order
      bool brk = nonskip && !((aq - bq - bl) <= P.band_base || mycnt <= P.band_count);
      u32 bb = __ballot_sync(FULLMASK, brk); int firstBrk = bb ? (__ffs(bb) - 1) : 32; if (bb) stop = true;
      i32 s = -1; bool valid = nonskip && lane < firstBrk;
      if (valid) { i32 g = abs(abs(aq - bq) - abs(at - bt)); if (g > P.max_gap) valid = false; else s = S[j] + bl - g; }
      i32 rs = valid ? s : INT32_MIN, rj = valid ? j : INT32_MAX;
      { i32 mx = __reduce_max_sync(FULLMASK, rs); rj = __reduce_min_sync(FULLMASK, rs == mx ? rj : INT32_MAX); rs = mx; }   // best score of the chunk, smallest j among equals (hardware warp reductions)
      if (rj != INT32_MAX && rs >= bs) { bs = rs; bj = rj; }     // s >= m: later (smaller j) wins ties
      cnt += __popc(bal);
    }
    i32 m = al; u32 mj = i; if (bj >= 0 && bs >= al) { m = bs; mj = (u32)bj; }
    if (lane == 0) { S[i] = m; Pd[i] = mj; } if (m > M) { M = m; Mi = i; }
    __syncwarp();
  }
  if (M < P.min_score) return;
  // ---- chainARegion, recursion -> explicit stack (right region first, then left)
  i32 sp = 0; if (lane == 0) ST[0] = ((u64)0 << 32) | n; __syncwarp(); sp = 1; bool top_level = true;
  while (sp > 0) {
    sp--; u64 e = ST[sp]; i32 rlo = (i32)(e >> 32), rhi = (i32)(e & 0xffffffffu); i32 len = rhi - rlo; i32 mi;
    if (top_level) { mi = (i32)Mi; top_level = false; }
    else { i32 bm = 0, bi = INT32_MAX; for (i32 x = rlo + lane; x < rhi; x += 32) { i32 v = S[x]; if (v > bm) { bm = v; bi = x; } }
      for (int o = 16; o; o >>= 1) { i32 om = __shfl_xor_sync(FULLMASK, bm, o), oi = __shfl_xor_sync(FULLMASK, bi, o); if (om > bm || (om == bm && oi < bi)) { bm = om; bi = oi; } }
      if (bm < P.min_score || bi == INT32_MAX) continue; mi = bi - rlo; }
    i32 nMatched = 0, nAQ = 0, nAT = 0, i = mi, j = 0, qb = 0, qe = 0, tb = 0, te = 0, beginOfNext = 0, nAnch = 0; bool firstA = true;
    for (;;) { j = (i32)Pd[rlo + i] - rlo; if (j < 0) break; u64 sv = C[rlo + i]; i32 sq = a_q(sv), sl = a_len(sv), stt = a_t(sv); nAnch++;
      if (firstA) { firstA = false; qe = sq + sl - 1; te = stt + sl - 1; qb = sq; tb = stt; nMatched += sl; } else { qb = sq; tb = stt; if (sq + sl - 1 >= beginOfNext) nMatched += beginOfNext - sq; else nMatched += sl; }
      beginOfNext = sq;
      if (i == j) { nAQ += qe - qb + 1; if (nAQ < P.min_align_len) break; nAT += te - tb + 1; double pid = (double)nMatched / (double)max(nAQ, nAT) * 100; if (pid < 15.0) break; emit(qb, qe, tb, te, nAQ, nAT, nMatched, nAnch); break; }
      i = j; }
    if (j < 0 && nAnch > 0) { nAQ += qe - qb + 1; nAT += te - tb + 1; if (nAQ >= P.min_align_len) { double pid = (double)nMatched / (double)max(nAQ, nAT) * 100; if (pid >= 15.0) emit(qb, qe, tb, te, nAQ, nAT, nMatched, nAnch); } }
    __syncwarp();
    if (i > 0) { if (lane == 0) ST[sp] = ((u64)(u32)rlo << 32) | (u32)(rlo + i); sp++; }                 // left, processed after the right region
    if (mi != len - 1) { if (lane == 0) ST[sp] = ((u64)(u32)(rlo + mi + 1) << 32) | (u32)rhi; sp++; }   // right
    __syncwarp();
  }
}

// =====================================================================================================
// K5: flank extension (extendMatch/_extendRight lib-index-search-util.go:34-201, Chainer3 lib-chaining3.go:111-299)
//     and wavefront alignment (github.com/shenwei356/wfa v0.5.0 Align; penalties 4/6/2, end-to-end, WFA2 backtrace priorities)
// =====================================================================================================
struct HspJob {          // one Chain2Result to align (host-built, lib-index-search.go:2223-2255 / :2490-2522)
  u32 q, g; i32 tBegin, tEnd; u32 rc; i32 qlen, tlen;          // window: target = genome[tBegin..tEnd] (reverse-complemented when rc), length tlen
  i32 start1, end1, start2, end2;                              // query [start1,end1) and window [start2,end2) before extension
  i32 ext, tb_arg, max_ext;                                    // _extLen2, c.TBegin, c.MaxExtLen
};
struct ExtOut { i32 s1, e1, s2, e2; i32 qs, qe, ts, te; };     // extension lengths and final half-open segments

struct SeqView { const u8* q2; const u8* g2; i32 tBegin, tEnd; u32 rc; const u8* qm; };
__device__ __forceinline__ u32 qcmp(const SeqView& v, i32 i) { return get_base(v.q2, (u64)i) | (((u32)(v.qm[i >> 3] >> (i & 7)) & 1u) << 2); }   // 4..7 = never equal to a target base
__device__ __forceinline__ u32 qbase(const SeqView& v, i32 i) { return get_base(v.q2, (u64)i); }
__device__ __forceinline__ u32 tbase(const SeqView& v, i32 i) { return win_base(v.g2, v.tBegin, v.tEnd, v.rc, i); }

// number of equal 2-mer pairs between q[qa..qa+n1) and t[ta..ta+n2) (dirq/dirt = +1 forward, -1 reversed flank)
__device__ u32 ext_count(const SeqView& v, i32 qa, i32 n1, int dq, i32 ta, i32 n2, int dt) {
  if (n1 < 2 || n2 < 2) return 0; u64 a0 = 0, a1 = 0, b0 = 0, b1 = 0;   // 16 byte-wide counters per sequence packed in two words (flanks <= 130 bases)
  u32 p = qbase(v, qa); for (i32 i = 1; i < n1; i++) { u32 c = qbase(v, qa + dq * i); u32 m = (p << 2) | c; if (m < 8) a0 += 1ull << (8 * m); else a1 += 1ull << (8 * (m - 8)); p = c; }
  p = tbase(v, ta); for (i32 i = 1; i < n2; i++) { u32 c = tbase(v, ta + dt * i); u32 m = (p << 2) | c; if (m < 8) b0 += 1ull << (8 * m); else b1 += 1ull << (8 * (m - 8)); p = c; }
  u32 t = 0; for (int i = 0; i < 8; i++) { t += (u32)((a0 >> (8 * i)) & 255) * (u32)((b0 >> (8 * i)) & 255) + (u32)((a1 >> (8 * i)) & 255) * (u32)((b1 >> (8 * i)) & 255); } return t;
}
// _extendRight, one WARP per (job, side): anchors (i1,i2) with equal 2-mers in (i1,i2) order, Chainer3 DP (MaxGap 5, MaxDistance 10, BandBase 10,
// BandCount 20), best chain end + 1. Anchor generation: the positions of each of the 16 target 2-mers as 192-bit sets in shared memory (flanks are
// at most 192 bases: checked on the host), rows (query 2-mers) expanded in parallel after a warp scan of their match counts. DP: anchors in order,
// the band of predecessors scanned 32 at a time (same ballot scheme as k_pa_chain). The thread-per-task version took 9 ms per batch: its critical
// path was a serial chain of dependent loads, and a single repetitive flank (thousands of anchors) held a whole wave.
__device__ void ext_run_warp(const SeqView& v, i32 qa, i32 n1, int dq, i32 ta, i32 n2, int dt, u16* __restrict__ anc, i16* __restrict__ sc, u16* __restrict__ pj, unsigned long long* __restrict__ Mk /*[48] shared, per warp*/, i32* e1, i32* e2) {
  const int lane = threadIdx.x & 31; *e1 = *e2 = 0; if (n1 < 2 || n2 < 2) return;
  for (int t = lane; t < 48; t += 32) Mk[t] = 0; __syncwarp();
  for (i32 i2 = lane; i2 + 1 < n2; i2 += 32) { u32 c = (tbase(v, ta + dt * i2) << 2) | tbase(v, ta + dt * (i2 + 1)); atomicOr(&Mk[c * 3 + (i2 >> 6)], 1ull << (i2 & 63)); } __syncwarp();
  u32 n = 0;
  for (i32 r0 = 0; r0 + 1 < n1; r0 += 32) { i32 i1 = r0 + lane; u32 c = 0, cnt = 0; bool act = i1 + 1 < n1;
    if (act) { c = (qbase(v, qa + dq * i1) << 2) | qbase(v, qa + dq * (i1 + 1)); cnt = __popcll(Mk[c * 3]) + __popcll(Mk[c * 3 + 1]) + __popcll(Mk[c * 3 + 2]); }
    u32 inc = cnt; for (int o = 1; o < 32; o <<= 1) { u32 t = __shfl_up_sync(FULLMASK, inc, o); if (lane >= o) inc += t; }
    u32 w = n + inc - cnt; if (act) for (int wd = 0; wd < 3; wd++) { unsigned long long mk = Mk[c * 3 + wd]; while (mk) { int b = __ffsll((long long)mk) - 1; mk &= mk - 1; anc[w++] = (u16)((i1 << 8) | (wd * 64 + b)); } }
    n += __shfl_sync(FULLMASK, inc, 31); }
  __syncwarp(); if (n == 0) return;
  i32 M = 0; u32 Mi = 0;
  for (u32 i = 0; i < n; i++) { u32 ai = anc[i]; i32 aq = ai >> 8, at = ai & 255; i32 m0 = 2 - max(aq, at) - abs(aq - at); i32 bs = INT32_MIN, bj = -1, cnt = 0; bool stop = false;   // m0 = Len - distance2(origin) - gap2(origin)
    for (i32 top = (i32)i - 1; top >= 0 && !stop; top -= 32) { i32 j = top - lane; bool nonskip = false; i32 bq = 0, bt = 0;
      if (j >= 0) { u32 aj = anc[j]; bq = aj >> 8; bt = aj & 255; nonskip = !(bq == aq || bt > at); }
      u32 bal = __ballot_sync(FULLMASK, nonskip); i32 mycnt = cnt + __popc(bal & ((2u << lane) - 1));
      bool brk = nonskip && !((aq - bq - 2) <= 10 || mycnt <= 20); u32 bb = __ballot_sync(FULLMASK, brk); int firstBrk = bb ? (__ffs(bb) - 1) : 32; if (bb) stop = true;
      i32 s = INT32_MIN; if (nonskip && lane < firstBrk) { i32 d = max(abs(aq - bq), abs(at - bt)); i32 g = abs(abs(aq - bq) - abs(at - bt)); if (d <= 10 && g <= 5) s = (i32)sc[j] + 2 - d - g; }
      i32 rs = s, rj = (s == INT32_MIN) ? INT32_MAX : j;
      { i32 mx = __reduce_max_sync(FULLMASK, rs); rj = __reduce_min_sync(FULLMASK, rs == mx ? rj : INT32_MAX); rs = mx; }
      if (rj != INT32_MAX && rs >= bs) { bs = rs; bj = rj; }   // s >= m: the later (smaller j) wins ties
      cnt += __popc(bal); }
    i32 m = m0; u32 mj = i; if (bj >= 0 && bs >= m0) { m = bs; mj = (u32)bj; }
    if (lane == 0) { sc[i] = (i16)m; pj[i] = (u16)mj; } if (i >= 1 && m > M) { M = m; Mi = i; }
    __syncwarp(); }
  if (M < 1) return;
  i32 i = (i32)Mi, nMatched = 0, beginOfNext = 0, qb = 0, qe = 0, tb = 0, te = 0; bool firstA = true;
  for (;;) { i32 j = pj[i]; u32 ai = anc[i]; i32 sq = ai >> 8, stt = ai & 255;
    if (firstA) { firstA = false; qe = sq + 1; te = stt + 1; qb = sq; tb = stt; nMatched += 2; } else { qb = sq; tb = stt; if (sq + 1 >= beginOfNext) nMatched += beginOfNext - sq; else nMatched += 2; }
    beginOfNext = sq;
    if (i == j) { i32 nAQ = qe - qb + 1; if (nAQ < 2) return; i32 nAT = te - tb + 1; double pid = (double)nMatched / (double)max(nAQ, nAT) * 100; if (pid < 15.0) return; *e1 = qe + 1; *e2 = te + 1; return; }
    i = j; }
}
__device__ __forceinline__ void ext_sides(const HspJob& J, i32* rext, i32* lext) {   // extension lengths for the 3' and 5' sides (0 = no attempt)
  *rext = 0; *lext = 0;
  if (J.end1 + 2 < J.qlen && J.end2 + 2 < J.tlen) { i32 e = J.rc ? min(J.ext, J.tb_arg) : min(J.ext, J.max_ext); if (e > 2) *rext = e; }
  if (J.start1 > 2 && J.start2 > 2) { i32 e = J.rc ? min(J.ext, J.max_ext) : min(J.ext, J.tb_arg); if (e > 2) *lext = e; }
}
__device__ __forceinline__ bool ext_task(const HspJob& J, int side, i32& qa, i32& n1, i32& ta, i32& n2, int& d) {   // flank geometry of one side; false = no extension attempt
  i32 rext, lext; ext_sides(J, &rext, &lext);
  if (side == 0) { if (!rext) return false; qa = J.end1; n1 = min(J.end1 + rext, J.qlen) - J.end1; ta = J.end2; n2 = min(J.end2 + rext, J.tlen) - J.end2; d = 1; }
  else { if (!lext) return false; i32 s1 = max(J.start1 - lext, 0), s2 = max(J.start2 - lext, 0); qa = J.start1 - 1; n1 = J.start1 - s1; ta = J.start2 - 1; n2 = J.start2 - s2; d = -1; }
  return true;
}
// scratch sizes: one thread per (job, side)
__global__ void k_extend_count(const HspJob* __restrict__ jobs, u32 njobs, const u8* __restrict__ qpacked, const u64* __restrict__ qboff, const u8* __restrict__ g2bit, const u64* __restrict__ g_off, u32* __restrict__ counts) {
  u32 t = blockIdx.x * blockDim.x + threadIdx.x; if (t >= njobs * 2) return; u32 jb = t >> 1; int side = t & 1; HspJob J = jobs[jb];
  SeqView v; v.q2 = qpacked + qboff[J.q]; v.qm = nullptr; v.g2 = g2bit + g_off[J.g]; v.tBegin = J.tBegin; v.tEnd = J.tEnd; v.rc = J.rc;
  i32 qa, n1, ta, n2; int d; counts[t] = ext_task(J, side, qa, n1, ta, n2, d) ? ext_count(v, qa, n1, d, ta, n2, d) : 0;
}
// extension: one warp per (job, side)
__global__ void __launch_bounds__(128) k_extend_run(const HspJob* __restrict__ jobs, u32 njobs, const u8* __restrict__ qpacked, const u64* __restrict__ qboff, const u8* __restrict__ g2bit, const u64* __restrict__ g_off,
                                                    const u64* __restrict__ soff, u16* __restrict__ anc, i16* __restrict__ sc, u16* __restrict__ pj, i32* __restrict__ res /*2 per (job,side)*/) {
  __shared__ unsigned long long Mk[4][48];
  u32 t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; if (t >= njobs * 2) return; u32 jb = t >> 1; int side = t & 1, lane = threadIdx.x & 31; HspJob J = jobs[jb];
  SeqView v; v.q2 = qpacked + qboff[J.q]; v.qm = nullptr; v.g2 = g2bit + g_off[J.g]; v.tBegin = J.tBegin; v.tEnd = J.tEnd; v.rc = J.rc;
  i32 qa, n1, ta, n2; int d; i32 e1 = 0, e2 = 0;
  if (ext_task(J, side, qa, n1, ta, n2, d)) { u64 o = soff[t]; ext_run_warp(v, qa, n1, d, ta, n2, d, anc + o, sc + o, pj + o, Mk[threadIdx.x >> 5], &e1, &e2); }
  if (lane == 0) { res[2 * t] = e1; res[2 * t + 1] = e2; }
}
__global__ void k_extend_final(const HspJob* __restrict__ jobs, u32 njobs, const i32* __restrict__ res, ExtOut* __restrict__ out) {
  u32 jb = blockIdx.x * blockDim.x + threadIdx.x; if (jb >= njobs) return; HspJob J = jobs[jb]; ExtOut o; o.e1 = res[4 * jb]; o.e2 = res[4 * jb + 1]; o.s1 = res[4 * jb + 2]; o.s2 = res[4 * jb + 3];
  i32 start1 = J.start1, end1 = J.end1, start2 = J.start2, end2 = J.end2;
  if (o.e1 > 0 || o.e2 > 0) { end1 += o.e1; end2 += o.e2; } if (o.s1 > 0 || o.s2 > 0) { start1 -= o.s1; start2 -= o.s2; }
  if (start1 < 0 || start2 < 0) { start1 = J.start1; start2 = J.start2; } if (end1 > J.qlen || end2 > J.tlen) { end1 = J.end1; end2 = J.end2; }
  o.qs = start1; o.qe = end1; o.ts = start2; o.te = end2; out[jb] = o;
}

// ---------------- WFA
#define WF_NULL (-1073741824)
struct WfaOut { i32 qbegin, qend, tbegin, tend, alen, matches, gaps, bscore, has_m, wscore, status; u32 ops_n; u64 ops_off; };   // status 0 ok, 1 workspace overflow
struct WfDir { i32 lo, hi; u32 base; u32 nullmask; i32 elo, ehi; };   // [lo,hi] = stored range (M at base, I at base+w, D at base+2w); [elo,ehi] = effective range after adaptive reduction; nullmask bit0 M, bit1 I, bit2 D

// One warp per alignment; persistent warps pull jobs from a queue. Wavefront offsets live in a per-warp HBM slab
// (directory + offsets), lanes own diagonals, extension compares bases straight from the 2-bit genome / query arrays.
__device__ __forceinline__ u64 fetch64(const u64* __restrict__ W, i32 pos);
__global__ void __launch_bounds__(128) k_wfa(const HspJob* __restrict__ jobs, const ExtOut* __restrict__ ext, const u32* __restrict__ job_ids, u32 njobs, u32* __restrict__ next_job, const u64* __restrict__ woff, const u64* __restrict__ words, const u32* __restrict__ has_amb,
                                             const u8* __restrict__ qpacked, const u8* __restrict__ qamask, const u64* __restrict__ qboff, const u8* __restrict__ g2bit, const u64* __restrict__ g_off,
                                             i32* __restrict__ slabs, u64 slab_words, WfaOut* __restrict__ outs, u64* __restrict__ ops_pool, u64* __restrict__ ops_cursor, u64 ops_cap, int want_ops, int adaptive) {
  const int X = 4, OE = 8, E = 2, STEP = 2; int lane = threadIdx.x & 31; u32 warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i32* slab = slabs + (u64)warp * slab_words;
  for (;;) {
    u32 ji = 0; if (lane == 0) ji = atomicAdd(next_job, 1u); ji = __shfl_sync(FULLMASK, ji, 0); if (ji >= njobs) return; u32 jb = job_ids[ji];
    HspJob J = jobs[jb]; ExtOut ex = ext[jb]; SeqView v; v.q2 = qpacked + qboff[J.q]; v.qm = qamask + qboff[J.q]; v.g2 = g2bit + g_off[J.g]; v.tBegin = J.tBegin; v.tEnd = J.tEnd; v.rc = J.rc;
    const i32 q0 = ex.qs, t0 = ex.ts, plen = ex.qe - ex.qs, tlen = ex.te - ex.ts, kend = tlen - plen;
    WfaOut R; R.qbegin = R.qend = R.tbegin = R.tend = R.alen = R.matches = R.gaps = R.bscore = R.has_m = 0; R.wscore = 0; R.status = 0; R.ops_n = 0; R.ops_off = 0;
    // slab layout: [directory: ndir x 4 words][offsets ...]
    const u32 ndir = (u32)((4 * (i64)max(plen, tlen) + 16) / STEP + 2); WfDir* dir = (WfDir*)slab; u64 used = (u64)ndir * 6; bool overflow = used + 64 > slab_words;
    i32* offs = slab;
    // extension on the packed words of k_wfa_prep (32 bases per XOR + CLZ; ambiguous query bases never match), as in k_wfa_fast / k_wfa_reg
    const u32 nqw = (u32)((plen + 31) / 32 + 2); const u64* Qw = words + woff[2 * jb]; const u64* Aw = Qw + nqw; const u64* Tw = words + woff[2 * jb + 1]; const bool amb = has_amb[jb] != 0; (void)v; (void)q0; (void)t0;
    auto extend = [&](i32 k, i32 h) { i32 vv = h - k; for (;;) { const i32 rem = min(plen - vv, tlen - h); if (rem <= 0) break; u64 x = fetch64(Qw, vv) ^ fetch64(Tw, h); if (amb) x |= fetch64(Aw, vv); i32 n = x ? (__clzll(x) >> 1) : 32; n = min(n, rem); vv += n; h += n; if (n < 32) break; } return h; };
    // wavefront value through a directory record held in registers (the three source levels of a score are loaded once per level)
    auto getd = [&](const WfDir& d, int comp, i32 k) -> i32 { if ((d.nullmask >> comp) & 1) return WF_NULL; if (k < d.lo || k > d.hi) return WF_NULL; return offs[d.base + (u32)comp * (u32)(d.hi - d.lo + 1) + (u32)(k - d.lo)]; };
    auto getw = [&](i32 sidx, int comp, i32 k) -> i32 { if (sidx < 0) return WF_NULL; WfDir d = dir[sidx]; if ((d.nullmask >> comp) & 1) return WF_NULL; if (k < d.lo || k > d.hi) return WF_NULL; return offs[d.base + (u32)comp * (u32)(d.hi - d.lo + 1) + (u32)(k - d.lo)]; };
    i32 s = 0; bool done = false;
    if (!overflow) { if (lane == 0) { WfDir d; d.lo = 0; d.hi = 0; d.elo = 0; d.ehi = 0; d.base = (u32)used; d.nullmask = 6; dir[0] = d; offs[used] = extend(0, 0); } used += 3; __syncwarp(); done = (getw(0, 0, kend) >= tlen); }
    while (!done && !overflow) {
      s += STEP; i32 si = s / STEP; if ((u32)si >= ndir) { overflow = true; break; }
      i32 ix = (s - X) / STEP, io = (s - OE) / STEP, ie = (s - E) / STEP; bool hx = s - X >= 0, ho = s - OE >= 0, he = s - E >= 0;
      WfDir dx, dop, de; dx.nullmask = 7; dop.nullmask = 7; de.nullmask = 7; dx.lo = dx.hi = dop.lo = dop.hi = de.lo = de.hi = 0; dx.elo = dx.ehi = dop.elo = dop.ehi = de.elo = de.ehi = 0; if (hx) dx = dir[ix]; if (ho) dop = dir[io]; if (he) de = dir[ie];
      bool nx = dx.nullmask & 1, no = dop.nullmask & 1, ni = (de.nullmask >> 1) & 1, nd = (de.nullmask >> 2) & 1;
      if (nx && no && ni && nd) { if (lane == 0) { WfDir d; d.lo = 0; d.hi = -1; d.elo = 0; d.ehi = -1; d.base = 0; d.nullmask = 7; dir[si] = d; } __syncwarp(); continue; }
      i32 lo = INT32_MAX, hi = INT32_MIN; if (!nx) { lo = min(lo, dx.elo); hi = max(hi, dx.ehi); } if (!no) { lo = min(lo, dop.elo - 1); hi = max(hi, dop.ehi + 1); } if (!ni || !nd) { lo = min(lo, de.elo - 1); hi = max(hi, de.ehi + 1); }
      u32 w = (u32)(hi - lo + 1); if (used + 3ull * w + 64 > slab_words) { overflow = true; break; }
      u32 base = (u32)used; bool anyM = false, anyI = false, anyD = false;
      for (i32 k = lo + lane; k <= hi; k += 32) {
        i32 a = getd(dop, 0, k - 1), b2 = getd(de, 1, k - 1); i32 ins = max(a, b2); ins = (ins <= WF_NULL) ? WF_NULL : ins + 1;
        a = getd(dop, 0, k + 1); b2 = getd(de, 2, k + 1); i32 del = max(a, b2);
        i32 mis = getd(dx, 0, k); mis = (mis <= WF_NULL) ? WF_NULL : mis + 1;
        if (!(ins > WF_NULL && ins >= 0 && ins - k >= 0 && ins <= tlen && ins - k <= plen)) ins = WF_NULL;
        if (!(del > WF_NULL && del >= 0 && del - k >= 0 && del <= tlen && del - k <= plen)) del = WF_NULL;
        if (!(mis > WF_NULL && mis >= 0 && mis - k >= 0 && mis <= tlen && mis - k <= plen)) mis = WF_NULL;
        i32 mm = max(mis, max(ins, del)); if (mm > WF_NULL) { mm = extend(k, mm); anyM = true; } anyI |= ins > WF_NULL; anyD |= del > WF_NULL;
        offs[base + (u32)(k - lo)] = mm; offs[base + w + (u32)(k - lo)] = ins; offs[base + 2 * w + (u32)(k - lo)] = del;
      }
      anyM = __any_sync(FULLMASK, anyM); anyI = __any_sync(FULLMASK, anyI); anyD = __any_sync(FULLMASK, anyD); __syncwarp();
      i32 elo = lo, ehi = hi;
      if (adaptive && anyM && hi - lo + 1 >= 10) {   // WFA-adaptive reduction (MinWFLen 10, MaxDistDiff 50), target diagonal preserved
        i32 mind = INT32_MAX; for (i32 k = lo + lane; k <= hi; k += 32) { i32 o = offs[base + (u32)(k - lo)]; if (o > WF_NULL) mind = min(mind, max(plen - (o - k), tlen - o)); }
        for (int o = 16; o; o >>= 1) mind = min(mind, __shfl_xor_sync(FULLMASK, mind, o));
        i32 top_limit = min(kend, hi); i32 nlo = lo; bool found = false;
        for (i32 b0 = lo; b0 < top_limit && !found; b0 += 32) { i32 k = b0 + lane; bool okk = false; if (k < top_limit) { i32 o = offs[base + (u32)(k - lo)]; okk = (o > WF_NULL) && (max(plen - (o - k), tlen - o) - mind <= 50); } u32 bal = __ballot_sync(FULLMASK, okk); if (bal) { nlo = b0 + __ffs(bal) - 1; found = true; } }
        if (!found && top_limit > lo) nlo = top_limit;
        i32 bottom_limit = max(kend, nlo); i32 nhi = hi; found = false;
        for (i32 b0 = hi; b0 > bottom_limit && !found; b0 -= 32) { i32 k = b0 - lane; bool okk = false; if (k > bottom_limit) { i32 o = offs[base + (u32)(k - lo)]; okk = (o > WF_NULL) && (max(plen - (o - k), tlen - o) - mind <= 50); } u32 bal = __ballot_sync(FULLMASK, okk); if (bal) { nhi = b0 - (__ffs(bal) - 1); found = true; } }
        if (!found && hi > bottom_limit) nhi = bottom_limit;
        if (nlo != lo || nhi != hi) { for (i32 k = lo + lane; k <= hi; k += 32) if (k < nlo || k > nhi) { offs[base + (u32)(k - lo)] = WF_NULL; offs[base + w + (u32)(k - lo)] = WF_NULL; offs[base + 2 * w + (u32)(k - lo)] = WF_NULL; }
          __syncwarp(); bool aM = false, aI = false, aD = false; for (i32 k = nlo + lane; k <= nhi; k += 32) { aM |= offs[base + (u32)(k - lo)] > WF_NULL; aI |= offs[base + w + (u32)(k - lo)] > WF_NULL; aD |= offs[base + 2 * w + (u32)(k - lo)] > WF_NULL; }
          anyM = __any_sync(FULLMASK, aM); anyI = __any_sync(FULLMASK, aI); anyD = __any_sync(FULLMASK, aD); elo = nlo; ehi = nhi; }
      }
      if (lane == 0) { WfDir d; d.lo = lo; d.hi = hi; d.elo = elo; d.ehi = ehi; d.base = base; d.nullmask = (anyM ? 0 : 1) | (anyI ? 0 : 2) | (anyD ? 0 : 4); dir[si] = d; }
      used += 3ull * w; __syncwarp();
      done = (getw(si, 0, kend) >= tlen);
    }
    if (overflow) { if (lane == 0) { R.status = 1; outs[jb] = R; } __syncwarp(); continue; }
    // ---- backtrace (lane 0): WFA2 priority mismatch(9) > D ext(6) > D open(5) > I ext(2) > I open(1); statistics over first-M..last-M
    if (lane == 0) {
      R.wscore = s; i32 k = kend, off = tlen, sc = s; int mat = 0; i32 vv = off - k, h = off; u64* ops = (u64*)(slab + ((used + 1) & ~1ull)); u64 ops_room = (slab_words - ((used + 1) & ~1ull)) / 2; u32 nops = 0; bool ops_over = false;
      int cur = 0; u32 curn = 0; i32 p_alen = 0, p_gaps = 0, p_bs = 0; int prev = 0;
      auto flush = [&]() { if (want_ops && curn) { if (nops < ops_room) ops[nops++] = ((u64)cur << 32) | curn; else ops_over = true; } curn = 0; };
      auto put = [&](int op, i32 cntp, i32 vend, i32 hend) {   // op run of cntp ending (exclusive) at query vend / target hend
        if (cntp <= 0) return; if (op != cur) { flush(); cur = op; } curn += (u32)cntp;
        if (op == 'M') { if (R.has_m) { R.alen += p_alen; R.gaps += p_gaps; R.bscore += p_bs; } else { R.has_m = 1; R.qend = vend; R.tend = hend; } p_alen = p_gaps = p_bs = 0; R.alen += cntp; R.matches += cntp; R.bscore += 2 * cntp; R.qbegin = vend - cntp + 1; R.tbegin = hend - cntp + 1; }
        else if (op == 'X') { p_alen += cntp; p_bs -= 3 * cntp; } else { p_alen += cntp; p_gaps += cntp; p_bs -= 2 * cntp; if (prev != op) p_bs -= 5; }
        prev = op; };
      while (vv > 0 && h > 0 && sc > 0) {
        i32 s_mis = sc - X, s_open = sc - OE, s_ext = sc - E; i64 c_mis = INT64_MIN, c_io = INT64_MIN, c_ie = INT64_MIN, c_do = INT64_MIN, c_de = INT64_MIN;
        auto pig = [](i32 o, int type) -> i64 { return o <= WF_NULL ? INT64_MIN : (((i64)o << 4) | type); };
        if (mat == 0) { i32 o = (s_mis >= 0) ? getw(s_mis / STEP, 0, k) : WF_NULL; c_mis = pig(o <= WF_NULL ? WF_NULL : o + 1, 9); }
        if (mat == 0 || mat == 1) { i32 o = (s_open >= 0) ? getw(s_open / STEP, 0, k - 1) : WF_NULL; c_io = pig(o <= WF_NULL ? WF_NULL : o + 1, 1); o = (s_ext >= 0) ? getw(s_ext / STEP, 1, k - 1) : WF_NULL; c_ie = pig(o <= WF_NULL ? WF_NULL : o + 1, 2); }
        if (mat == 0 || mat == 2) { c_do = pig((s_open >= 0) ? getw(s_open / STEP, 0, k + 1) : WF_NULL, 5); c_de = pig((s_ext >= 0) ? getw(s_ext / STEP, 2, k + 1) : WF_NULL, 6); }
        i64 best = max(c_mis, max(max(c_io, c_ie), max(c_do, c_de))); if (best == INT64_MIN) { R.status = 2; break; }
        if (mat == 0) { i32 mo = (i32)(best >> 4); i32 nm = off - mo; put('M', nm, off - k, off); off = mo; vv = off - k; h = off; if (vv <= 0 || h <= 0) continue; }
        int type = (int)(best & 15);
        switch (type) { case 9: sc = s_mis; mat = 0; put('X', 1, off - k, off); off--; break;
          case 1: sc = s_open; mat = 0; put('I', 1, off - k, off); k--; off--; break; case 2: sc = s_ext; mat = 1; put('I', 1, off - k, off); k--; off--; break;
          case 5: sc = s_open; mat = 0; put('D', 1, off - k, off); k++; break; case 6: sc = s_ext; mat = 2; put('D', 1, off - k, off); k++; break; }
        vv = off - k; h = off;
      }
      if (R.status == 0) { if (sc == 0) put('M', off, off - k, off); else { if (vv > 0) put('D', vv, vv, h); if (h > 0) put('I', h, 0, h); } }
      flush();
      if (want_ops && R.status == 0) { if (ops_over) R.status = 1; else { u64 o = atomicAdd((unsigned long long*)ops_cursor, (unsigned long long)nops); if (o + nops <= ops_cap) { for (u32 i = 0; i < nops; i++) ops_pool[o + i] = ops[i]; R.ops_off = o; R.ops_n = nops; } else R.status = 3; } }
      outs[jb] = R;
    }
    __syncwarp();
  }
}

// ---------------- WFA fast path: packed sequences, shared-memory wavefront ring, directory-free HBM layout
// Per job the query / target segments are first re-packed (k_wfa_prep) into 64-bit words (32 bases each, first base in the top
// bits; a third word stream marks ambiguous query bases) so that wavefront extension compares 32 bases per XOR+CLZ.
// Wavefronts live in HBM as u16 offsets in a fixed-stride layout: level L (score 2L), component c, diagonal k at
// slab[L*3*WFS + c*WFS + k + WFK0]; every level writes [glo-PAD, ghi+PAD] (NULL outside its reach) so neither the forward pass nor
// the backtrace needs bounds checks or a directory. The last 5 M levels and 2 I/D levels are mirrored in shared memory (ring).
// Jobs that do not fit (|k| >= 128, score >= 2*WF_LMAX, sequences >= 32000) report status 1 and go to the general kernel k_wfa.
#define WF_WMAX 256
#define WF_PAD 8
#define WFS (WF_WMAX + 2 * WF_PAD)
#define WFK0 (WF_WMAX / 2 + WF_PAD)
#define WF_LMAX 2048
#define WF_OPSMAX 4096
#define WF_WARPS 8
#define WF_SEQW 144
#define X2_LV 2
#define OE2_LV 4
#define E2_LV 1
#define WF_SMEM_BYTES ((size_t)WF_WARPS * (WF_SEQW * 8 + 9 * WFS * 2 + WFS * 2))
struct WfaSeg { u64 qw, tw; };   // word offsets of the job's packed query / target (query: 2 streams: bases at qw, ambiguity at qw + nqw)

__global__ void k_wfa_prep(const HspJob* __restrict__ jobs, const ExtOut* __restrict__ ext, u32 njobs, const u64* __restrict__ woff /*2 per job +1*/, const u8* __restrict__ qpacked, const u8* __restrict__ qamask, const u64* __restrict__ qboff,
                           const u8* __restrict__ g2bit, const u64* __restrict__ g_off, u64* __restrict__ words, u32* __restrict__ has_amb) {
  u32 jb = blockIdx.x; if (jb >= njobs) return; if (threadIdx.x == 0) has_amb[jb] = 0; __syncthreads(); HspJob J = jobs[jb]; ExtOut e = ext[jb]; const u8* q2 = qpacked + qboff[J.q]; const u8* qm = qamask + qboff[J.q]; const u8* g2 = g2bit + g_off[J.g];
  i32 plen = e.qe - e.qs, tlen = e.te - e.ts; u32 nq = (u32)((plen + 31) / 32 + 2), nt = (u32)((tlen + 31) / 32 + 2); u64* Q = words + woff[2 * jb]; u64* A = Q + nq; u64* T = words + woff[2 * jb + 1];
  for (u32 w = threadIdx.x; w < nq; w += blockDim.x) { u64 b = 0, a = 0; for (int j = 0; j < 32; j++) { i32 i = (i32)w * 32 + j; u64 c = 0, am = 0; if (i < plen) { i32 p = e.qs + i; c = get_base(q2, (u64)p); am = ((qm[p >> 3] >> (p & 7)) & 1) ? 3 : 0; } b = (b << 2) | c; a = (a << 2) | am; } Q[w] = b; A[w] = a; if (a) has_amb[jb] = 1; }
  for (u32 w = threadIdx.x; w < nt; w += blockDim.x) { u64 b = 0; for (int j = 0; j < 32; j++) { i32 i = (i32)w * 32 + j; u64 c = (i < tlen) ? win_base(g2, J.tBegin, J.tEnd, J.rc, e.ts + i) : 0; b = (b << 2) | c; } T[w] = b; }
}
__device__ __forceinline__ u64 fetch64(const u64* __restrict__ W, i32 pos) { u32 i = (u32)pos >> 5, sh = ((u32)pos & 31) * 2; u64 a = W[i]; if (sh == 0) return a; return (a << sh) | (W[i + 1] >> (64 - sh)); }

__global__ void __launch_bounds__(WF_WARPS * 32) k_wfa_fast(const ExtOut* __restrict__ ext, const u64* __restrict__ woff, const u64* __restrict__ words, const u32* __restrict__ has_amb, const u32* __restrict__ job_ids, u32 job0, u32 njobs, u32* __restrict__ next_job,
                                                          u16* __restrict__ slabs, WfaOut* __restrict__ outs, int adaptive, int lmax) {
  extern __shared__ __align__(16) u8 wf_smem[];   // per warp: packed sequences | ring of 9 wavefronts (0-4: M levels L%5, 5-6: I L%2, 7-8: D L%2) | distances of the current M wavefront
  const int X2 = 2, OE2 = 4, E2 = 1;        // penalties 4 / 8 / 2 in units of levels (score = 2*level)
  int lane = threadIdx.x & 31, wib = threadIdx.x >> 5; u64* seqb = (u64*)wf_smem + (size_t)wib * WF_SEQW; u16 (*R)[WFS] = (u16 (*)[WFS])(wf_smem + (size_t)WF_WARPS * WF_SEQW * 8) + (size_t)wib * 9;
  u16* Dst = (u16*)(wf_smem + (size_t)WF_WARPS * WF_SEQW * 8 + (size_t)WF_WARPS * 9 * WFS * 2) + (size_t)wib * WFS;
  for (;;) {
    u32 jr = 0; if (lane == 0) jr = atomicAdd(next_job, 1u); jr = __shfl_sync(FULLMASK, jr, 0); if (jr >= njobs) return; const u32 jb = job_ids ? job_ids[job0 + jr] : job0 + jr; u16* slab = slabs + (u64)jr * lmax * 3 * WFS;   // one slab per alignment of the round
    ExtOut ex = ext[jb]; const i32 plen = ex.qe - ex.qs, tlen = ex.te - ex.ts, kend = tlen - plen; const u32 nqw = (u32)((plen + 31) / 32 + 2);
    const u64* Q = words + woff[2 * jb]; const u64* A = Q + nqw; const u64* T = words + woff[2 * jb + 1]; const bool amb = has_amb[jb] != 0;
    { // stage the packed sequences in shared memory when they fit (alignments of up to ~3 kb each side): the extension loop is a chain of dependent word fetches
      const u32 ntw = (u32)((tlen + 31) / 32 + 2), nA = amb ? nqw : 0; __syncwarp();
      if (nqw + nA + ntw <= WF_SEQW) { u64* sq = seqb; for (u32 i = lane; i < nqw + nA; i += 32) sq[i] = Q[i]; for (u32 i = lane; i < ntw; i += 32) sq[nqw + nA + i] = T[i]; __syncwarp(); Q = sq; A = sq + nqw; T = sq + nqw + nA; } }
    WfaOut Rz; Rz.qbegin = Rz.qend = Rz.tbegin = Rz.tend = Rz.alen = Rz.matches = Rz.gaps = Rz.bscore = Rz.has_m = 0; Rz.wscore = 0; Rz.status = 0; Rz.ops_n = 0; Rz.ops_off = 0;
    if (plen >= 32000 || tlen >= 32000 || kend <= -(WF_WMAX / 2) + 2 || kend >= WF_WMAX / 2 - 2 || plen <= 0 || tlen <= 0) { if (lane == 0) { Rz.status = 1; outs[jb] = Rz; } __syncwarp(); continue; }
    auto extend = [&](i32 k, i32 h) { i32 v = h - k; for (;;) { i32 rem = min(plen - v, tlen - h); if (rem <= 0) break; u64 x = fetch64(Q, v) ^ fetch64(T, h); if (amb) x |= fetch64(A, v); i32 n = x ? (__clzll(x) >> 1) : 32; n = min(n, rem); v += n; h += n; if (n < 32) break; } return h; };
    // level 0. History of levels L-1..L-4: effective lo/hi (after reduction), null bits (1 M, 2 I, 4 D). Every level writes the diagonals
    // [min(lo, rlo) - PAD, max(hi, rhi) + PAD] where [rlo,rhi] spans the effective ranges of the last 4 non-null levels, so later levels and
    // the backtrace can read k-1 / k+1 of any source level without bounds checks.
    i32 hlo[4], hhi[4]; u32 hnull[4]; for (int i = 0; i < 4; i++) { hlo[i] = 0; hhi[i] = -1; hnull[i] = 7; }
    for (int t = lane; t < 9 * WFS / 2; t += 32) ((u32*)&R[0][0])[t] = 0xFFFFFFFFu; __syncwarp();   // every ring row starts null: levels below 0 need no special case
    { i32 h0 = 0; if (lane == 0) h0 = extend(0, 0); h0 = __shfl_sync(FULLMASK, h0, 0);
      for (i32 k = -WF_PAD + lane; k <= WF_PAD; k += 32) { u16 mv = (k == 0) ? (u16)h0 : (u16)0xFFFF; R[0][k + WFK0] = mv; slab[k + WFK0] = mv; slab[WFS + k + WFK0] = 0xFFFF; slab[2 * WFS + k + WFK0] = 0xFFFF; R[5][k + WFK0] = 0xFFFF; R[7][k + WFK0] = 0xFFFF; }
      hlo[0] = 0; hhi[0] = 0; hnull[0] = 6; __syncwarp(); }
    i32 L = 0; bool done = (kend == 0 && R[0][WFK0] != 0xFFFF && (i32)R[0][WFK0] >= tlen), overflow = false;
    while (!done) {
      L++; if (L >= lmax) { overflow = true; break; }
      bool nx = (L - X2 < 0) || (hnull[X2 - 1] & 1), no = (L - OE2 < 0) || (hnull[OE2 - 1] & 1), ni = (hnull[E2 - 1] >> 1) & 1, nd = (hnull[E2 - 1] >> 2) & 1;
      i32 lo = INT32_MAX, hi = INT32_MIN; bool allnull = nx && no && ni && nd;
      if (!allnull) { if (!nx) { lo = min(lo, hlo[X2 - 1]); hi = max(hi, hhi[X2 - 1]); } if (!no) { lo = min(lo, hlo[OE2 - 1] - 1); hi = max(hi, hhi[OE2 - 1] + 1); } if (!ni || !nd) { lo = min(lo, hlo[E2 - 1] - 1); hi = max(hi, hhi[E2 - 1] + 1); } }
      i32 rlo = INT32_MAX, rhi = INT32_MIN; for (int i = 0; i < 4; i++) if (hnull[i] != 7) { rlo = min(rlo, hlo[i]); rhi = max(rhi, hhi[i]); }
      i32 wl = min(allnull ? INT32_MAX : lo, rlo), wh = max(allnull ? INT32_MIN : hi, rhi); if (wl > wh) { wl = 0; wh = 0; }
      if (wl <= -(WF_WMAX / 2) || wh >= WF_WMAX / 2) { overflow = true; break; }
      const u16* M2 = R[(L + 5 - X2) % 5]; const u16* M4 = R[(L + 5 - OE2) % 5]; const u16* I1 = R[5 + ((L + 1) & 1)]; const u16* D1 = R[7 + ((L + 1) & 1)];
      u16* Mo = R[L % 5]; u16* Io = R[5 + (L & 1)]; u16* Do = R[7 + (L & 1)]; u16* G = slab + (u64)L * 3 * WFS; bool anyM = false, anyI = false, anyD = false;
      u32 bitM = 0, bitI = 0, bitD = 0, itb = 1; i32 mind = INT32_MAX; const i32 k0 = wl - WF_PAD + lane;   // bit `it` of bitX: this lane's cell of iteration `it` (k = k0 + 32*it) is non-null
      for (i32 k = k0; k <= wh + WF_PAD; k += 32, itb <<= 1) {
        u16 om = 0xFFFF, oi = 0xFFFF, od = 0xFFFF, dd = 0xFFFF; int x = k + WFK0;
        if (!allnull && k >= lo && k <= hi) {
          // offsets are < 32000 here, so the u16 cells read as signed 16-bit give -1 for NULL (0xFFFF) directly
          i32 ins = max((i32)(i16)M4[x - 1], (i32)(i16)I1[x - 1]); ins = (ins < 0) ? -1 : ins + 1;
          i32 del = max((i32)(i16)M4[x + 1], (i32)(i16)D1[x + 1]);
          i32 mis = (i32)(i16)M2[x]; mis = (mis < 0) ? -1 : mis + 1;
          const i32 hmax = min(tlen, plen + k);   // h <= tlen and v = h - k <= plen; h >= 0 and v >= 0 hold by construction
          if (ins > hmax) ins = -1; if (del > hmax || del - k < 0) del = -1; if (mis > hmax) mis = -1;
          i32 mm = max(mis, max(ins, del)); if (mm >= 0) { mm = extend(k, mm); bitM |= itb; om = (u16)mm; i32 dv = max(plen - (mm - k), tlen - mm); dd = (u16)dv; mind = min(mind, dv); } if (ins >= 0) { bitI |= itb; oi = (u16)ins; } if (del >= 0) { bitD |= itb; od = (u16)del; }
        }
        Mo[x] = om; Io[x] = oi; Do[x] = od; Dst[x] = dd; G[x] = om; G[WFS + x] = oi; G[2 * WFS + x] = od;
      }
      anyM = __any_sync(FULLMASK, bitM != 0); anyI = __any_sync(FULLMASK, bitI != 0); anyD = __any_sync(FULLMASK, bitD != 0); __syncwarp();
      if (adaptive && !allnull && anyM && hi - lo + 1 >= 10) {   // WFA-adaptive reduction, same rule as k_wfa and the oracle
        for (int o = 16; o; o >>= 1) mind = min(mind, __shfl_xor_sync(FULLMASK, mind, o));   // distances (max of the remaining query / target lengths) were taken in the cell loop; Dst holds them, 0xFFFF = null
        const i32 thr = mind + 50; i32 top_limit = min(kend, hi); i32 nlo = lo; bool found = false;
        for (i32 b0 = lo; b0 < top_limit && !found; b0 += 32) { i32 k = b0 + lane; bool okk = (k < top_limit) && ((i32)Dst[k + WFK0] <= thr); u32 bal = __ballot_sync(FULLMASK, okk); if (bal) { nlo = b0 + __ffs(bal) - 1; found = true; } }
        if (!found && top_limit > lo) nlo = top_limit;
        i32 bottom_limit = max(kend, nlo); i32 nhi = hi; found = false;
        for (i32 b0 = hi; b0 > bottom_limit && !found; b0 -= 32) { i32 k = b0 - lane; bool okk = (k > bottom_limit) && ((i32)Dst[k + WFK0] <= thr); u32 bal = __ballot_sync(FULLMASK, okk); if (bal) { nhi = b0 - (__ffs(bal) - 1); found = true; } }
        if (!found && hi > bottom_limit) nhi = bottom_limit;
        if (nlo != lo || nhi != hi) { for (i32 k = lo + lane; k <= hi; k += 32) if (k < nlo || k > nhi) { int x = k + WFK0; Mo[x] = 0xFFFF; Io[x] = 0xFFFF; Do[x] = 0xFFFF; G[x] = 0xFFFF; G[WFS + x] = 0xFFFF; G[2 * WFS + x] = 0xFFFF; }
          // which of this lane's cells (k = k0 + 32*it) survive: it in [ceil((nlo-k0)/32), floor((nhi-k0)/32)]
          i32 a0 = nlo - k0, a1 = nhi - k0; i32 itlo = a0 <= 0 ? 0 : (a0 + 31) >> 5, ithi = a1 < 0 ? -1 : (a1 >> 5); u32 keep = (ithi < itlo) ? 0u : ((ithi >= 31 ? 0xFFFFFFFFu : ((2u << ithi) - 1)) & ~((1u << itlo) - 1));
          anyM = __any_sync(FULLMASK, (bitM & keep) != 0); anyI = __any_sync(FULLMASK, (bitI & keep) != 0); anyD = __any_sync(FULLMASK, (bitD & keep) != 0); lo = nlo; hi = nhi; __syncwarp(); }
      }
      for (int i = 3; i > 0; i--) { hlo[i] = hlo[i - 1]; hhi[i] = hhi[i - 1]; hnull[i] = hnull[i - 1]; }
      hlo[0] = allnull ? 0 : lo; hhi[0] = allnull ? -1 : hi; hnull[0] = allnull ? 7u : ((anyM ? 0u : 1u) | (anyI ? 0u : 2u) | (anyD ? 0u : 4u));
      __syncwarp();
      if (!allnull && kend >= lo && kend <= hi) { u16 v = Mo[kend + WFK0]; done = (v != 0xFFFF && (i32)v >= tlen); }
    }
    if (lane == 0) { Rz.status = overflow ? 1 : 0; Rz.wscore = 2 * L; outs[jb] = Rz; }   // the backtrace runs in k_wfa_bt, one thread per alignment
    __syncwarp();
  }
}
// Backtrace of the fast path, ONE THREAD per alignment: the walk is a chain of dependent HBM reads (~1 us each), so 32 of them per warp
// (instead of lane 0 only) hide the latency. Same decision rule as k_wfa: mismatch(9) > D ext(6) > D open(5) > I ext(2) > I open(1).
__global__ void __launch_bounds__(128) k_wfa_bt(const ExtOut* __restrict__ ext, const u32* __restrict__ job_ids, u32 job0, u32 njobs, const u16* __restrict__ slabs, u64* __restrict__ ops_scratch, WfaOut* __restrict__ outs, u64* __restrict__ ops_pool, u64* __restrict__ ops_cursor, u64 ops_cap, int want_ops, int lmax) {
  const int X2 = 2, OE2 = 4, E2 = 1; u32 t = blockIdx.x * blockDim.x + threadIdx.x; if (t >= njobs) return; u32 jb = job_ids ? job_ids[job0 + t] : job0 + t; WfaOut Rr = outs[jb]; if (Rr.status != 0) return;
  ExtOut ex = ext[jb]; const i32 plen = ex.qe - ex.qs, tlen = ex.te - ex.ts, kend = tlen - plen; const u16* slab = slabs + (u64)t * lmax * 3 * WFS; u64* ops = ops_scratch + (u64)t * WF_OPSMAX;
  i32 L = Rr.wscore / 2; i32 k = kend, off = tlen, lv = L; int mat = 0; i32 vv = off - k, h = off; u32 nops = 0; bool ops_over = false; int cur = 0; u32 curn = 0; i32 p_alen = 0, p_gaps = 0, p_bs = 0; int prev = 0; (void)plen;
  auto flush = [&]() { if (want_ops && curn) { if (nops < WF_OPSMAX) ops[nops++] = ((u64)cur << 32) | curn; else ops_over = true; } curn = 0; };
  auto put = [&](int op, i32 cntp, i32 vend, i32 hend) { if (cntp <= 0) return; if (op != cur) { flush(); cur = op; } curn += (u32)cntp;
    if (op == 'M') { if (Rr.has_m) { Rr.alen += p_alen; Rr.gaps += p_gaps; Rr.bscore += p_bs; } else { Rr.has_m = 1; Rr.qend = vend; Rr.tend = hend; } p_alen = p_gaps = p_bs = 0; Rr.alen += cntp; Rr.matches += cntp; Rr.bscore += 2 * cntp; Rr.qbegin = vend - cntp + 1; Rr.tbegin = hend - cntp + 1; }
    else if (op == 'X') { p_alen += cntp; p_bs -= 3 * cntp; } else { p_alen += cntp; p_gaps += cntp; p_bs -= 2 * cntp; if (prev != op) p_bs -= 5; } prev = op; };
  auto ld = [&](i32 lvl, int comp, i32 kk) -> i32 { if (lvl < 0) return -1; u16 v = slab[(u64)lvl * 3 * WFS + comp * WFS + kk + WFK0]; return v == 0xFFFF ? -1 : (i32)v; };
  auto pig = [](i32 o, int type) -> i64 { return o < 0 ? INT64_MIN : (((i64)o << 4) | type); };
  while (vv > 0 && h > 0 && lv > 0) {
    i32 l_mis = lv - X2, l_open = lv - OE2, l_ext = lv - E2; i64 c_mis = INT64_MIN, c_io = INT64_MIN, c_ie = INT64_MIN, c_do = INT64_MIN, c_de = INT64_MIN;
    i32 v_mis = -1, v_io = -1, v_ie = -1, v_do = -1, v_de = -1;
    if (mat == 0) v_mis = ld(l_mis, 0, k); if (mat == 0 || mat == 1) { v_io = ld(l_open, 0, k - 1); v_ie = ld(l_ext, 1, k - 1); } if (mat == 0 || mat == 2) { v_do = ld(l_open, 0, k + 1); v_de = ld(l_ext, 2, k + 1); }
    if (mat == 0) c_mis = pig(v_mis < 0 ? -1 : v_mis + 1, 9); if (mat == 0 || mat == 1) { c_io = pig(v_io < 0 ? -1 : v_io + 1, 1); c_ie = pig(v_ie < 0 ? -1 : v_ie + 1, 2); } if (mat == 0 || mat == 2) { c_do = pig(v_do, 5); c_de = pig(v_de, 6); }
    i64 best = max(c_mis, max(max(c_io, c_ie), max(c_do, c_de))); if (best == INT64_MIN) { Rr.status = 2; break; }
    if (mat == 0) { i32 mo = (i32)(best >> 4); i32 nm = off - mo; put('M', nm, off - k, off); off = mo; vv = off - k; h = off; if (vv <= 0 || h <= 0) continue; }
    int type = (int)(best & 15);
    switch (type) { case 9: lv = l_mis; mat = 0; put('X', 1, off - k, off); off--; break;
      case 1: lv = l_open; mat = 0; put('I', 1, off - k, off); k--; off--; break; case 2: lv = l_ext; mat = 1; put('I', 1, off - k, off); k--; off--; break;
      case 5: lv = l_open; mat = 0; put('D', 1, off - k, off); k++; break; case 6: lv = l_ext; mat = 2; put('D', 1, off - k, off); k++; break; }
    vv = off - k; h = off;
  }
  if (Rr.status == 0) { if (lv == 0) put('M', off, off - k, off); else { if (vv > 0) put('D', vv, vv, h); if (h > 0) put('I', h, 0, h); } }
  flush();
  if (want_ops && Rr.status == 0) { if (ops_over) Rr.status = 1; else { u64 o = atomicAdd((unsigned long long*)ops_cursor, (unsigned long long)nops); if (o + nops <= ops_cap) { for (u32 i = 0; i < nops; i++) ops_pool[o + i] = ops[i]; Rr.ops_off = o; Rr.ops_n = nops; } else Rr.status = 3; } }
  outs[jb] = Rr;
}

// ---------------- WFA register path: one diagonal per lane, wavefronts in registers
// After WFA-adaptive reduction (MaxDistDiff 50) the live band of an alignment at <= ~15 % divergence is 10-20 diagonals wide, so the whole
// wavefront of a level fits the 32 lanes of a warp with one FIXED diagonal per lane (k = kb + lane, kb chosen so that diagonals 0 and
// kend = tlen - plen sit inside the window). The M offsets of the last four levels and the I / D offsets of the last level then live in
// registers, neighbours (k-1, k+1) come from warp shuffles, and a level costs ~200 warp instructions instead of the ~850 of the
// shared-memory-ring kernel (k_wfa_fast: range loops over a padded band, ring index arithmetic, three HBM stores per cell).
// Backtrace data: ONE u32 per (level, lane) — the backtrace decision taken with exactly the values the backtrace of k_wfa_fast would
// load: pre-extension offset (15 bits) | source of the M cell (4 bits: 9 mismatch, 6 D-ext, 5 D-open, 2 I-ext, 1 I-open) | "I came from
// I-ext" | "D came from D-ext" — 128 B per level (k_wfa_fast: 1,632 B), written as one coalesced line; the backtrace reads one word per step.
// A level whose range leaves the window (wide bands: high divergence, long gaps) reports status 1 and the alignment is redone by k_wfa_fast /
// k_wfa: results are identical whichever kernel runs (same recurrences, same reduction rule, same tie-breaking).
#define WR_WARPS 8
#define WR_SEQW 144
#define WR_SMEM_BYTES ((size_t)WR_WARPS * WR_SEQW * 8)
#define WR_MAXLEN (1 << 20)      // offsets are stored in 20 bits of the backtrace word
#define WR_NS 4                  // diagonals per lane ("slots"): lane l holds window positions l, 32 + l, 64 + l, 96 + l
#define WR_W (32 * WR_NS)        // diagonals in the window
// The 128-diagonal window follows the band: kb moves (registers shifted by warp shuffles) whenever the range of the new level or of one of the four
// levels it reads from would touch the window's edge positions; the alignment falls back to k_wfa_fast / k_wfa only when that span itself exceeds
// 126 diagonals. Only the slots the span touches do any work (warp-uniform branches): bands of up to 30 diagonals (<= ~12 % divergence) are kept in
// slot 0, ~25 % divergence needs two or three slots. Window position 0 never holds a live cell: its slab word stores kb of the level for the backtrace.
// This is what lets long, indel-rich alignments (ONT reads: |tlen - plen| of hundreds, tens of thousands of levels) stay on the register path.
struct WrCell { i32 om, oi, od, dv; u32 word; };
template <int NS>
__global__ void __launch_bounds__(WR_WARPS * 32) k_wfa_reg(const ExtOut* __restrict__ ext, const u64* __restrict__ woff, const u64* __restrict__ words, const u32* __restrict__ has_amb, const u32* __restrict__ job_ids, u32 job0, u32 njobs, u32* __restrict__ next_job,
                                                         u32* __restrict__ slabs, const u64* __restrict__ slab_off, WfaOut* __restrict__ outs, int adaptive) {
  extern __shared__ __align__(16) u8 wr_smem[];
  constexpr int W = 32 * NS;   /* diagonals in the window */
  const int X2 = 2, OE2 = 4, E2 = 1; const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5; u64* seqb = (u64*)wr_smem + (size_t)wib * WR_SEQW; const int lm1 = (lane + 31) & 31, lp1 = (lane + 1) & 31;
  for (;;) {
    u32 jr = 0; if (lane == 0) jr = atomicAdd(next_job, 1u); jr = __shfl_sync(FULLMASK, jr, 0); if (jr >= njobs) return; const u32 jb = job_ids ? job_ids[job0 + jr] : job0 + jr;
    u32* slab = slabs + slab_off[jr] * W; const i32 lmax = (i32)(slab_off[jr + 1] - slab_off[jr]);   // levels this alignment may use
    ExtOut ex = ext[jb]; const i32 plen = ex.qe - ex.qs, tlen = ex.te - ex.ts, kend = tlen - plen; const u32 nqw = (u32)((plen + 31) / 32 + 2);
    const u64* Q = words + woff[2 * jb]; const u64* A = Q + nqw; const u64* T = words + woff[2 * jb + 1]; const bool amb = has_amb[jb] != 0;
    { const u32 ntw = (u32)((tlen + 31) / 32 + 2), nA = amb ? nqw : 0; __syncwarp();
      if (nqw + nA + ntw <= WR_SEQW) { u64* sq = seqb; for (u32 i = lane; i < nqw + nA; i += 32) sq[i] = Q[i]; for (u32 i = lane; i < ntw; i += 32) sq[nqw + nA + i] = T[i]; __syncwarp(); Q = sq; A = sq + nqw; T = sq + nqw + nA; } }
    WfaOut Rz; Rz.qbegin = Rz.qend = Rz.tbegin = Rz.tend = Rz.alen = Rz.matches = Rz.gaps = Rz.bscore = Rz.has_m = 0; Rz.wscore = 0; Rz.status = 0; Rz.ops_n = 0; Rz.ops_off = 0;
    if (plen >= WR_MAXLEN || tlen >= WR_MAXLEN || plen <= 0 || tlen <= 0 || lmax < 2) { if (lane == 0) { Rz.status = 1; Rz.wscore = -3; outs[jb] = Rz; } __syncwarp(); continue; }
    auto extend = [&](i32 kk, i32 h) { i32 v = h - kk; for (;;) { i32 rem = min(plen - v, tlen - h); if (rem <= 0) break; u64 x = fetch64(Q, v) ^ fetch64(T, h); if (amb) x |= fetch64(A, v); i32 n = x ? (__clzll(x) >> 1) : 32; n = min(n, rem); v += n; h += n; if (n < 32) break; } return h; };
    // one cell: recurrences + extension + the backtrace word (decisions from the unfiltered neighbour values: offset first, then mismatch(9) > D-ext(6) > D-open(5) > I-ext(2) > I-open(1))
    auto cell = [&](i32 k, bool live, i32 c_m2, i32 m4l, i32 i1l, i32 m4r, i32 d1r) { WrCell c; c.om = -1; c.oi = -1; c.od = -1; c.dv = INT32_MAX; c.word = 0;
      if (live) { i32 ins = max(m4l, i1l); ins = (ins < 0) ? -1 : ins + 1; i32 del = max(m4r, d1r); i32 mis = (c_m2 < 0) ? -1 : c_m2 + 1;
        const i32 hmax = min(tlen, plen + k); if (ins > hmax) ins = -1; if (del > hmax || del - k < 0) del = -1; if (mis > hmax) mis = -1;
        i32 mm = max(mis, max(ins, del)); if (mm >= 0) { c.om = extend(k, mm); c.dv = max(plen - (c.om - k), tlen - c.om); } c.oi = ins; c.od = del;
        i32 best = -1; if (c_m2 >= 0) best = ((c_m2 + 1) << 4) | 9; if (m4l >= 0) best = max(best, ((m4l + 1) << 4) | 1); if (i1l >= 0) best = max(best, ((i1l + 1) << 4) | 2); if (m4r >= 0) best = max(best, (m4r << 4) | 5); if (d1r >= 0) best = max(best, (d1r << 4) | 6);
        if (best >= 0) c.word = (u32)(best >> 4) | ((u32)(best & 15) << 20);
        if (i1l >= 0 && i1l >= m4l) c.word |= 1u << 24; if (d1r >= 0 && d1r >= m4r) c.word |= 1u << 25; if (m4l < 0 && i1l < 0) c.word |= 1u << 26; if (m4r < 0 && d1r < 0) c.word |= 1u << 27; }   // bits 26/27: no source for the I / D state (backtrace error as in k_wfa_bt)
      return c; };
    i32 hlo[4], hhi[4]; u32 hnull[4]; for (int i = 0; i < 4; i++) { hlo[i] = 0; hhi[i] = -1; hnull[i] = 7; }
    i32 kb = -16;   // diagonal 0 on lane 16 of slot 0
    i32 m1[NS], m2[NS], m3[NS], m4[NS], i1[NS], d1[NS];   // per slot: M of levels L-1..L-4, I / D of level L-1 on this lane's diagonals; -1 = null
#pragma unroll
    for (int s_ = 0; s_ < NS; s_++) { m1[s_] = m2[s_] = m3[s_] = m4[s_] = i1[s_] = d1[s_] = -1; }
    if (kb + lane == 0) m1[0] = extend(0, 0); hlo[0] = 0; hhi[0] = 0; hnull[0] = 6; slab[lane] = lane == 0 ? (u32)kb : 0u;
    auto at = [&](const i32 (&v)[NS], i32 j) { i32 r = -1;   // value on window position j, warp-uniform j
#pragma unroll
      for (int t = 0; t < NS; t++) { const i32 x = __shfl_sync(FULLMASK, v[t], j & 31); if ((j >> 5) == t) r = x; } return r; };
    i32 L = 0, why = 0; bool done = false, overflow = false; { const i32 v0 = at(m1, 0 - kb); done = (kend == 0 && v0 >= tlen); }   // why: reason for leaving the register path (-1 level capacity, -2 band wider than the window, -3 not eligible)
    while (!done) {
      L++; if (L >= lmax) { overflow = true; why = -1; break; }
      const bool nx = (L - X2 < 0) || (hnull[X2 - 1] & 1), no = (L - OE2 < 0) || (hnull[OE2 - 1] & 1), ni = (hnull[E2 - 1] >> 1) & 1, nd = (hnull[E2 - 1] >> 2) & 1;
      i32 lo = INT32_MAX, hi = INT32_MIN; const bool allnull = nx && no && ni && nd;
      if (!allnull) { if (!nx) { lo = min(lo, hlo[X2 - 1]); hi = max(hi, hhi[X2 - 1]); } if (!no) { lo = min(lo, hlo[OE2 - 1] - 1); hi = max(hi, hhi[OE2 - 1] + 1); } if (!ni || !nd) { lo = min(lo, hlo[E2 - 1] - 1); hi = max(hi, hhi[E2 - 1] + 1); } }
      // the window must hold this level's range and the ranges of the levels still in registers on positions 1..W-2 (the edge positions stay null: k-1 / k+1 of a live cell is always inside)
      i32 slo = allnull ? INT32_MAX : lo, shi = allnull ? INT32_MIN : hi; for (int i = 0; i < 4; i++) if (hnull[i] != 7) { slo = min(slo, hlo[i]); shi = max(shi, hhi[i]); }
      if (slo <= shi) { const i32 span = shi - slo + 1;
        if (slo < kb + 1 || shi > kb + W - 2 || (span <= 30 && shi > kb + 30)) { if (span > W - 2) { overflow = true; why = -2; break; }
          const i32 room = min(W - 2, ((span + 33) >> 5) * 32 - 2); const i32 nkb = slo - 1 - (room - span) / 2, dlt = nkb - kb;   // centred in as few slots as the span needs
          const int sl = (lane + (dlt & 31)) & 31, carry = (lane + (dlt & 31)) >> 5, q = dlt >> 5;   // new position p takes old position p + dlt: old lane sl, old slot s + q + carry
          auto shift = [&](i32 (&v)[NS]) { i32 x[NS];
#pragma unroll
            for (int t = 0; t < NS; t++) x[t] = __shfl_sync(FULLMASK, v[t], sl);
#pragma unroll
            for (int s_ = 0; s_ < NS; s_++) { const int src = s_ + q + carry; i32 r = -1;
#pragma unroll
              for (int t = 0; t < NS; t++) if (src == t) r = x[t]; v[s_] = r; } };
          shift(m1); shift(m2); shift(m3); shift(m4); shift(i1); shift(d1); kb = nkb; } }
      const int sa = (slo <= shi) ? max(0, (slo - kb) >> 5) : 0, sb = (slo <= shi) ? min(NS - 1, (shi - kb) >> 5) : -1;   // slots that can hold anything but nulls
      WrCell c[NS];
#pragma unroll
      for (int s_ = 0; s_ < NS; s_++) { c[s_].om = c[s_].oi = c[s_].od = -1; c[s_].dv = INT32_MAX; c[s_].word = 0; }
      if (!allnull) {   // raw neighbour values (what the backtrace of k_wfa_fast loads from its slab); positions 31|32, 63|64, 95|96 are the seams between the slots
        i32 r4[NS], ri[NS], t4[NS], td[NS];
#pragma unroll
        for (int t = 0; t < NS; t++) { r4[t] = ri[t] = t4[t] = td[t] = -1; if (t >= sa && t <= sb) { r4[t] = __shfl_sync(FULLMASK, m4[t], lm1); ri[t] = __shfl_sync(FULLMASK, i1[t], lm1); t4[t] = __shfl_sync(FULLMASK, m4[t], lp1); td[t] = __shfl_sync(FULLMASK, d1[t], lp1); } }
#pragma unroll
        for (int s_ = 0; s_ < NS; s_++) if (s_ >= sa && s_ <= sb && lo <= kb + 32 * s_ + 31 && hi >= kb + 32 * s_) { const i32 k = kb + 32 * s_ + lane;
            const i32 l4 = lane ? r4[s_] : (s_ ? r4[s_ - 1] : -1), li = lane ? ri[s_] : (s_ ? ri[s_ - 1] : -1), g4 = lane < 31 ? t4[s_] : (s_ < NS - 1 ? t4[s_ + 1] : -1), gd = lane < 31 ? td[s_] : (s_ < NS - 1 ? td[s_ + 1] : -1);
            c[s_] = cell(k, k >= lo && k <= hi, m2[s_], l4, li, g4, gd); } }
      bool hm = false, hi_ = false, hd = false;
#pragma unroll
      for (int s_ = 0; s_ < NS; s_++) { hm |= c[s_].om >= 0; hi_ |= c[s_].oi >= 0; hd |= c[s_].od >= 0; }
      bool anyM = __any_sync(FULLMASK, hm), anyI = __any_sync(FULLMASK, hi_), anyD = __any_sync(FULLMASK, hd);
      if (adaptive && !allnull && anyM && hi - lo + 1 >= 10) {   // WFA-adaptive reduction, same rule as k_wfa / k_wfa_fast / the oracle
        i32 mind = INT32_MAX;
#pragma unroll
        for (int s_ = 0; s_ < NS; s_++) mind = min(mind, c[s_].dv);
        for (int o = 16; o; o >>= 1) mind = min(mind, __shfl_xor_sync(FULLMASK, mind, o)); const i32 thr = mind + 50;   // distances of null cells are +inf
        const i32 top_limit = min(kend, hi); i32 nlo = lo; bool found = false;
#pragma unroll
        for (int s_ = 0; s_ < NS; s_++) if (s_ >= sa && s_ <= sb) { const i32 k = kb + 32 * s_ + lane; const u32 bal = __ballot_sync(FULLMASK, c[s_].om >= 0 && c[s_].dv <= thr && k >= lo && k < top_limit); if (bal && !found) { nlo = kb + 32 * s_ + __ffs(bal) - 1; found = true; } }
        if (!found && top_limit > lo) nlo = top_limit;
        const i32 bottom_limit = max(kend, nlo); i32 nhi = hi; found = false;
#pragma unroll
        for (int s_ = NS - 1; s_ >= 0; s_--) if (s_ >= sa && s_ <= sb) { const i32 k = kb + 32 * s_ + lane; const u32 bal = __ballot_sync(FULLMASK, c[s_].om >= 0 && c[s_].dv <= thr && k <= hi && k > bottom_limit); if (bal && !found) { nhi = kb + 32 * s_ + 31 - __clz(bal); found = true; } }
        if (!found && hi > bottom_limit) nhi = bottom_limit;
        if (nlo != lo || nhi != hi) { hm = hi_ = hd = false;
#pragma unroll
          for (int s_ = 0; s_ < NS; s_++) { const i32 k = kb + 32 * s_ + lane; if (k < nlo || k > nhi) { c[s_].om = -1; c[s_].oi = -1; c[s_].od = -1; } hm |= c[s_].om >= 0; hi_ |= c[s_].oi >= 0; hd |= c[s_].od >= 0; }
          anyM = __any_sync(FULLMASK, hm); anyI = __any_sync(FULLMASK, hi_); anyD = __any_sync(FULLMASK, hd); lo = nlo; hi = nhi; }
      }
      { u32* row = slab + (u64)L * W;
#pragma unroll
        for (int s_ = 0; s_ < NS; s_++) if (s_ == 0 || (s_ >= sa && s_ <= sb)) row[32 * s_ + lane] = (s_ == 0 && lane == 0) ? (u32)kb : c[s_].word; }
#pragma unroll
      for (int s_ = 0; s_ < NS; s_++) { m4[s_] = m3[s_]; m3[s_] = m2[s_]; m2[s_] = m1[s_]; m1[s_] = c[s_].om; i1[s_] = c[s_].oi; d1[s_] = c[s_].od; }
      for (int i = 3; i > 0; i--) { hlo[i] = hlo[i - 1]; hhi[i] = hhi[i - 1]; hnull[i] = hnull[i - 1]; }
      hlo[0] = allnull ? 0 : lo; hhi[0] = allnull ? -1 : hi; hnull[0] = allnull ? 7u : ((anyM ? 0u : 1u) | (anyI ? 0u : 2u) | (anyD ? 0u : 4u));
      if (!allnull && kend >= lo && kend <= hi) { const i32 v = at(m1, kend - kb); done = (v >= tlen); }
    }
    if (lane == 0) { Rz.status = overflow ? 1 : 0; Rz.wscore = overflow ? why : 2 * L; outs[jb] = Rz; }
    __syncwarp();
  }
}
// backtrace of the register path, one thread per alignment: two words of one 128-byte line per step (the level's window base and the cell)
__global__ void __launch_bounds__(128) k_wfa_bt2(const ExtOut* __restrict__ ext, const u32* __restrict__ job_ids, u32 job0, u32 njobs, const u32* __restrict__ slabs, const u64* __restrict__ slab_off, u64* __restrict__ ops_scratch, const u64* __restrict__ ops_off, WfaOut* __restrict__ outs, u64* __restrict__ ops_pool, u64* __restrict__ ops_cursor, u64 ops_cap, int want_ops, int W) {   /* W: diagonals per level row (32 x the slots of the forward kernel) */
  u32 t = blockIdx.x * blockDim.x + threadIdx.x; if (t >= njobs) return; u32 jb = job_ids ? job_ids[job0 + t] : job0 + t; WfaOut Rr = outs[jb]; if (Rr.status != 0) return;
  ExtOut ex = ext[jb]; const i32 plen = ex.qe - ex.qs, tlen = ex.te - ex.ts, kend = tlen - plen; const u32* slab = slabs + slab_off[t] * W; u64* ops = ops_scratch + (want_ops ? ops_off[t] : 0); const u32 ops_max = want_ops ? (u32)(ops_off[t + 1] - ops_off[t]) : 0;
  i32 k = kend, off = tlen, lv = Rr.wscore / 2; int mat = 0; i32 vv = off - k, h = off; u32 nops = 0; bool ops_over = false; int cur = 0; u32 curn = 0; i32 p_alen = 0, p_gaps = 0, p_bs = 0; int prev = 0;
  auto flush = [&]() { if (want_ops && curn) { if (nops < ops_max) ops[nops++] = ((u64)cur << 32) | curn; else ops_over = true; } curn = 0; };
  auto put = [&](int op, i32 cntp, i32 vend, i32 hend) { if (cntp <= 0) return; if (op != cur) { flush(); cur = op; } curn += (u32)cntp;
    if (op == 'M') { if (Rr.has_m) { Rr.alen += p_alen; Rr.gaps += p_gaps; Rr.bscore += p_bs; } else { Rr.has_m = 1; Rr.qend = vend; Rr.tend = hend; } p_alen = p_gaps = p_bs = 0; Rr.alen += cntp; Rr.matches += cntp; Rr.bscore += 2 * cntp; Rr.qbegin = vend - cntp + 1; Rr.tbegin = hend - cntp + 1; }
    else if (op == 'X') { p_alen += cntp; p_bs -= 3 * cntp; } else { p_alen += cntp; p_gaps += cntp; p_bs -= 2 * cntp; if (prev != op) p_bs -= 5; } prev = op; };
  while (vv > 0 && h > 0 && lv > 0) {
    const u32* row = slab + (u64)lv * W; const i32 kb = (i32)row[0]; const u32 li = (u32)(k - kb); if (li - 1u > (u32)(W - 3)) { Rr.status = 2; break; } const u32 cell = row[li]; int type;
    if (mat == 0) { type = (int)((cell >> 20) & 15); if (type == 0) { Rr.status = 2; break; } const i32 mo = (i32)(cell & 0xFFFFF); put('M', off - mo, off - k, off); off = mo; vv = off - k; h = off; if (vv <= 0 || h <= 0) continue; }
    else if (mat == 1) { if ((cell >> 26) & 1) { Rr.status = 2; break; } type = ((cell >> 24) & 1) ? 2 : 1; }
    else { if ((cell >> 27) & 1) { Rr.status = 2; break; } type = ((cell >> 25) & 1) ? 6 : 5; }
    switch (type) { case 9: lv -= X2_LV; mat = 0; put('X', 1, off - k, off); off--; break;
      case 1: lv -= OE2_LV; mat = 0; put('I', 1, off - k, off); k--; off--; break; case 2: lv -= E2_LV; mat = 1; put('I', 1, off - k, off); k--; off--; break;
      case 5: lv -= OE2_LV; mat = 0; put('D', 1, off - k, off); k++; break; case 6: lv -= E2_LV; mat = 2; put('D', 1, off - k, off); k++; break; default: Rr.status = 2; break; }
    if (Rr.status) break; vv = off - k; h = off;
  }
  if (Rr.status == 0) { if (lv == 0) put('M', off, off - k, off); else { if (vv > 0) put('D', vv, vv, h); if (h > 0) put('I', h, 0, h); } }
  flush();
  if (want_ops && Rr.status == 0) { if (ops_over) Rr.status = 1; else { u64 o = atomicAdd((unsigned long long*)ops_cursor, (unsigned long long)nops); if (o + nops <= ops_cap) { for (u32 i = 0; i < nops; i++) ops_pool[o + i] = ops[i]; Rr.ops_off = o; Rr.ops_n = nops; } else Rr.status = 3; } }
  (void)plen; outs[jb] = Rr;
}

// the dynamic shared-memory ceiling of a kernel is a per-device attribute: raise it once on every device this process runs alignments on
static void wfa_raise_smem_limit() { static std::mutex mu; static std::set<int> done; int dev = 0; CUDA_CHECK(cudaGetDevice(&dev)); std::lock_guard<std::mutex> lk(mu); if (done.count(dev)) return;
  CUDA_CHECK(cudaFuncSetAttribute(k_wfa_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WF_SMEM_BYTES)); done.insert(dev); }
static void wfa_run_all(cudaStream_t st, int sm_count, DBuf<HspJob>& d_jobs, DBuf<ExtOut>& d_ext, const std::vector<ExtOut>& hext, u32 nj, const u8* qpacked, const u8* qamask, const u64* qboff, const u8* g2bit, const u64* g_off,
                        int want_ops, int adaptive, std::vector<WfaOut>& hw, std::vector<u64>& hops, u64* counters, double* ms, size_t total_mem, int active_lanes = 1) {
    // WFA: fast kernel (packed words + smem ring) for every job, then the general kernel for whatever did not fit
    wfa_raise_smem_limit();
    DBuf<WfaOut> d_out(nj, st); std::vector<u32> ids; hw.resize(nj);
    u64 ops_cap = 0; if (want_ops) { for (u32 j = 0; j < nj; j++) ops_cap += (u64)(hext[j].qe - hext[j].qs) + (hext[j].te - hext[j].ts) + 4; } DBuf<u64> ops_pool(ops_cap + 2, st); DBuf<u64> ops_cur(1, st); ops_cur.zero();
    DBuf<u64> woff, words; DBuf<u32> hasamb;   /* packed sequences of every job (k_wfa_prep): all four WFA kernels read them */
    { std::vector<u64> hwoff(2 * (u64)nj + 1, 0); for (u32 j = 0; j < nj; j++) { u64 nqw = (u64)((hext[j].qe - hext[j].qs + 31) / 32 + 2), ntw = (u64)((hext[j].te - hext[j].ts + 31) / 32 + 2); hwoff[2 * j + 1] = hwoff[2 * j] + 2 * nqw; hwoff[2 * j + 2] = hwoff[2 * j + 1] + ntw; }
      woff.alloc(hwoff.size(), st); woff.from_host(hwoff.data(), hwoff.size()); words.alloc(hwoff.back() + 4, st); hasamb.alloc(nj + 1, st);
      if (g_lap) (*g_lap)("wfa host prep"); { KTimer kt(st, &ms[10]); k_wfa_prep<<<nj, 64, 0, st>>>(d_jobs.p, d_ext.p, nj, woff.p, qpacked, qamask, qboff, g2bit, g_off, words.p, hasamb.p); KERNEL_CHECK(); }
      // Levels per slab follow the longest sequence of the batch (score <= 1.6 x length covers ~40 % divergence; deeper ones use k_wfa).
      i32 maxlen = 1; for (u32 j = 0; j < nj; j++) maxlen = std::max(maxlen, std::max(hext[j].qe - hext[j].qs, hext[j].te - hext[j].ts));
      const int lmax = (int)std::min<i64>(WF_LMAX * 2, std::max<i64>(256, ((i64)(0.8 * maxlen) + 63) / 64 * 64));
      const size_t tb_ = total_mem; const u64 budgetF = std::min<u64>((u64)(tb_ * 0.25), 36ull << 30) / (u64)std::max(1, active_lanes);
      if (g_lap) (*g_lap)("wfa prep kernel");
      // pass 1, every job: the register kernel (one diagonal per lane, 128 B of backtrace words per level). Rounds only when the slabs of all jobs exceed the budget.
      std::vector<u32> rest;   // jobs whose band left the 32-lane window (or too deep / too long): pass 2
      static const bool use_reg = getenv("LMG_NO_WFA_REG") == nullptr;
      if (use_reg) {   // per-job slabs: levels for ~40 % divergence of THIS alignment (0.8 x its longer side), op runs up to its length; rounds fill the HBM budget
        auto lv_cap = [&](u32 j) { const i64 len = std::max(hext[j].qe - hext[j].qs, hext[j].te - hext[j].ts); return (u64)std::max<i64>(256, ((i64)(0.8 * (double)len) + 63) / 64 * 64); };
        auto op_cap = [&](u32 j) { const i64 len = std::max(hext[j].qe - hext[j].qs, hext[j].te - hext[j].ts); return want_ops ? (u64)std::max<i64>(1024, len + 64) : 0ull; };
        std::vector<u32> skipped; u32 round_max = 0; std::vector<i32> len(nj); for (u32 j = 0; j < nj; j++) len[j] = std::max(hext[j].qe - hext[j].qs, hext[j].te - hext[j].ts);
        // one pass of the register kernel over the jobs `ord` with NS slots per lane (window of 32 x NS diagonals, 128 x NS bytes of backtrace words per level); rounds fill the HBM budget
        auto run_reg = [&](const std::vector<u32>& ord, int NS, std::vector<u32>& skip) { const u32 cnt = (u32)ord.size(); const u64 W = 32ull * (u64)NS; std::vector<u64> hso, hoo; DBuf<u32> d_order(cnt, st); d_order.from_host(ord.data(), cnt);
          KTimer kt(st, &ms[11]);
          for (u32 j0 = 0; j0 < cnt;) { hso.assign(1, 0); hoo.assign(1, 0); u64 bytes = 0; u32 n = 0;
            while (j0 + n < cnt) { const u32 jid = ord[j0 + n]; const u64 lv = lv_cap(jid), oc = op_cap(jid), b = lv * W * 4 + oc * 8; if (n > 0 && (bytes + b > budgetF || (g_arena == nullptr && n >= 8192))) break; if (n == 0 && b > budgetF) { skip.push_back(jid); j0++; continue; } hso.push_back(hso.back() + lv); hoo.push_back(hoo.back() + oc); bytes += b; n++; }
            if (n == 0) continue;
            DBuf<u64> soff(n + 1, st), ooff(n + 1, st); soff.from_host(hso.data(), n + 1); ooff.from_host(hoo.data(), n + 1); DBuf<u32> rslabs(hso.back() * W + 64, st); DBuf<u64> oscr(hoo.back() + 8, st); DBuf<u32> next(1, st); next.zero();
            const u32 blocks = (u32)std::min<u64>((u64)sm_count * 8, (n + WR_WARPS - 1) / WR_WARPS);
            if (NS == 4) k_wfa_reg<4><<<blocks, WR_WARPS * 32, WR_SMEM_BYTES, st>>>(d_ext.p, woff.p, words.p, hasamb.p, d_order.p, j0, n, next.p, rslabs.p, soff.p, d_out.p, adaptive);
            else k_wfa_reg<8><<<blocks, WR_WARPS * 32, WR_SMEM_BYTES, st>>>(d_ext.p, woff.p, words.p, hasamb.p, d_order.p, j0, n, next.p, rslabs.p, soff.p, d_out.p, adaptive);
            KERNEL_CHECK();
            k_wfa_bt2<<<cdiv(n, 128), 128, 0, st>>>(d_ext.p, d_order.p, j0, n, rslabs.p, soff.p, oscr.p, ooff.p, d_out.p, ops_pool.p, ops_cur.p, ops_cap, want_ops, (int)W); KERNEL_CHECK();
            CUDA_CHECK(cudaStreamSynchronize(st));   // the round's buffers go back to the arena
            j0 += n; round_max = std::max(round_max, n); } };
        // longest alignments first: the persistent warps pull jobs in this order, so the expensive ones start early and the tail of a launch is made of short ones
        std::vector<u32> order(nj); std::iota(order.begin(), order.end(), 0u); std::stable_sort(order.begin(), order.end(), [&](u32 a, u32 b) { return len[a] > len[b]; });
        run_reg(order, 4, skipped);
        std::vector<WfaOut> o = d_out.to_host(nj); std::vector<char> sk(nj, 0); for (u32 j : skipped) sk[j] = 1;
        u32 why[5] = {0, 0, 0, 0, 0}; std::vector<u32> wide;   // wide: the band left the 128-diagonal window
        for (u32 j = 0; j < nj; j++) { if (sk[j] || o[j].status == 1) { const int w = sk[j] ? 4 : (o[j].wscore == -1 ? 1 : o[j].wscore == -2 ? 2 : o[j].wscore == -3 ? 3 : 0); why[w]++; if (w == 2) wide.push_back(j); else rest.push_back(j); } else if (o[j].status != 0) throw std::runtime_error("WFA backtrace failed (register kernel)"); else hw[j] = o[j]; }
        if (g_lap && g_lap->on && (!rest.empty() || !wide.empty())) fprintf(stderr, "[lmg host] register WFA pass: %zu of %u alignments left (op scratch %u, level capacity %u, band > %d diagonals %u, not eligible %u, slab over budget %u)\n", rest.size() + wide.size(), nj, why[0], why[1], 32 * 4 - 2, why[2], why[3], why[4]);
        // pass 1b, wide bands (long, indel-rich alignments: ONT reads against the less similar genomes): the same kernel with 8 slots per lane = a 256-diagonal window
        static const bool use_reg8 = getenv("LMG_NO_WFA_REG8") == nullptr;
        if (!wide.empty() && use_reg8) { std::stable_sort(wide.begin(), wide.end(), [&](u32 a, u32 b) { return len[a] > len[b]; }); std::vector<u32> sk8; run_reg(wide, 8, sk8); std::vector<WfaOut> o8 = d_out.to_host(nj); std::vector<char> s8(nj, 0); for (u32 j : sk8) s8[j] = 1; u32 left = 0;
          for (u32 j : wide) { if (s8[j] || o8[j].status == 1) { rest.push_back(j); left++; } else if (o8[j].status != 0) throw std::runtime_error("WFA backtrace failed (register kernel, wide window)"); else hw[j] = o8[j]; }
          if (g_lap && g_lap->on) fprintf(stderr, "[lmg host] register WFA pass, 256-diagonal window: %u of %zu alignments left\n", left, wide.size()); }
        else rest.insert(rest.end(), wide.begin(), wide.end());
        counters[14] = round_max; if (g_lap) (*g_lap)("wfa register pass"); }
      else { rest.resize(nj); std::iota(rest.begin(), rest.end(), 0u); }
      counters[9] = nj; counters[10] = rest.size(); counters[15] = (u64)lmax;
      // pass 2, the jobs left: shared-memory-ring kernel (bands up to 256 diagonals), per-alignment slabs of 3 x u16 per cell in rounds bounded by the HBM budget
      if (!rest.empty()) { const u32 nr = (u32)rest.size(); DBuf<u32> d_rest(nr, st); d_rest.from_host(rest.data(), nr); const u64 slab_bytes = (u64)lmax * 3 * WFS * 2; const u32 nwarps = (u32)sm_count * 4 * WF_WARPS;
        u32 per_round = (u32)std::min<u64>(nr, std::max<u64>(nwarps, budgetF / slab_bytes)); if (per_round > nwarps) per_round = per_round / nwarps * nwarps; if (g_arena == nullptr) per_round = std::min<u32>(per_round, 2048);
        DBuf<u16> fslabs((u64)per_round * lmax * 3 * WFS, st); DBuf<u64> oscr(want_ops ? (u64)per_round * WF_OPSMAX : 8, st); DBuf<u32> next(1, st);
        { KTimer kt(st, &ms[11]);
          for (u32 j0 = 0; j0 < nr; j0 += per_round) { u32 n = std::min(per_round, nr - j0); next.zero(); u32 blocks = (u32)std::min<u64>((u64)sm_count * 4, (n + WF_WARPS - 1) / WF_WARPS);
            k_wfa_fast<<<blocks, WF_WARPS * 32, WF_SMEM_BYTES, st>>>(d_ext.p, woff.p, words.p, hasamb.p, d_rest.p, j0, n, next.p, fslabs.p, d_out.p, adaptive, lmax); KERNEL_CHECK();
            k_wfa_bt<<<cdiv(n, 128), 128, 0, st>>>(d_ext.p, d_rest.p, j0, n, fslabs.p, oscr.p, d_out.p, ops_pool.p, ops_cur.p, ops_cap, want_ops, lmax); KERNEL_CHECK(); } }
        if (!use_reg) counters[14] = per_round; if (g_lap) (*g_lap)("wfa rounds");
        std::vector<WfaOut> o = d_out.to_host(nj); for (u32 j : rest) { if (o[j].status == 1) ids.push_back(j); else if (o[j].status != 0) throw std::runtime_error("WFA backtrace failed (fast kernel)"); else hw[j] = o[j]; } }
      if (g_lap) (*g_lap)("wfa d2h"); }
    u64 budget = 0; if (!ids.empty()) { size_t freeb = 0, totalb = 0; CUDA_CHECK(cudaMemGetInfo(&freeb, &totalb)); budget = (u64)(freeb * 0.6); }
    u64 slab_words = 1ull << 20;  // 4 MB per warp to start
    for (int round = 0; round < 6 && !ids.empty(); round++) {
      u32 n = (u32)ids.size(); u32 warps = (u32)std::min<u64>(std::min<u64>((u64)sm_count * 32, n), std::max<u64>(1, budget / (slab_words * 4))); warps = std::max(1u, (warps / 4) * 4); if (warps < 4) warps = 4;
      if ((u64)warps * slab_words * 4 > budget) throw std::runtime_error("WFA workspace does not fit in HBM for an alignment in this batch");
      DBuf<i32> slabs((u64)warps * slab_words, st); DBuf<u32> d_ids(n, st); d_ids.from_host(ids.data(), n); DBuf<u32> next(1, st); next.zero();
      { KTimer kt(st, &ms[10]); k_wfa<<<warps / 4, 128, 0, st>>>(d_jobs.p, d_ext.p, d_ids.p, n, next.p, woff.p, words.p, hasamb.p, qpacked, qamask, qboff, g2bit, g_off, slabs.p, slab_words, d_out.p, ops_pool.p, ops_cur.p, ops_cap, want_ops, adaptive); KERNEL_CHECK(); } counters[11] += n;
      std::vector<WfaOut> o = d_out.to_host(nj); std::vector<u32> again; for (u32 id : ids) { if (o[id].status == 1) again.push_back(id); else if (o[id].status != 0) throw std::runtime_error("WFA backtrace failed"); else hw[id] = o[id]; }
      ids.swap(again); slab_words *= 8;
    }
    if (!ids.empty()) throw std::runtime_error("WFA workspace exhausted after 6 rounds");
    if (want_ops) { u64 used = ops_cur.to_host()[0]; hops = ops_pool.to_host(used); }
}

// =====================================================================================================
// host orchestration of K4/K5 + finishing (lib-index-search.go:1834-2932)
// =====================================================================================================
struct StageTimer { cudaEvent_t ev[10]; cudaStream_t st; int n = 0; StageTimer(cudaStream_t s) : st(s) { for (auto& e : ev) cudaEventCreate(&e); } ~StageTimer() { for (auto& e : ev) cudaEventDestroy(e); }
  void mark() { cudaEventRecord(ev[n++], st); } float ms(int a, int b) { float f = 0; cudaEventElapsedTime(&f, ev[a], ev[b]); return f; } };

static void build_tree_tables(lmg_index* ix, QBatch& B, DBuf<u64>& tkeys, DBuf<u32>& tvals, DBuf<u32>& toff) {
  cudaStream_t st = ix->st; u64 n = B.total_k; toff.alloc(B.nq + 2, st); if (n == 0) { toff.zero(); tkeys.alloc(2, st); tvals.alloc(2, st); return; }
  DBuf<u32> flags(n + 1, st), pos(n + 1, st); k_tree_flags<<<cdiv((i64)n, 256), 256, 0, st>>>(B.qvals.p, n, flags.p); KERNEL_CHECK(); CUDA_CHECK(cudaMemsetAsync(flags.p + n, 0, 4, st));
  size_t tb = 0; cub::DeviceScan::ExclusiveSum(nullptr, tb, flags.p, pos.p, (int)(n + 1), st); cub::DeviceScan::ExclusiveSum(ix->tmp.get(tb), tb, flags.p, pos.p, (int)(n + 1), st); CUB_CHECK();
  u32 total; CUDA_CHECK(cudaMemcpyAsync(&total, pos.p + n, 4, cudaMemcpyDeviceToHost, st)); CUDA_CHECK(cudaStreamSynchronize(st));
  tkeys.alloc((u64)total + 2, st); tvals.alloc((u64)total + 2, st); k_tree_scatter<<<cdiv((i64)n, 256), 256, 0, st>>>(B.qkeys.p, B.qvals.p, pos.p, n, tkeys.p, tvals.p); KERNEL_CHECK();
  k_tree_offsets<<<cdiv(B.nq + 1, 128), 128, 0, st>>>(pos.p, B.koff.p, B.nq, n, total, toff.p); KERNEL_CHECK(); CUDA_CHECK(cudaStreamSynchronize(st));
}

struct HostHsp { i32 qb, qe, tb, te, aligned_q, tpo, max_ext; int job = -1; bool dead = false; i32 alen = 0, matched = 0, gaps = 0, score = 0, bitscore = 0; double evalue = 0, af = 0, pident = 0; std::string cigar, text; };   // text = qseq | sseq | align, alen bytes each (-a output)
struct HostCluster { u32 seg; u32 item; bool rc, variantA; int nseeds, iseq; std::vector<HostHsp> hsps; double sim = 0; bool has = false; };

struct lmg_results { std::vector<lmg_hsp> rows; std::string pool; std::vector<u32> row_genome; std::shared_ptr<const std::vector<std::vector<std::string>>> seq_ids; };   // sseqid of row i = (*seq_ids)[row_genome[i]][rows[i].seq_idx]; the host-side id table is shared with the index and outlives lmg_index_close

static void search_pipeline(lmg_index* ix, const lmg_params* prm, const u8* seqs, const u64* off, int nq, lmg_results& R, QBatch* staged, std::vector<lmg_pa>* pa_sink = nullptr) {
  cudaStream_t st = ix->st; const Image& I = ix->img; StageTimer T(st); T.mark();
  LapTimer lap; lap.lane = ix->lane_id; const bool dbgt = lap.on; struct LapScope { LapTimer* prev; LapScope(LapTimer* l) : prev(g_lap) { g_lap = l; } ~LapScope() { g_lap = prev; } } lapscope(&lap);
  QBatch Blocal; if (!staged) upload_queries(ix, seqs, off, nq, Blocal); QBatch& B = staged ? *staged : Blocal; nq = B.nq; T.mark();   // [0] h2d (zero when the queries were staged beforehand)
  lap("upload"); sketch_tables(ix, B); if (dbgt) cudaStreamSynchronize(st); lap("sketch tables"); Survivors SV; probe_survivors(ix, B, prm, SV, nullptr, nullptr); lap("capture+filter"); T.mark();   // [1] sketch (tables, capture, anchor-bit filter)
  if (prm->ext_len2 < 0 || prm->ext_len2 + 80 > 190) throw std::runtime_error("align-ext-len2 must be in [0, 110] for the GPU path (flank windows are held as 192-bit sets)");
  Anchors A; seed_probe(ix, B, prm, SV, A, false); SV.d.free(); if (dbgt) cudaStreamSynchronize(st); lap("probe+anchor sort"); T.mark();              // [2] probe
  Segments S; Chains Cn; chain_stage(ix, prm, A, S, Cn); A.hi.free(); A.lo.free(); lap("chain stage"); T.mark();                // [3] chain
  for (int i = 0; i < 16; i++) if (i != 8 && i != 9) ix->ms[i] = 0; ix->counters[11] = 0;
  auto finish_times = [&](int upto) { const int map_[6] = {0, 1, 2, 3, 4, 5}; (void)map_; CUDA_CHECK(cudaStreamSynchronize(st)); for (int i = 0; i < upto; i++) ix->ms[i] = T.ms(i, i + 1); ix->ms[7] = T.ms(0, upto); ix->counters[6] = B.total_bases; ix->counters[7] = (u64)B.nq; };
  if (Cn.n == 0) { T.mark(); finish_times(4); return; }
  // ---- windows (lib-index-search.go:1987-2051)
  const int K = I.k, extLen = prm->ext_len; std::vector<WinItem> items(Cn.n);
  const int HT = ix->host_threads();
  ix->pool.run(HT, HT, [&](int ci) { for (i64 c = (i64)Cn.n * ci / HT, cE = (i64)Cn.n * (ci + 1) / HT; c < cE; c++) { const ChainRec& r = Cn.h[c]; u64 key = S.h_key[r.seg]; WinItem w; w.q = (u32)(key >> 36); w.g = (u32)((key >> 2) & 0x3FFFFFFFFull); w.chain = (u32)c;
    i32 qlen = (i32)(B.h_off[w.q + 1] - B.h_off[w.q]); i32 qb = r.q0, tb = r.t0, qe = r.q1 + r.len1 - 1, te = r.t1 + r.len1 - 1; bool qrc = (r.flags1 >> 1) & 1, trc = r.flags1 & 1;
    bool rc = (r.nseeds == 1) ? (qrc != trc) : (tb > r.t1); i32 tBegin, tEnd;
    if (rc) { tBegin = r.t1 - extLen; if (tBegin < 0) tBegin = 0; tEnd = tb + r.len1 - 1 + extLen; } else { tBegin = tb - extLen; if (tBegin < 0) tBegin = 0; tEnd = te + extLen; }
    w.qBegin = qb - std::min(qb, extLen); w.qEnd = qe + std::min(qlen - qe - 1, extLen);
    i32 nb = (i32)I.h_nbases[w.g];
    i32 start = std::max(tBegin, 0), end = tEnd; if (end >= nb - 1) end = nb - 1; if (end < start) end = start; i32 sl = end - start + 1; if (sl < tEnd - tBegin + 1) tEnd -= tEnd - tBegin + 1 - sl;   // SubSeq3 clamp + :2045-2047
    w.tBegin = tBegin; w.tEnd = tEnd; w.W = sl; w.rc = rc; w.mp = 11 + (sl >= 1000000 ? 8 : sl >= 250000 ? 6 : sl >= 50000 ? 4 : sl >= 10000 ? 2 : 0); items[c] = w; } });
  u32 nit = Cn.n; DBuf<WinItem> d_items(nit, st); d_items.from_host(items.data(), nit);
  lap("windows host"); DBuf<u64> tkeys; DBuf<u32> tvals, toff; build_tree_tables(ix, B, tkeys, tvals, toff); lap("tree tables");
  // ---- K4 anchors: per-query prefix hash, one pass into capacity-bounded regions (exact rerun for the rare overflow), per-window sort
  std::vector<u32> htoff = toff.to_host(B.nq + 1); std::vector<u64> hhoff(B.nq + 1, 0); for (int q = 0; q < B.nq; q++) { u32 n = htoff[q + 1] - htoff[q]; u64 H = 0; if (n) { H = 16; while (H < 2ull * n) H <<= 1; } hhoff[q + 1] = hhoff[q] + H; }
  // queries -> item ranges (items are ordered by (query, genome)); per-query kernel when table + window fit in shared memory
  std::vector<u32> qbeg(B.nq, 0), qend(B.nq, 0), qlist, rest; { u32 i = 0; while (i < nit) { u32 q = items[i].q, j = i; while (j < nit && items[j].q == q) j++; qbeg[q] = i; qend[q] = j; i = j; } }
  u32 max_tn = 0; i32 maxW3 = 0; const size_t smem_cap3 = std::min<size_t>(ix->smem_optin - 48 * 1024, 178 * 1024);   // dynamic part; the 1,024-thread variant has 46 KB of static queues + Bloom filter
  for (int q = 0; q < B.nq; q++) if (qend[q] > qbeg[q]) { u32 tn = htoff[q + 1] - htoff[q]; i32 mw = 0; for (u32 i = qbeg[q]; i < qend[q]; i++) mw = std::max(mw, items[i].W); size_t need = (size_t)tn * 12 + 2 * ((size_t)(mw + 15) / 16 + 2) * 4 + 64;   // two window buffers: the 1,024-thread variant runs two groups
      if (need <= smem_cap3) { qlist.push_back(q); max_tn = std::max(max_tn, tn); maxW3 = std::max(maxW3, mw); } else for (u32 i = qbeg[q]; i < qend[q]; i++) rest.push_back(i); }
  if (rest.empty()) std::fill(hhoff.begin(), hhoff.end(), 0);   // the L2-resident hash index is only needed by the fallback kernel
  DBuf<u64> hoff(B.nq + 1, st); hoff.from_host(hhoff.data(), B.nq + 1); DBuf<u64> htab(hhoff[B.nq] + 2, st); htab.fill_ff();
  if (!rest.empty()) { u32 maxn = 0; for (int q = 0; q < B.nq; q++) maxn = std::max(maxn, htoff[q + 1] - htoff[q]); if (maxn) { dim3 g((unsigned)std::max(1, std::min(32, cdiv(maxn, 256))), B.nq); k_tree_hash_build<<<g, 256, 0, st>>>(tkeys.p, toff.p, hoff.p, B.nq, htab.p); KERNEL_CHECK(); } }
  std::vector<u32> hcap(nit); std::vector<u64> habeg(nit + 1, 0); i32 maxW = 0; for (u32 i = 0; i < nit; i++) { hcap[i] = (u32)std::min<i64>((i64)std::max(0, items[i].W - 30) + 256, 0x7fffffff); maxW = std::max(maxW, items[i].W); }
  size_t smemW = ((size_t)(maxW + 15) / 16 + 2) * 4; if (smemW > ix->smem_optin - 4096) throw std::runtime_error("target window too long for the shared-memory pseudo-alignment kernel");
  DBuf<u32> cnt(nit + 1, st), dcap(nit, st); DBuf<u64> abeg(nit + 1, st); DBuf<u64> lo0; std::vector<u32> hcnt; std::vector<u64> haend(nit); u64 NA = 0;
  max_tn = (max_tn + 3) & ~3u; const u32 win_words = (u32)((maxW3 + 15) / 16 + 2); size_t smem3 = (size_t)max_tn * 12 + 2 * (size_t)win_words * 4 + 64; if (smem3 > ix->smem_optin - 48 * 1024) { rest.clear(); qlist.clear(); for (u32 i = 0; i < nit; i++) rest.push_back(i); }   // mixed extremes: everything through the L2 kernel
  DBuf<u32> d_qlist(qlist.size() + 1, st), d_qbeg(B.nq + 1, st), d_qend(B.nq + 1, st), d_rest(rest.size() + 1, st); d_qlist.from_host(qlist.data(), qlist.size()); d_qbeg.from_host(qbeg.data(), B.nq); d_qend.from_host(qend.data(), B.nq); d_rest.from_host(rest.data(), rest.size());
  lap("k4 host prep");
  for (int pass = 0; pass < 2; pass++) {   // pass 1 only when some window produced more anchors than W+226: capacities become the exact counts
    for (u32 i = 0; i < nit; i++) habeg[i + 1] = habeg[i] + hcap[i];
    u64 slot_limit = 1ull << 31; if (const char* e = getenv("LMG_SLOT_LIMIT")) slot_limit = strtoull(e, nullptr, 10);   // test hook: forces the halving path on small batches
    if (habeg[nit] >= slot_limit) throw BatchTooLarge("more than 2^31 pseudo-alignment anchor slots in one batch; use smaller batches");
    dcap.from_host(hcap.data(), nit); abeg.from_host(habeg.data(), nit + 1); lo0.alloc(habeg[nit] + 2, st);
    { KTimer kt(st, &ix->ms[14]);
      if (!qlist.empty()) { if (smem3 > 64 * 1024) { k_pa_anchors3<1024><<<(u32)qlist.size(), 1024, smem3, st>>>(d_items.p, d_qlist.p, d_qbeg.p, d_qend.p, I.d_g2bit, I.d_g_off, tkeys.p, tvals.p, toff.p, abeg.p, dcap.p, cnt.p, lo0.p, max_tn, win_words); KERNEL_CHECK(); }
        else { k_pa_anchors3<256><<<(u32)qlist.size(), 256, smem3, st>>>(d_items.p, d_qlist.p, d_qbeg.p, d_qend.p, I.d_g2bit, I.d_g_off, tkeys.p, tvals.p, toff.p, abeg.p, dcap.p, cnt.p, lo0.p, max_tn, win_words); KERNEL_CHECK(); } }
      if (!rest.empty()) { k_pa_anchors2<<<(u32)rest.size(), 128, smemW, st>>>(d_items.p, d_rest.p, (u32)rest.size(), I.d_g2bit, I.d_g_off, tkeys.p, tvals.p, toff.p, htab.p, hoff.p, abeg.p, dcap.p, cnt.p, lo0.p, habeg[nit]); KERNEL_CHECK(); } }
    hcnt = cnt.to_host(nit);
    bool over = false; NA = 0; for (u32 i = 0; i < nit; i++) { if (hcnt[i] > hcap[i]) over = true; haend[i] = habeg[i] + hcnt[i]; NA += hcnt[i]; }
    if (!over) break; if (pass == 1) throw std::runtime_error("pseudo-alignment anchor capacity overflow after exact sizing"); for (u32 i = 0; i < nit; i++) hcap[i] = hcnt[i]; }
  lap("pa_anchors");
  std::vector<C2Rec> c2;
  if (NA > 0) {
    // compact layout for everything but the raw anchors: cbeg = exclusive prefix sum of the per-window counts
    DBuf<u64> aend(nit, st); aend.from_host(haend.data(), nit); std::vector<u64> hcbeg(nit + 1, 0); std::vector<u32> bins[5];
    for (u32 i = 0; i < nit; i++) { u32 c = hcnt[i]; hcbeg[i + 1] = hcbeg[i] + c; if (c) bins[c <= 64 ? 0 : c <= 256 ? 1 : c <= 1024 ? 2 : c <= 4096 ? 3 : 4].push_back(i); }
    DBuf<u64> cbeg(nit + 1, st); cbeg.from_host(hcbeg.data(), nit + 1); DBuf<u64> lo1(NA + 2, st);
    { std::vector<u32> all; u32 boff[6] = {0, 0, 0, 0, 0, 0}; for (int k3 = 0; k3 < 5; k3++) { all.insert(all.end(), bins[k3].begin(), bins[k3].end()); boff[k3 + 1] = (u32)all.size(); } DBuf<u32> dl(all.size() + 1, st); dl.from_host(all.data(), all.size());
      if (boff[1] > boff[0]) { k_pa_sort<64, 32><<<cdiv(boff[1] - boff[0], 4), 128, 4 * 64 * 8, st>>>(dl.p + boff[0], boff[1] - boff[0], abeg.p, cbeg.p, cnt.p, lo0.p, lo1.p); KERNEL_CHECK(); }
      if (boff[2] > boff[1]) { k_pa_sort<256, 128><<<boff[2] - boff[1], 128, 256 * 8, st>>>(dl.p + boff[1], boff[2] - boff[1], abeg.p, cbeg.p, cnt.p, lo0.p, lo1.p); KERNEL_CHECK(); }
      if (boff[3] > boff[2]) { k_pa_sort<1024, 256><<<boff[3] - boff[2], 256, 1024 * 8, st>>>(dl.p + boff[2], boff[3] - boff[2], abeg.p, cbeg.p, cnt.p, lo0.p, lo1.p); KERNEL_CHECK(); }
      if (boff[4] > boff[3]) { k_pa_sort<4096, 512><<<boff[4] - boff[3], 512, 4096 * 8, st>>>(dl.p + boff[3], boff[4] - boff[3], abeg.p, cbeg.p, cnt.p, lo0.p, lo1.p); KERNEL_CHECK(); }
      if (boff[5] > boff[4]) {   // rare: > 4096 anchors in one window
        std::vector<u64> hb(nit), he(nit); for (u32 i = 0; i < nit; i++) { hb[i] = hcbeg[i]; he[i] = hcnt[i] > 4096 ? hcbeg[i + 1] : hcbeg[i]; } DBuf<u64> sb(nit, st), se(nit, st), tmpc(NA + 2, st); sb.from_host(hb.data(), nit); se.from_host(he.data(), nit);
        k_pa_gather<<<boff[5] - boff[4], 256, 0, st>>>(dl.p + boff[4], boff[5] - boff[4], abeg.p, cbeg.p, cnt.p, lo0.p, tmpc.p); KERNEL_CHECK();
        size_t tb = 0; cub::DeviceSegmentedSort::SortKeys(nullptr, tb, tmpc.p, lo1.p, (int)NA, (int)nit, sb.p, se.p, st); cub::DeviceSegmentedSort::SortKeys(ix->tmp.get(tb), tb, tmpc.p, lo1.p, (int)NA, (int)nit, sb.p, se.p, st); CUB_CHECK(); CUDA_CHECK(cudaStreamSynchronize(st)); }
      if (dbgt) fprintf(stderr, "[lmg host] pa anchors %llu in %u windows (%.1f GB of slots): bins %zu / %zu / %zu / %zu / %zu\n", (unsigned long long)NA, nit, habeg[nit] * 8e-9, bins[0].size(), bins[1].size(), bins[2].size(), bins[3].size(), bins[4].size()); }
    if (dbgt) cudaStreamSynchronize(st); lap("k4 seg sort");
    Chain2Params P2; P2.max_gap = prm->align_max_gap; P2.min_score = (int)((double)prm->align_min_len * prm->min_pident / 100); P2.min_align_len = prm->align_min_len; P2.band_base = prm->align_band; P2.band_count = prm->align_band / 2; P2.k = K;
    DBuf<i32> sc(NA + 2, st); DBuf<u32> pred(NA + 2, st); DBuf<u64> stack(NA + 2, st); u32 capc = (u32)std::min<u64>(NA, 0x7fffffffu); DBuf<C2Rec> d_c2(capc, st); DBuf<u32> nout(1, st); nout.zero();
    { KTimer kt(st, &ix->ms[15]); k_pa_chain<<<cdiv((i64)nit * 32, 128), 128, 0, st>>>(lo1.p, abeg.p, aend.p, cbeg.p, nit, P2, lo0.p, sc.p, pred.p, stack.p, d_c2.p, nout.p, capc); KERNEL_CHECK(); }
    u32 nc2 = nout.to_host()[0]; if (nc2 > capc) throw std::runtime_error("chain2 list overflow"); c2 = d_c2.to_host(nc2);
    lap("pa_chain+d2h"); bucket_sort(c2, nit, [](const C2Rec& a) { return a.item; }, [](const C2Rec& a, const C2Rec& b) { if (a.qb != b.qb) return a.qb < b.qb; return a.ord < b.ord; });   // lib-seq_compare.go:501-508
  }
  lap("c2 bucket sort"); tkeys.free(); tvals.free(); T.mark();                                                                  // [4] pseudo-align
  if (pa_sink) { for (const C2Rec& r : c2) { const WinItem& w = items[r.item]; lmg_pa o; o.genome = I.genome_bgi[w.g]; o.query = w.q; o.t_begin = w.tBegin; o.t_end = w.tEnd; o.rc = (i32)w.rc; o.qb = r.qb; o.qe = r.qe; o.tb = r.tb; o.te = r.te; o.aligned_q = r.aligned_q; o.aligned_t = r.aligned_t; o.matched = r.matched; o.n_anchors = r.n_anchors; pa_sink->push_back(o); } T.mark(); finish_times(5); return; }
  if (c2.empty()) { T.mark(); finish_times(5); return; }
  // ---- contig mapping, clusters, jobs (lib-index-search.go:2083-2469) — host, sequential per (query, genome)
  std::vector<HostCluster> clusters; std::vector<HspJob> jobs; const int contigInterval = I.contig_interval;
  { // segments (query, genome) are independent: process them in parallel, concatenate in order
    std::vector<std::pair<u32, u32>> segs; { u32 c = 0; while (c < nit) { u32 seg = Cn.h[c].seg, e = c; while (e < nit && Cn.h[e].seg == seg) e++; segs.push_back({c, e}); c = e; } }
    std::vector<size_t> c2beg(nit + 1, c2.size()); { size_t x = c2.size(); for (i64 it = (i64)nit - 1; it >= 0; it--) { while (x > 0 && c2[x - 1].item >= (u32)it) x--; c2beg[it] = x; } }
    const int NT = HT; std::vector<std::vector<HostCluster>> segCl(NT); std::vector<std::vector<HspJob>> segJobs(NT);
    ix->pool.run(NT, HT, [&](int ti) { std::vector<HostCluster>& clusters_l = segCl[ti]; std::vector<HspJob>& jobs_l = segJobs[ti]; size_t s0 = segs.size() * ti / NT, s1 = segs.size() * (ti + 1) / NT; std::vector<std::array<int, 6>> keys;
     for (size_t si = s0; si < s1; si++) { u32 c = segs[si].first, cEnd = segs[si].second, seg = Cn.h[c].seg;
      keys.clear(); int iSeq = 0, iSeqPre = -1;
      for (u32 it = c; it < cEnd; it++) { const WinItem& w = items[it]; const auto& SS = I.seq_sizes[w.g]; const int numSeqs = (int)SS.size(); const bool rc = w.rc; const i32 tBegin = w.tBegin, tEnd = w.tEnd, tlenSeq = w.W;
        size_t c0 = c2beg[it], ci = c2beg[it + 1]; if (ci == c0) continue;
        iSeqPre = -1; HostCluster cur; cur.seg = seg; cur.item = it; cur.rc = rc; cur.nseeds = Cn.h[it].nseeds; cur.variantA = false; cur.iseq = 0;
        auto convert = [&](HostHsp& h, const C2Rec& r, int tpo, int iS) { h.qb = r.qb; h.qe = r.qe; h.aligned_q = r.aligned_q; h.tpo = tpo;
          if (rc) { h.tb = tBegin - tpo + (tlenSeq - r.te - 1); if (h.tb < 0) { h.qe += h.tb; h.aligned_q += h.tb; h.tb = 0; } h.te = tBegin - tpo + (tlenSeq - r.tb - 1); if (h.te > (i32)SS[iS] - 1) { h.qb += h.te - ((i32)SS[iS] - 1); h.te = (i32)SS[iS] - 1; } }
          else { h.tb = tBegin - tpo + r.tb; if (h.tb < 0) { h.qb -= h.tb; h.aligned_q += h.tb; h.tb = 0; } h.te = tBegin - tpo + r.te; if (h.te > (i32)SS[iS] - 1) { h.qe -= h.te - ((i32)SS[iS] - 1); h.te = (i32)SS[iS] - 1; } }
          h.max_ext = (i32)SS[iS] - 1 - h.te; };
        auto flush = [&](bool variantA, int iS) { if (cur.hsps.empty()) return; cur.variantA = variantA; cur.iseq = iS; i32 qlen = (i32)(B.h_off[w.q + 1] - B.h_off[w.q]);
          for (HostHsp& h : cur.hsps) { if (h.qb >= h.qe + 1) { h.dead = true; continue; } i32 start, end; if (rc) { start = tEnd - h.te - h.tpo; end = tEnd - h.tb - h.tpo + 1; } else { start = h.tpo + h.tb - tBegin; end = h.tpo + h.te - tBegin + 1; }
            if (start >= end) { h.dead = true; continue; } if (start < 0 || end > tlenSeq || h.qb < 0 || h.qe + 1 > qlen) { h.dead = true; continue; }
            int ext2 = prm->ext_len2; if (h.aligned_q > 1000000) ext2 += 80; else if (h.aligned_q > 250000) ext2 += 40; else if (h.aligned_q > 50000) ext2 += 20; else if (h.aligned_q > 10000) ext2 += 10;
            HspJob J; J.q = w.q; J.g = w.g; J.tBegin = tBegin; J.tEnd = tEnd; J.rc = rc; J.qlen = qlen; J.tlen = tlenSeq; J.start1 = h.qb; J.end1 = h.qe + 1; J.start2 = start; J.end2 = end; J.ext = ext2; J.tb_arg = h.tb; J.max_ext = h.max_ext; h.job = (int)jobs_l.size(); jobs_l.push_back(J); }
          clusters_l.push_back(cur); cur.hsps.clear(); };
        for (size_t x = c0; x < ci; x++) { const C2Rec& r = c2[x]; iSeq = 0; int tpoB = 0, tpoE = 0;
          if (numSeqs > 1) { iSeq = -1; int _b, _e; if (rc) { _b = tEnd - r.te + K; _e = tEnd - r.tb - K; } else { _b = tBegin + r.tb + K; _e = tBegin + r.te - K; }
            if (_b >= _e) { if (rc) { _b = tEnd - r.te; _e = tEnd - r.tb; } else { _b = tBegin + r.tb; _e = tBegin + r.te; } }
            for (int j = 0; j < numSeqs; j++) { int l = (int)SS[j]; tpoE += l - 1; if (_b + K >= tpoB && _e - K <= tpoE) { iSeq = j; break; } else if (_e < tpoB) { iSeq = -1; break; } tpoE += contigInterval + 1; tpoB = tpoE; }
            if (iSeq < 0) continue;
            if (iSeqPre >= 0 && iSeq != iSeqPre) { int iSeq0 = iSeq; iSeq = iSeqPre; HostHsp h; convert(h, r, tpoB, iSeq); flush(true, iSeq); iSeqPre = -1;
              std::array<int, 6> key{h.qb, h.qe, h.tb, h.te, iSeq, (int)rc}; if (std::find(keys.begin(), keys.end(), key) == keys.end()) { cur.hsps.push_back(h); keys.push_back(key); } iSeq = iSeq0; continue; } }
          iSeqPre = iSeq; HostHsp h; convert(h, r, tpoB, iSeq); std::array<int, 6> key{h.qb, h.qe, h.tb, h.te, iSeq, (int)rc}; if (std::find(keys.begin(), keys.end(), key) == keys.end()) { cur.hsps.push_back(h); keys.push_back(key); } }
        if (iSeq >= 0) flush(false, iSeq); } } });
    size_t ncl = 0, njb = 0; for (size_t si = 0; si < segCl.size(); si++) { ncl += segCl[si].size(); njb += segJobs[si].size(); } clusters.reserve(ncl); jobs.reserve(njb);
    for (size_t si = 0; si < segCl.size(); si++) { int jo = (int)jobs.size(); for (HostCluster& cl : segCl[si]) { for (HostHsp& h : cl.hsps) if (h.job >= 0) h.job += jo; clusters.push_back(std::move(cl)); } jobs.insert(jobs.end(), segJobs[si].begin(), segJobs[si].end()); } }
  lap("contig mapping");
  // ---- K5: extension + WFA
  u32 nj = (u32)jobs.size(); std::vector<ExtOut> hext; std::vector<WfaOut> hw; std::vector<u64> hops;
  if (nj) {
    DBuf<HspJob> d_jobs(nj, st); d_jobs.from_host(jobs.data(), nj); DBuf<u32> ecnt(2 * (u64)nj + 1, st); DBuf<i32> eres(4 * (u64)nj, st);
    { KTimer kt(st, &ix->ms[13]); k_extend_count<<<cdiv(2 * (i64)nj, 128), 128, 0, st>>>(d_jobs.p, nj, B.packed.p, B.boff.p, I.d_g2bit, I.d_g_off, ecnt.p); KERNEL_CHECK(); }
    std::vector<u32> hec = ecnt.to_host(2 * (u64)nj); std::vector<u64> hso(2 * (u64)nj + 1, 0); for (u64 i = 0; i < 2 * (u64)nj; i++) hso[i + 1] = hso[i] + hec[i]; u64 ES = hso.back();
    DBuf<u64> soff(2 * (u64)nj + 1, st); soff.from_host(hso.data(), hso.size()); DBuf<u16> anc(ES + 2, st), pj(ES + 2, st); DBuf<i16> esc(ES + 2, st);
    { KTimer kt(st, &ix->ms[13]); k_extend_run<<<cdiv(2 * (i64)nj * 32, 128), 128, 0, st>>>(d_jobs.p, nj, B.packed.p, B.boff.p, I.d_g2bit, I.d_g_off, soff.p, anc.p, esc.p, pj.p, eres.p); KERNEL_CHECK(); }
    lap("extend kernels"); DBuf<ExtOut> d_ext(nj, st); k_extend_final<<<cdiv(nj, 128), 128, 0, st>>>(d_jobs.p, nj, eres.p, d_ext.p); KERNEL_CHECK(); hext = d_ext.to_host(nj);
    wfa_run_all(st, ix->sm_count, d_jobs, d_ext, hext, nj, B.packed.p, B.amask.p, B.boff.p, I.d_g2bit, I.d_g_off, prm->output_seq, prm->wfa_adaptive, hw, hops, ix->counters, ix->ms, ix->total_mem, ix->active_lanes);
  }
  lap("K5 total");
  T.mark();                                                                                              // [5] extend + wfa
  // ---- finishing: scores, filters, ordering, rows (lib-index-search.go:2266-2357, :2701-2932; search.go:437-533)
  const double lnK = std::log(0.41), totalBases = (double)I.total_bases;
  ix->pool.run(HT, HT, [&](int ci) { for (i64 cli = (i64)clusters.size() * ci / HT, cE = (i64)clusters.size() * (ci + 1) / HT; cli < cE; cli++) { HostCluster& cl = clusters[cli]; i32 qlen = 0; const WinItem& w = items[cl.item]; qlen = (i32)(B.h_off[w.q + 1] - B.h_off[w.q]); double maxSim = 0; bool has = false;
    for (HostHsp& h : cl.hsps) { if (h.dead) continue; const WfaOut& o = hw[h.job]; const ExtOut& e = hext[h.job]; i32 ql = e.qe - e.qs, tl = e.te - e.ts;
      if (!o.has_m) { h.dead = true; continue; }   // trimOps == nil -> evalue MaxFloat64 > max_evalue
      int _s = o.bscore; if (_s & 1) _s--; double bs = (0.625 * (double)_s - lnK) / M_LN2; h.score = o.bscore; h.bitscore = (int)bs; h.evalue = totalBases * std::pow(2, -bs) * (double)ql; if (h.evalue > prm->max_evalue) { h.dead = true; continue; }
      h.qb -= e.s1; h.qe += e.e1; h.qb = h.qb + o.qbegin - 1; h.qe = h.qe - (ql - o.qend);
      if (cl.rc) { h.tb -= e.e2; h.te += e.s2; h.tb = h.tb + (tl - o.tend); h.te = cl.variantA ? (h.te - o.tbegin - 1) : (h.te - (o.tbegin - 1)); } else { h.tb -= e.s2; h.te += e.e2; h.tb = h.tb + o.tbegin - 1; h.te = h.te - (tl - o.tend); }
      h.aligned_q = h.qe - h.qb + 1; h.alen = o.alen; h.matched = o.matches; h.gaps = o.gaps; h.af = (double)h.aligned_q / (double)qlen * 100; if (h.af > 100) h.af = 100; h.pident = (double)h.matched / (double)o.alen * 100;
      if (h.af < prm->min_qcov_hsp || h.pident < prm->min_pident) { h.dead = true; continue; }
      if (prm->output_seq) { // ops were written last-op-first; trim to first M..last M; swap I/D for SAM (:2331-2338)
        std::vector<u64> ops(hops.begin() + o.ops_off, hops.begin() + o.ops_off + o.ops_n); std::reverse(ops.begin(), ops.end()); int a = -1, b = -1; for (size_t i = 0; i < ops.size(); i++) if ((ops[i] >> 32) == 'M') { if (a < 0) a = (int)i; b = (int)i; }
        // alignment text (cigar.AlignmentText(&_qseq, &_tseq, true), :2342): first M .. last M over the query bytes and the target strand that was aligned
        const u8* qs = B.h_ascii.data() + B.h_off[w.q]; const u8* g2h = I.h_g2bit.data() + I.h_g_off[w.g]; i32 qi = h.qb; const i64 t0 = (i64)h.tpo + h.tb, t1 = (i64)h.tpo + h.te; i64 ti = 0; std::string qt, tt, at; qt.reserve(h.alen); tt.reserve(h.alen); at.reserve(h.alen);
        auto qch = [&](i32 x) -> char { char c = (char)qs[x]; return (c >= 'a' && c <= 'z') ? (char)(c - 32) : c; };   // the reference upper-cases the query on input (search.go:582-587)
        auto tch = [&](i64 x) -> char { i64 pos = cl.rc ? t1 - x : t0 + x; u32 bb = (g2h[pos >> 2] >> (6 - 2 * (pos & 3))) & 3; return "ACGT"[cl.rc ? 3 - bb : bb]; };
        for (int i = a; i >= 0 && i <= b; i++) { char c = (char)(ops[i] >> 32); u32 n = (u32)(ops[i] & 0xffffffffu);
          for (u32 x = 0; x < n; x++) { if (c == 'M' || c == 'X') { qt.push_back(qch(qi++)); tt.push_back(tch(ti++)); at.push_back(c == 'M' ? '|' : ' '); } else if (c == 'I') { qt.push_back('-'); tt.push_back(tch(ti++)); at.push_back(' '); } else { qt.push_back(qch(qi++)); tt.push_back('-'); at.push_back(' '); } }
          if (c == 'D') c = 'I'; else if (c == 'I') c = 'D'; h.cigar += std::to_string(n); h.cigar.push_back(c); }
        h.text = qt + tt + at; }
      double sim = (double)h.bitscore * h.pident; if (sim > maxSim) maxSim = sim; has = true; }
    cl.has = has; cl.sim = maxSim; } });
  lap("score clusters");
  // group clusters per segment -> genomes -> queries; whole queries are independent, so static chunks of queries run in parallel
  struct GenomeOut { u32 seg; std::vector<const HostCluster*> sds; double af; };
  const int NTF = HT; std::vector<size_t> cut(NTF + 1, clusters.size()); cut[0] = 0;
  for (int t = 1; t < NTF; t++) { size_t x = clusters.size() * t / NTF; while (x > 0 && x < clusters.size() && (u32)(S.h_key[clusters[x].seg] >> 36) == (u32)(S.h_key[clusters[x - 1].seg] >> 36)) x++; cut[t] = std::max(x, cut[t - 1]); }
  std::vector<std::vector<lmg_hsp>> trows(NTF); std::vector<std::string> tpool(NTF); std::vector<std::vector<u32>> trg(NTF);
  ix->pool.run(NTF, HT, [&](int ti) {
    std::vector<GenomeOut> gouts; std::vector<lmg_hsp>& rows = trows[ti]; std::string& pool = tpool[ti]; std::vector<u32>& rg = trg[ti];
    auto seg_q = [&](u32 seg) { return (u32)(S.h_key[seg] >> 36); }; auto seg_g = [&](u32 seg) { return (u32)((S.h_key[seg] >> 2) & 0x3FFFFFFFFull); };
    // query coverage of a genome = union of its HSPs' query intervals (coverageLen lib-seq_compare.go:270-308); false = below -Q
    auto coverage = [&](GenomeOut& g) { u32 q = seg_q(g.seg); i32 qlen = (i32)(B.h_off[q + 1] - B.h_off[q]); std::vector<std::array<int, 2>> reg; for (auto* sd : g.sds) for (const HostHsp& h : sd->hsps) if (!h.dead) reg.push_back({h.qb, h.qe});
      int cov = 0; if (reg.size() == 1) cov = reg[0][1] - reg[0][0] + 1; else if (!reg.empty()) { std::stable_sort(reg.begin(), reg.end(), [](const std::array<int, 2>& a, const std::array<int, 2>& b) { return a[0] < b[0]; }); int s0 = reg[0][0], e0 = reg[0][1]; for (size_t i = 1; i < reg.size(); i++) { if (reg[i][0] > e0) { cov += e0 - s0 + 1; s0 = reg[i][0]; e0 = reg[i][1]; continue; } if (reg[i][1] <= e0) continue; e0 = reg[i][1]; } cov += e0 - s0 + 1; }
      g.af = (double)cov / (double)qlen * 100; if (g.af > 100) g.af = 100; return g.af >= prm->min_qcov_genome; };
    { size_t x = cut[ti]; while (x < cut[ti + 1]) { u32 seg = clusters[x].seg; GenomeOut g; g.seg = seg; g.af = 0; size_t y = x; while (y < cut[ti + 1] && clusters[y].seg == seg) { if (clusters[y].has) g.sds.push_back(&clusters[y]); y++; } x = y; if (g.sds.empty()) continue;
        if (!I.has_chunks) { if (!coverage(g)) continue; std::stable_sort(g.sds.begin(), g.sds.end(), [](const HostCluster* a, const HostCluster* b) { return a->sim > b->sim; }); }   // with split genomes in the index both wait for the merge (:2701)
        gouts.push_back(std::move(g)); } }
    size_t x = 0; while (x < gouts.size()) { u32 q = seg_q(gouts[x].seg); size_t y = x; while (y < gouts.size() && seg_q(gouts[y].seg) == q) y++;
      std::vector<GenomeOut*> rs; for (size_t z = x; z < y; z++) rs.push_back(&gouts[z]);
      auto bgi_of = [&](const GenomeOut* g) { return I.genome_bgi[seg_g(g->seg)]; };
      if (q < Cn.topn_sorted.size() && Cn.topn_sorted[q]) std::stable_sort(rs.begin(), rs.end(), [&](GenomeOut* a, GenomeOut* b) { return Cn.seg_score[a->seg] > Cn.seg_score[b->seg]; });   // the order the top-N selection left (:1780-1805)
      std::stable_sort(rs.begin(), rs.end(), [&](GenomeOut* a, GenomeOut* b) { return (bgi_of(a) & 131071) < (bgi_of(b) & 131071); });    // :1848-1853
      if (I.has_chunks) {   // merge the chunks of a split genome into the first one (:2797-2852), then coverage, -Q and cluster order on the merged result (:2856-2897)
        std::vector<GenomeOut*> kept; std::vector<std::pair<u32, GenomeOut*>> first;
        for (GenomeOut* g : rs) { u32 grp = I.chunk_group[seg_g(g->seg)]; GenomeOut* tgt = nullptr; if (grp != 0xFFFFFFFFu) { for (auto& f : first) if (f.first == grp) tgt = f.second; if (!tgt) first.push_back({grp, g}); }
          if (tgt) { tgt->sds.insert(tgt->sds.end(), g->sds.begin(), g->sds.end()); g->sds.clear(); } else kept.push_back(g); }
        rs.clear(); for (GenomeOut* g : kept) { if (!coverage(*g)) continue; std::stable_sort(g->sds.begin(), g->sds.end(), [](const HostCluster* a, const HostCluster* b) { return a->sim > b->sim; }); rs.push_back(g); } }
      std::stable_sort(rs.begin(), rs.end(), [](GenomeOut* a, GenomeOut* b) { return a->sds[0]->sim > b->sds[0]->sim; });                  // :2919-2921
      for (GenomeOut* g : rs) { u32 gd = seg_g(g->seg);
        std::vector<const HostCluster*> ord; std::vector<char> used(g->sds.size(), 0);   // SortBySeqID :1042-1096 (compares the sequence IDs)
        auto sid = [&](const HostCluster* c) -> const std::string& { return I.seq_ids[seg_g(c->seg)][c->iseq]; };
        for (size_t i = 0; i < g->sds.size(); i++) { if (used[i]) continue; for (size_t j = i; j < g->sds.size(); j++) if (!used[j] && (g->sds[j] == g->sds[i] || sid(g->sds[j]) == sid(g->sds[i]))) { used[j] = 1; ord.push_back(g->sds[j]); } }
        int cls = 1, j = 1; for (const HostCluster* sd : ord) { const u32 sg = seg_g(sd->seg);   // sg: the chunk this cluster was found in (== gd unless chunks were merged)
          for (const HostHsp& h : sd->hsps) { if (h.dead) continue; lmg_hsp r; memset(&r, 0, sizeof r); r.query = q; r.hits = (u32)rs.size(); r.genome = I.genome_bgi[gd]; r.seq_idx = sd->iseq; r.n_seqs = (u32)I.seq_ids[sg].size(); r.chunk_idx = I.chunk_idx[sg]; r.n_chunks = I.chunk_n[sg]; r.seq_len = (i32)I.seq_sizes[sg][sd->iseq];
            r.cls = cls; r.hsp = j; r.qb = h.qb; r.qe = h.qe; r.tb = h.tb; r.te = h.te; r.rc = sd->rc; r.alen = h.alen; r.matches = h.matched; r.gaps = h.gaps; r.score = h.score; r.bitscore = h.bitscore; r.evalue = h.evalue; r.qcov_hsp = h.af; r.pident = h.pident; r.qcov_gnm = g->af;
            r.cigar_off = pool.size(); r.cigar_len = (u32)h.cigar.size(); pool += h.cigar; pool += h.text; rows.push_back(r); rg.push_back(sg); j++; } cls++; } }
      x = y; } });
  { size_t nr = 0; for (auto& v : trows) nr += v.size(); R.rows.reserve(nr); R.row_genome.reserve(nr); for (int ti = 0; ti < NTF; ti++) { u64 po = R.pool.size(); for (lmg_hsp& r : trows[ti]) { r.cigar_off += po; R.rows.push_back(r); } R.pool += tpool[ti]; R.row_genome.insert(R.row_genome.end(), trg[ti].begin(), trg[ti].end()); } R.seq_ids = I.seq_ids_p; }
  lap("group+rows");
  T.mark(); finish_times(7);                                                                             // [6] finish (host)
}

// =====================================================================================================
// C ABI (part 1)
// =====================================================================================================
extern "C" {

void lmg_default_params(lmg_params* p) { p->min_prefix = 15; p->min_single_prefix = 17; p->top_n_genomes = 0; p->top_n_chains = 0; p->max_gap = 50; p->max_distance = 1000; p->ext_len = 1000; p->ext_len2 = 50;
  p->min_qcov_genome = 0; p->max_evalue = 10; p->align_max_gap = 20; p->align_min_len = 50; p->align_band = 100; p->output_seq = 0; p->min_pident = 70; p->min_qcov_hsp = 0; p->wfa_adaptive = 1; p->lanes = 0; }
const char* lmg_last_error(void) { return g_err.c_str(); }

static lmg_index* make_ctx(Image* im, bool owner, int device) {
  lmg_index* ix = new lmg_index(im, owner); CUDA_CHECK(cudaStreamCreateWithFlags(&ix->st, cudaStreamNonBlocking)); ix->tmp.st = ix->st;
  { int lo_p = 0, hi_p = 0; CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p)); CUDA_CHECK(cudaStreamCreateWithPriority(&ix->st_hi, cudaStreamNonBlocking, hi_p)); CUDA_CHECK(cudaEventCreateWithFlags(&ix->ev_hi[0], cudaEventDisableTiming)); CUDA_CHECK(cudaEventCreateWithFlags(&ix->ev_hi[1], cudaEventDisableTiming)); }
  cudaDeviceProp pr; CUDA_CHECK(cudaGetDeviceProperties(&pr, device)); ix->sm_count = pr.multiProcessorCount; ix->smem_optin = (u32)pr.sharedMemPerBlockOptin; if (pr.major < 9) ix->use_tma = 0;
  if (getenv("LMG_NO_TMA")) ix->use_tma = 0; ix->total_mem = pr.totalGlobalMem;
  if (const char* e = getenv("LMG_L2_FETCH")) { const int g = atoi(e); if (g == 32 || g == 64 || g == 128) CUDA_CHECK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)g)); }   /* experiment: DRAM bytes fetched per missing 32-byte sector (a device-wide hint) */
  // dynamic shared memory ceilings are per function and device-global: raise them once to the opt-in limit so concurrent lanes never race on them
  auto raise = [&](const void* f) { cudaFuncAttributes fa; CUDA_CHECK(cudaFuncGetAttributes(&fa, f)); CUDA_CHECK(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(ix->smem_optin - fa.sharedSizeBytes))); };
  if (owner) { raise((const void*)k_capture); raise((const void*)k_capture2<true>); raise((const void*)k_capture2<false>); raise((const void*)k_pa_anchors2); raise((const void*)k_pa_anchors3<256>); raise((const void*)k_pa_anchors3<1024>); raise((const void*)k_pa_sort<4096, 512>); }
  return ix;
}
static void free_ctx(lmg_index* ix) { cudaStreamSynchronize(ix->st); if (ix->st_hi) { cudaStreamSynchronize(ix->st_hi); cudaStreamDestroy(ix->st_hi); } for (auto& e : ix->ev_hi) if (e) cudaEventDestroy(e); ix->arena.release(); if (ix->tmp.p) { cudaFree(ix->tmp.p); ix->tmp.p = nullptr; } for (auto& e : ix->kev) if (e) cudaEventDestroy(e); cudaStreamDestroy(ix->st); }
static lmg_index* lane_ctx(lmg_index* ix, int l) { if (l == 0) return ix; while ((int)ix->lanes.size() < l) { ix->lanes.push_back(make_ctx(ix->imgp, false, ix->img.device)); ix->lanes.back()->lane_id = (int)ix->lanes.size(); } return ix->lanes[l - 1]; }
// number of concurrent sub-batches: explicit (params.lanes / LMG_LANES) or up to 6 for batches big enough to amortise the split
// (10,000 x 1-kb bench, ms per batch: 1 lane 124-140, 3 lanes 102, 4 lanes 88, 6 lanes 81-84, 8 lanes 85, 12 lanes 87; exclusive GPU phases were slower)
static int pick_lanes(const lmg_params* p, int nq, u64 bases) { int L = (p && p->lanes > 0) ? p->lanes : 0; if (!L) { const char* e = getenv("LMG_LANES"); if (e) L = atoi(e); } bool forced = L > 0; if (!L) L = 6; L = std::max(1, std::min(L, 16));
  if (!forced) while (L > 1 && (nq < 800 * L || bases < (u64)800000 * L)) L--; return std::max(1, std::min(L, std::max(nq, 1))); }
static std::vector<int> lane_cuts(const u64* off, int nq, int L) { std::vector<int> cut(L + 1, nq); cut[0] = 0; u64 tot = off[nq] - off[0]; int q = 0; for (int l = 1; l < L; l++) { u64 want = off[0] + tot * l / L; while (q < nq && off[q] < want) q++; cut[l] = std::max(q, cut[l - 1]); } return cut; }

int lmg_index_open(const char* dir, int device, int shard, int n_shards, lmg_index** out) {
  try { int ndev = 0; CUDA_CHECK(cudaGetDeviceCount(&ndev)); if (ndev == 0) throw std::runtime_error("no CUDA device: the LexicMap GPU path has no CPU fallback");
    Image* im = new Image; lmg_index* ix = nullptr; try { im->load(dir, device, shard, std::max(1, n_shards)); ix = make_ctx(im, true, device); } catch (...) { if (!ix) { im->release(); delete im; } throw; }
    cudaMemPool_t pool; CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, device)); u64 thr = ~0ull; CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    *out = ix; return 0; } catch (std::exception& e) { g_err = e.what(); return -1; }
}
int lmg_index_info(const lmg_index* ix, lmg_info* o) { const Image& I = ix->img; o->k = I.k; o->masks = I.m; o->chunks = I.info.chunks; o->partitions = I.info.partitions; o->genomes = I.G; o->genome_batches = I.info.genome_batches;
  o->contig_interval = I.contig_interval; o->mask_prefix = I.mask_prefix; o->anchor_prefix = I.anchor_prefix; o->input_bases = I.total_bases; o->seed_keys = I.E; o->seed_values = I.V; o->image_bytes = I.bytes; return 0; }
int lmg_genome_name(const lmg_index* ix, uint64_t genome, const char** name) { auto it = ix->img.bgi2dense.find(genome); if (it == ix->img.bgi2dense.end()) { *name = ""; return -1; } *name = ix->img.genome_names[it->second].c_str(); return 0; }
void lmg_index_close(lmg_index* ix) { if (!ix) return; cudaSetDevice(ix->img.device); for (lmg_index* l : ix->lanes) { free_ctx(l); delete l; } free_ctx(ix); ix->img.release(); Image* im = ix->imgp; delete ix; delete im; }
void lmg_free(void* p) { free(p); }
int lmg_last_timing(const lmg_index* ix, double* ms16, uint64_t* c16) { for (int i = 0; i < 16; i++) { if (ms16) ms16[i] = ix->ms[i]; if (c16) c16[i] = ix->counters[i]; } if (c16) c16[15] = g_launches; return 0; }

int lmg_mask_batch(lmg_index* ix, const uint8_t* seqs, const uint64_t* off, int32_t n, uint64_t* kmers, uint32_t* nlocs, uint32_t* minloc, uint64_t* suf, uint64_t suf_cap, uint64_t* n_suf) {
  try { std::lock_guard<std::mutex> lk(ix->mu); CUDA_CHECK(cudaSetDevice(ix->img.device)); ArenaReset ar_(ix); QBatch B; upload_queries(ix, seqs, off, n, B); sketch_tables(ix, B); CapBufs cap; DBuf<u32> owner; { lmg_params dp; lmg_default_params(&dp); Survivors SV; probe_survivors(ix, B, &dp, SV, &cap, &owner); }
    const int m = ix->img.m, k = ix->img.k; std::vector<Capture> hc((u64)n * m); { auto a = cap.kmer.to_host(); auto b = cap.lo.to_host(); auto c = cap.n.to_host(); auto d = cap.smask.to_host(); for (size_t i = 0; i < hc.size(); i++) { hc[i].kmer = a[i]; hc[i].lo = b[i]; hc[i].n = c[i]; hc[i].smask = d[i]; } } auto hv = B.qvals.to_host(); auto ho = owner.to_host(); u64 ns = 0; std::vector<std::array<u64, 4>> trip;
    for (int q = 0; q < n; q++) { trip.clear();
      for (int i = 0; i < m; i++) { const Capture& c = hc[(u64)q * m + i]; u64 o = (u64)q * m + i; kmers[o] = c.kmer; nlocs[o] = c.kmer ? c.n : 0; u32 mn = 0xffffffffu; if (c.kmer) for (u32 t = 0; t < c.n; t++) mn = std::min(mn, hv[B.h_koff[q] + c.lo + t] & 0x7fffffffu); minloc[o] = c.kmer ? mn : 0;
        if (c.kmer && ho[B.h_koff[q] + c.lo] == (u32)i) trip.push_back({(u64)q, (u64)c.smask, (u64)i, kmer_reverse62(c.kmer, k)}); }
      std::sort(trip.begin(), trip.end()); for (auto& t : trip) { if (ns < suf_cap) memcpy(suf + 4 * ns, t.data(), 32); ns++; } }
    *n_suf = ns; return 0; } catch (std::exception& e) { g_err = e.what(); return -1; }
}

int lmg_anchor_batch(lmg_index* ix, const lmg_params* p, const uint8_t* seqs, const uint64_t* off, int32_t n, lmg_anchor** out, uint64_t* n_out) {
  try { std::lock_guard<std::mutex> lk(ix->mu); CUDA_CHECK(cudaSetDevice(ix->img.device)); ArenaReset ar_(ix); QBatch B; upload_queries(ix, seqs, off, n, B); sketch_tables(ix, B); Survivors SV; probe_survivors(ix, B, p, SV, nullptr, nullptr);
    Anchors A; seed_probe(ix, B, p, SV, A, true); auto hi = A.hi.to_host(A.n), lo = A.lo.to_host(A.n); lmg_anchor* o = (lmg_anchor*)malloc(sizeof(lmg_anchor) * (A.n + 1));
    for (u64 i = 0; i < A.n; i++) { lmg_anchor& a = o[i]; u32 g = (u32)((hi[i] >> 2) & 0x3FFFFFFFFull); a.genome = ix->img.genome_bgi[g]; a.query = (u32)(hi[i] >> 36); a.qbegin = (i32)(lo[i] >> 36); a.len = (u8)(63 - ((lo[i] >> 30) & 63)); a.tbegin = (i32)((lo[i] >> 2) & 0x0FFFFFFF); a.qrc = (lo[i] >> 1) & 1; a.trc = lo[i] & 1; a.pad = 0; }
    *out = o; *n_out = A.n; return 0; } catch (std::exception& e) { g_err = e.what(); return -1; }
}


int lmg_chain_batch(lmg_index* ix, const lmg_params* p, const uint8_t* seqs, const uint64_t* off, int32_t n, lmg_chain** out, uint64_t* n_out) {
  try { std::lock_guard<std::mutex> lk(ix->mu); CUDA_CHECK(cudaSetDevice(ix->img.device)); ArenaReset ar_(ix); QBatch B; upload_queries(ix, seqs, off, n, B); sketch_tables(ix, B); Survivors SV; probe_survivors(ix, B, p, SV, nullptr, nullptr);
    Anchors A; seed_probe(ix, B, p, SV, A, false); Segments S; Chains C; chain_stage(ix, p, A, S, C);
    lmg_chain* o = (lmg_chain*)malloc(sizeof(lmg_chain) * (C.n + 1));
    for (u32 i = 0; i < C.n; i++) { const ChainRec& r = C.h[i]; lmg_chain& c = o[i]; u64 key = S.h_key[r.seg]; c.query = (u32)(key >> 36); c.genome = ix->img.genome_bgi[(u32)((key >> 2) & 0x3FFFFFFFFull)]; c.score = r.score; c.n_seeds = r.nseeds;
      c.q0 = r.q0; c.t0 = r.t0; c.len0 = r.len0; c.q1 = r.q1; c.t1 = r.t1; c.len1 = r.len1; bool qrc = (r.flags1 >> 1) & 1, trc = r.flags1 & 1; c.rc = (r.nseeds == 1) ? (qrc != trc) : (r.t0 > r.t1); }
    *out = o; *n_out = C.n; return 0; } catch (std::exception& e) { g_err = e.what(); return -1; }
}

int lmg_pseudoalign_batch(lmg_index* ix, const lmg_params* p, const uint8_t* seqs, const uint64_t* off, int32_t n, lmg_pa** out, uint64_t* n_out) {
  try { std::lock_guard<std::mutex> lk(ix->mu); CUDA_CHECK(cudaSetDevice(ix->img.device)); if (ix->img.synth_per) throw std::runtime_error("synthetic seeds-only index"); std::vector<lmg_pa> v; if (n > 0) { ArenaReset ar_(ix); ix->active_lanes = 1; lmg_results R; search_pipeline(ix, p, seqs, off, n, R, nullptr, &v); }
    lmg_pa* o = (lmg_pa*)malloc(sizeof(lmg_pa) * (v.size() + 1)); if (!v.empty()) memcpy(o, v.data(), sizeof(lmg_pa) * v.size()); *out = o; *n_out = v.size(); return 0; } catch (std::exception& e) { g_err = e.what(); cudaGetLastError(); return -1; }
}

struct lmg_queries { std::vector<QBatch> parts; std::vector<int> cut; };
// Runs the pipeline over L sub-batches concurrently (one host thread, stream and arena per lane) and concatenates the rows in query order.
// Queries are independent all the way through the reference path (search.go:437-533 handles one query at a time), so the split is exact.
// one lane's share of a call; halves the range (recursively) when an intermediate list does not fit
static void search_range(lmg_index* lx, const lmg_params* p, const u8* seqs, const u64* off, int n, lmg_results& R, QBatch* staged) {
  try { ArenaReset ar_(lx); search_pipeline(lx, p, seqs, off, n, R, staged); return; }
  catch (BatchTooLarge& e) { if (n < 2) throw std::runtime_error(std::string(e.what()) + " (a single query)"); }
  catch (ArenaExhausted& e) { cudaGetLastError(); if (n < 2) throw std::runtime_error(std::string(e.what()) + " (a single query)"); }   // the sub-batch's buffers do not fit in HBM next to the image: halve it
  if (staged) { seqs = staged->h_ascii.data(); off = staged->h_off.data(); }   // the host copy made at upload time
  R = lmg_results(); const u64 mid = off[0] + (off[n] - off[0]) / 2; int h = 1; while (h < n - 1 && off[h] < mid) h++;
  lmg_results R2; double ms[16]; u64 cn[16]; search_range(lx, p, seqs, off, h, R, nullptr); for (int i = 0; i < 16; i++) { ms[i] = lx->ms[i]; cn[i] = lx->counters[i]; }
  search_range(lx, p, seqs, off + h, n - h, R2, nullptr); for (int i = 0; i < 16; i++) if (i != 12) { lx->ms[i] += ms[i]; if (i != 14) lx->counters[i] += cn[i]; }
  u64 po = R.pool.size(); for (lmg_hsp& r : R2.rows) { r.query += (u32)h; r.cigar_off += po; R.rows.push_back(r); } R.pool += R2.pool; R.row_genome.insert(R.row_genome.end(), R2.row_genome.begin(), R2.row_genome.end()); R.seq_ids = lx->img.seq_ids_p;
}

static void search_lanes(lmg_index* ix, const lmg_params* p, const u8* seqs, const u64* off, int nq, lmg_results& R, lmg_queries* staged) {
  auto w0 = std::chrono::steady_clock::now(); const int dev = ix->img.device; CUDA_CHECK(cudaSetDevice(dev));
  if (ix->img.synth_per) throw std::runtime_error("this index is a synthetic seeds-only image (lmg_index_synth): only lmg_probe_bench runs on it");
  if (ix->img.n_shards > 1 && p->top_n_genomes > 0) throw std::runtime_error("--top-n-genomes cannot be applied inside one genome shard (the top N are chosen over all genomes): search the shards without it and select after merging");
  std::vector<int> cut = staged ? staged->cut : lane_cuts(off, nq, pick_lanes(p, nq, off[nq] - off[0])); const int L = (int)cut.size() - 1;
  if (L == 1) { ix->active_lanes = 1; search_range(ix, p, seqs, off, nq, R, staged ? &staged->parts[0] : nullptr); }
  else {
    std::vector<lmg_results> Rl(L); std::vector<std::string> err(L); std::vector<lmg_index*> lx(L); for (int l = 0; l < L; l++) { lx[l] = lane_ctx(ix, l); lx[l]->active_lanes = L; }
    auto work = [&](int l) { try { CUDA_CHECK(cudaSetDevice(dev)); int n = cut[l + 1] - cut[l]; if (n > 0) search_range(lx[l], p, seqs, staged ? nullptr : off + cut[l], n, Rl[l], staged ? &staged->parts[l] : nullptr); else { for (double& m : lx[l]->ms) m = 0; for (u64& c : lx[l]->counters) c = 0; } }
      catch (std::exception& e) { err[l] = e.what(); if (err[l].empty()) err[l] = "error"; cudaGetLastError(); } };
    std::vector<std::thread> th; for (int l = 1; l < L; l++) th.emplace_back(work, l); work(0); for (auto& t : th) t.join();
    for (int l = 0; l < L; l++) if (!err[l].empty()) throw std::runtime_error(err[l]);
    size_t nr = 0, np = 0; for (auto& r : Rl) { nr += r.rows.size(); np += r.pool.size(); } R.rows.reserve(nr); R.row_genome.reserve(nr); R.pool.reserve(np); R.seq_ids = ix->img.seq_ids_p;
    for (int l = 0; l < L; l++) { u64 po = R.pool.size(); for (lmg_hsp& r : Rl[l].rows) { r.query += (u32)cut[l]; r.cigar_off += po; R.rows.push_back(r); } R.pool += Rl[l].pool; R.row_genome.insert(R.row_genome.end(), Rl[l].row_genome.begin(), Rl[l].row_genome.end()); }
    // timers and counters: summed over the lanes (the lanes overlap, so stage sums exceed the wall time in ms[7])
    double msum[16] = {0}; u64 csum[16] = {0}; for (int l = 0; l < L; l++) for (int i = 0; i < 16; i++) { msum[i] += lx[l]->ms[i]; csum[i] += lx[l]->counters[i]; }
    csum[14] = lx[0]->counters[14]; for (int i = 0; i < 16; i++) { ix->ms[i] = msum[i]; ix->counters[i] = csum[i]; }
    ix->ms[7] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
  }
  ix->ms[12] = (double)L; ix->ms[9] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
}
int lmg_search_batch(lmg_index* ix, const lmg_params* p, const uint8_t* seqs, const uint64_t* off, int32_t n, lmg_results** out) {
  try { std::lock_guard<std::mutex> lk(ix->mu); lmg_results* R = new lmg_results; if (n <= 0) { R->seq_ids = ix->img.seq_ids_p; *out = R; return 0; }   /* empty batch: no rows */ try { search_lanes(ix, p, seqs, off, n, *R, nullptr); } catch (...) { delete R; throw; } *out = R; return 0; }
  catch (std::exception& e) { g_err = e.what(); cudaGetLastError(); return -1; }
}
int lmg_results_rows(const lmg_results* r, const lmg_hsp** rows, uint64_t* n_rows, const char** pool, uint64_t* pool_len) { *rows = r->rows.data(); *n_rows = r->rows.size(); if (pool) *pool = r->pool.data(); if (pool_len) *pool_len = r->pool.size(); return 0; }
int lmg_results_seq_id(const lmg_results* r, uint64_t row, const char** seqid) { if (row >= r->rows.size() || !r->seq_ids) return -1; *seqid = (*r->seq_ids)[r->row_genome[row]][r->rows[row].seq_idx].c_str(); return 0; }
void lmg_results_free(lmg_results* r) { delete r; }

int lmg_wfa_batch(int device, const uint8_t* seqs, const uint64_t* off, int32_t n, int32_t adaptive, char** cigars, uint64_t* cigars_len) {
  try { CUDA_CHECK(cudaSetDevice(device)); cudaStream_t st = 0; std::vector<u8> qp, tp, qmk; std::vector<u64> qo(n + 1), to(n + 1); std::vector<HspJob> jobs(n); std::vector<ExtOut> ex(n);
    auto pack = [](const u8* s, u64 len, std::vector<u8>& out) { while (out.size() & 15) out.push_back(0); u64 o = out.size(); out.resize(o + (len + 3) / 4 + 16, 0); for (u64 i = 0; i < len; i++) out[o + (i >> 2)] |= (u8)(base2bit(s[i]) << (6 - 2 * (i & 3))); return o; };
    u64 opsCap = 0;
    for (int i = 0; i < n; i++) { u64 ql = off[2 * i + 1] - off[2 * i], tl = off[2 * i + 2] - off[2 * i + 1]; qo[i] = pack(seqs + off[2 * i], ql, qp); qmk.resize(qp.size(), 0); for (u64 x = 0; x < ql; x++) { u8 c = seqs[off[2 * i] + x] & 0xDF; if (!(c == 'A' || c == 'C' || c == 'G' || c == 'T')) qmk[qo[i] + (x >> 3)] |= (u8)(1u << (x & 7)); } to[i] = pack(seqs + off[2 * i + 1], tl, tp);
      HspJob J; memset(&J, 0, sizeof J); J.q = i; J.g = i; J.tBegin = 0; J.tEnd = (i32)tl - 1; J.rc = 0; J.qlen = (i32)ql; J.tlen = (i32)tl; jobs[i] = J; ExtOut e; memset(&e, 0, sizeof e); e.qs = 0; e.qe = (i32)ql; e.ts = 0; e.te = (i32)tl; ex[i] = e; opsCap += ql + tl + 4; }
    qp.resize(qp.size() + 64, 0); tp.resize(tp.size() + 64, 0); qmk.resize(qp.size(), 0);
    DBuf<u8> dq(qp.size(), st), dt(tp.size(), st), dqm(qmk.size(), st); dqm.from_host(qmk.data(), qmk.size()); dq.from_host(qp.data(), qp.size()); dt.from_host(tp.data(), tp.size()); DBuf<u64> dqo(n + 1, st), dto(n + 1, st); dqo.from_host(qo.data(), n + 1); dto.from_host(to.data(), n + 1);
    DBuf<HspJob> dj(n, st); dj.from_host(jobs.data(), n); DBuf<ExtOut> de(n, st); de.from_host(ex.data(), n); (void)opsCap;
    cudaDeviceProp pr; CUDA_CHECK(cudaGetDeviceProperties(&pr, device)); std::vector<WfaOut> hw; std::vector<u64> ops; u64 counters[16] = {0}; double dms[16] = {0};
    wfa_run_all(st, pr.multiProcessorCount, dj, de, ex, (u32)n, dq.p, dqm.p, dqo.p, dt.p, dto.p, 1, adaptive, hw, ops, counters, dms, pr.totalGlobalMem); std::string out;
    for (int i = 0; i < n; i++) { for (i64 x = (i64)hw[i].ops_n - 1; x >= 0; x--) { u64 op = ops[hw[i].ops_off + x]; out += std::to_string((u32)(op & 0xffffffffu)); out.push_back((char)(op >> 32)); } out.push_back('\n'); }
    char* c = (char*)malloc(out.size() + 1); memcpy(c, out.data(), out.size() + 1); *cigars = c; *cigars_len = out.size(); return 0; } catch (std::exception& e) { g_err = e.what(); cudaGetLastError(); return -1; }
}

int lmg_index_set_total_bases(lmg_index* ix, int64_t tb) { if (!ix || tb <= 0) { g_err = "lmg_index_set_total_bases: total_bases must be positive"; return -1; } std::lock_guard<std::mutex> lk(ix->mu); ix->imgp->total_bases = tb; return 0; }
int lmg_index_load_times(const lmg_index* ix, double* ms4) { for (int i = 0; i < 4; i++) ms4[i] = ix->img.load_ms[i]; return 0; }
int lmg_probe_model(const lmg_index* ix, uint64_t* s4) { for (int i = 0; i < 4; i++) s4[i] = ix->pstat[i]; return 0; }
// Random 32-byte-sector read rate of the device: every thread reads `per_thread` independent pseudo-random 8-byte words of a buffer far larger than L2.
// This is the physical ceiling of a lookup kernel whose accesses are dependent random sectors (DRAM row activations, not streaming bandwidth, bound it).
__global__ void __launch_bounds__(256) k_gather_bench(const u64* __restrict__ buf, u64 nwords, u64 n, int per_thread, u64 seed, u64* __restrict__ sink) {
  const u64 t = blockIdx.x * (u64)blockDim.x + threadIdx.x; if (t >= n) return; u64 acc = 0, r = mix64(seed ^ t);
  for (int i = 0; i < per_thread; i++) { r = mix64(r); acc += buf[(r % nwords) & ~3ull]; }   // sector-aligned, independent addresses (no pointer chasing: the rate, not the latency)
  if (acc == 0x1234567ull) sink[0] = acc;
}
int lmg_gather_bench(int device, uint64_t bytes, uint64_t n_threads, int32_t per_thread, int32_t iters, double* out4) {
  try { CUDA_CHECK(cudaSetDevice(device)); u64* buf = nullptr; u64* sink = nullptr; const u64 nwords = bytes / 8; CUDA_CHECK(cudaMalloc((void**)&buf, nwords * 8)); CUDA_CHECK(cudaMalloc((void**)&sink, 64)); CUDA_CHECK(cudaMemset(buf, 1, nwords * 8));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); double best = 1e30, sum = 0;
    for (int it = -1; it < iters; it++) { cudaEventRecord(e0); k_gather_bench<<<(unsigned)((n_threads + 255) / 256), 256>>>(buf, nwords, n_threads, per_thread, 77 + it, sink); KERNEL_CHECK(); cudaEventRecord(e1); CUDA_CHECK(cudaEventSynchronize(e1)); float f = 0; cudaEventElapsedTime(&f, e0, e1); if (it >= 0) { best = std::min(best, (double)f); sum += f; } }
    const double acc = (double)n_threads * per_thread; out4[0] = acc / (best * 1e-3); out4[1] = out4[0] * 32 / 1e9; out4[2] = best; out4[3] = iters > 0 ? sum / iters : 0; cudaFree(buf); cudaFree(sink); cudaEventDestroy(e0); cudaEventDestroy(e1); return 0; }
  catch (std::exception& e) { g_err = e.what(); cudaGetLastError(); return -1; }
}
int lmg_index_synth(int device, int32_t masks, uint64_t per_mask, uint64_t seed, int32_t mask_lo, int32_t mask_hi, int32_t with_values, lmg_index** out) {
  try { int ndev = 0; CUDA_CHECK(cudaGetDeviceCount(&ndev)); if (ndev == 0) throw std::runtime_error("no CUDA device: the LexicMap GPU path has no CPU fallback"); if (masks < 64 || per_mask == 0 || per_mask >= (1ull << 31) || mask_lo < 0 || mask_hi > masks || mask_lo >= mask_hi) throw std::runtime_error("lmg_index_synth: bad arguments");
    Image* im = new Image; lmg_index* ix = nullptr; try { im->synth(device, masks, per_mask, seed, mask_lo, mask_hi, with_values != 0); im->info.chunks = 0; im->info.partitions = im->NA; im->info.genome_batches = 0; im->synth_per = per_mask; im->synth_seed = seed; ix = make_ctx(im, true, device); } catch (...) { if (!ix) { im->release(); delete im; } throw; }
    *out = ix; return 0; } catch (std::exception& e) { g_err = e.what(); cudaGetLastError(); return -1; }
}
int lmg_probe_bench(lmg_index* ix, uint64_t n_queries, uint64_t seed, int32_t min_prefix, int32_t iters, double* out16) {
  try { std::lock_guard<std::mutex> lk(ix->mu); const Image& I = ix->img; if (!I.synth_per) throw std::runtime_error("lmg_probe_bench needs an index made by lmg_index_synth"); CUDA_CHECK(cudaSetDevice(I.device)); cudaStream_t st = ix->st;
    if (n_queries == 0 || n_queries >= (1ull << 31)) throw std::runtime_error("lmg_probe_bench: 1 <= n_queries < 2^31"); ProbeParams P = probe_params(I, min_prefix); for (int i = 0; i < 16; i++) out16[i] = 0;
    DBuf<Surv> surv(2 * n_queries + 64, st); DBuf<u32> nsv(1, st), nh(1, st), bcnt0((u64)I.m + 1, st); DBuf<u64> dstats(8, st); nsv.zero(); dstats.zero(); bcnt0.zero(); P.bcnt = bcnt0.p; cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, st); k_c5_gen<<<cdiv((i64)n_queries, 256), 256, 0, st>>>(P, I.d_masks, I.d_mask_pstart, I.mask_pbits, I.synth_per, I.synth_seed, seed, n_queries, I.mask_lo, I.mask_hi, surv.p, nsv.p, (u32)(2 * n_queries), dstats.p); KERNEL_CHECK(); cudaEventRecord(e1, st);
    const u32 ns = nsv.to_host()[0]; float fg = 0; cudaEventElapsedTime(&fg, e0, e1); out16[0] = (double)dstats.to_host()[0]; out16[1] = (double)ns; out16[9] = fg; DBuf<ProbeHit> hits((u64)ns + 64, st); double tsum = 0, tmin = 1e30, rsum = 0; DBuf<Surv> grouped((u64)ns + 64, st); DBuf<u32> bcnt((u64)I.m + 1, st); const Surv* sp = surv.p; cudaEvent_t er; cudaEventCreate(&er);
    for (int it = -1; it < iters && ns; it++) { nh.zero(); cudaEventRecord(er, st); CUDA_CHECK(cudaMemcpyAsync(bcnt.p, bcnt0.p, ((size_t)I.m + 1) * 4, cudaMemcpyDeviceToDevice, st)); if (regroup_survivors(st, surv.p, ns, I.m, grouped.p, bcnt.p, true)) sp = grouped.p;   /* the regrouping pass is part of every lookup: timed next to the kernel ([11]) */
      cudaEventRecord(e0, st); k_probe_find2<false><<<cdiv(ns, 256), 256, 0, st>>>(P, sp, ns, hits.p, nh.p, ns, nullptr); KERNEL_CHECK(); cudaEventRecord(e1, st); CUDA_CHECK(cudaEventSynchronize(e1)); float f = 0, fr = 0; cudaEventElapsedTime(&f, e0, e1); cudaEventElapsedTime(&fr, er, e0); if (it >= 0) { tsum += f; rsum += fr; tmin = std::min(tmin, (double)f); } }
    out16[2] = iters > 0 ? tsum / iters : 0; out16[11] = iters > 0 ? rsum / iters : 0; out16[10] = tmin < 1e29 ? tmin : 0; out16[3] = (double)nh.to_host()[0]; cudaEventDestroy(er);
    if (ns) { nh.zero(); dstats.zero(); k_probe_find2<true><<<cdiv(ns, 256), 256, 0, st>>>(P, sp, ns, hits.p, nh.p, ns, dstats.p); KERNEL_CHECK(); auto sd = dstats.to_host(); out16[4] = (double)sd[4]; out16[5] = (double)sd[5]; out16[6] = (double)sd[6]; out16[7] = (double)sd[2]; out16[8] = (double)sd[3]; }
    cudaEventDestroy(e0); cudaEventDestroy(e1); CUDA_CHECK(cudaStreamSynchronize(st)); return 0; } catch (std::exception& e) { g_err = e.what(); cudaGetLastError(); return -1; }
}

int lmg_queries_upload(lmg_index* ix, const uint8_t* seqs, const uint64_t* off, int32_t n, lmg_queries** out) {
  try { std::lock_guard<std::mutex> lk(ix->mu); CUDA_CHECK(cudaSetDevice(ix->img.device)); lmg_queries* Q = new lmg_queries; Q->cut = lane_cuts(off, n, pick_lanes(nullptr, n, off[n] - off[0])); const int L = (int)Q->cut.size() - 1; Q->parts.resize(L);
    for (int l = 0; l < L && n > 0; l++) { lmg_index* lx = lane_ctx(ix, l); upload_queries(lx, seqs, off + Q->cut[l], Q->cut[l + 1] - Q->cut[l], Q->parts[l]); CUDA_CHECK(cudaStreamSynchronize(lx->st)); } *out = Q; return 0; } catch (std::exception& e) { g_err = e.what(); return -1; }
}
int lmg_search_staged(lmg_index* ix, const lmg_params* p, lmg_queries* q, lmg_results** out) {
  try { std::lock_guard<std::mutex> lk(ix->mu); lmg_results* R = new lmg_results; if (q->cut.back() <= 0) { R->seq_ids = ix->img.seq_ids_p; *out = R; return 0; }   /* empty batch: no rows, nothing launched */ try { search_lanes(ix, p, nullptr, nullptr, q->cut.back(), *R, q); } catch (...) { delete R; throw; } *out = R; return 0; }
  catch (std::exception& e) { g_err = e.what(); cudaGetLastError(); return -1; }
}
void lmg_queries_free(lmg_index* ix, lmg_queries* q) { if (!q) return; cudaSetDevice(ix->img.device); delete q; }
}  // extern "C"
