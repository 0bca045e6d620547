// common.cuh — shared host/device helpers for the B200 engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <stdexcept>
#include <vector>
#include <algorithm>

typedef uint8_t u8; typedef uint16_t u16; typedef uint32_t u32; typedef uint64_t u64; typedef int16_t i16; typedef int32_t i32; typedef int64_t i64;

#define CUDA_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) throw std::runtime_error(std::string("CUDA error ") + cudaGetErrorString(e_) + " at " __FILE__ ":" + std::to_string(__LINE__)); } while (0)
#define KERNEL_CHECK() CUDA_CHECK(cudaGetLastError())

// Per-index bump arena for the per-batch device buffers: cudaMallocAsync pools showed sporadic 200-300 ms stalls when multi-GB
// blocks had to be re-mapped between calls; a grow-only arena that is reset after every batch is deterministic.
struct ArenaExhausted : std::runtime_error { using std::runtime_error::runtime_error; };   // the caller may retry with a smaller batch
struct Arena {
  struct Chunk { char* base; size_t size; };
  std::vector<Chunk> chunks; size_t cur = 0, off = 0; u64 epoch = 1;   // epoch: bumped by reset(); a buffer from an earlier epoch is already gone and must not rewind the cursor
  static size_t al(size_t b) { return (b + 511) & ~(size_t)511; }
  void* alloc(size_t bytes) {
    bytes = al(bytes + 64);
    while (cur < chunks.size() && off + bytes > chunks[cur].size) { cur++; off = 0; }
    if (cur >= chunks.size()) { size_t tot = 0; for (auto& c : chunks) tot += c.size; size_t sz = std::max(bytes, std::max<size_t>(tot / 4, (size_t)1 << 30)); Chunk c; c.size = sz;
      cudaError_t e = cudaMalloc((void**)&c.base, sz); if (e != cudaSuccess) { cudaGetLastError(); c.size = sz = bytes; e = cudaMalloc((void**)&c.base, sz); }
      if (e != cudaSuccess) { cudaGetLastError(); throw ArenaExhausted(std::string("device arena: cudaMalloc failed: ") + cudaGetErrorString(e)); } chunks.push_back(c); cur = chunks.size() - 1; off = 0; }
    void* p = chunks[cur].base + off; off += bytes; return p;
  }
  void free(void* p, size_t bytes) { bytes = al(bytes + 64); if (cur < chunks.size() && (char*)p + bytes == chunks[cur].base + off) off -= bytes; }   // LIFO frees are recycled
  void reset() { cur = 0; off = 0; epoch++; if (chunks.size() > 1) { size_t tot = 0; for (auto& c : chunks) { tot += c.size; cudaFree(c.base); } chunks.clear(); Chunk c; c.size = tot; if (cudaMalloc((void**)&c.base, tot) == cudaSuccess) chunks.push_back(c); else cudaGetLastError(); } }
  void release() { for (auto& c : chunks) cudaFree(c.base); chunks.clear(); cur = off = 0; }
};
static thread_local Arena* g_arena = nullptr;
struct ArenaScope { Arena* prev; ArenaScope(Arena* a) : prev(g_arena) { g_arena = a; } ~ArenaScope() { g_arena = prev; } };

// device buffer: from the current arena when one is active, otherwise stream-ordered cudaMallocAsync
template <class T> struct DBuf {
  T* p = nullptr; size_t n = 0; cudaStream_t st = 0; Arena* ar = nullptr; u64 ep = 0;
  DBuf() {}
  DBuf(size_t n_, cudaStream_t s) { alloc(n_, s); }
  void alloc(size_t n_, cudaStream_t s) { free(); n = n_; st = s; if (!n) return; if (g_arena) { ar = g_arena; ep = ar->epoch; p = (T*)ar->alloc(n * sizeof(T)); } else { ar = nullptr; CUDA_CHECK(cudaMallocAsync((void**)&p, n * sizeof(T) + 64, s)); } }
  void free() { if (p) { if (ar) { if (ar->epoch == ep) ar->free(p, n * sizeof(T)); } else cudaFreeAsync(p, st); p = nullptr; n = 0; ar = nullptr; } }
  void zero() { if (p) CUDA_CHECK(cudaMemsetAsync(p, 0, n * sizeof(T), st)); }
  void fill_ff() { if (p) CUDA_CHECK(cudaMemsetAsync(p, 0xff, n * sizeof(T), st)); }
  ~DBuf() { free(); }
  DBuf(const DBuf&) = delete; DBuf& operator=(const DBuf&) = delete;
  DBuf(DBuf&& o) noexcept { p = o.p; n = o.n; st = o.st; ar = o.ar; ep = o.ep; o.p = nullptr; o.n = 0; }
  DBuf& operator=(DBuf&& o) noexcept { if (this != &o) { free(); p = o.p; n = o.n; st = o.st; ar = o.ar; ep = o.ep; o.p = nullptr; o.n = 0; } return *this; }
  std::vector<T> to_host(size_t cnt = (size_t)-1) const { if (cnt == (size_t)-1) cnt = n; std::vector<T> h(cnt); if (cnt) { CUDA_CHECK(cudaMemcpyAsync(h.data(), p, cnt * sizeof(T), cudaMemcpyDeviceToHost, st)); CUDA_CHECK(cudaStreamSynchronize(st)); } return h; }
  void from_host(const T* h, size_t cnt) { if (cnt) CUDA_CHECK(cudaMemcpyAsync(p, h, cnt * sizeof(T), cudaMemcpyHostToDevice, st)); }
};

// ---------------------------------------------------------------- k-mer helpers (device)
// 2-bit codes A0 C1 G2 T3; degenerate bases as genome.base2bit (reference genome/genome.go:1427-1444)
__host__ __device__ inline u32 base2bit(u8 c) {
  switch (c) { case 'C': case 'c': case 'B': case 'b': case 'S': case 's': case 'Y': case 'y': return 1;
    case 'G': case 'g': case 'K': case 'k': return 2; case 'T': case 't': case 'U': case 'u': return 3; default: return 0; }
}
// base i of a 2-bit packed sequence (first base in bits 7-6 of byte 0)
__device__ __forceinline__ u32 get_base(const u8* __restrict__ p, u64 i) { return (p[i >> 2] >> (6 - 2 * (i & 3))) & 3; }

__host__ __device__ inline u64 kmer_reverse62(u64 c, int k) {  // kmers.MustReverse: reverse base order
  // swap 2-bit groups: full 64-bit group reversal then shift
  c = ((c >> 2) & 0x3333333333333333ull) | ((c & 0x3333333333333333ull) << 2);
  c = ((c >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((c & 0x0F0F0F0F0F0F0F0Full) << 4);
  c = ((c >> 8) & 0x00FF00FF00FF00FFull) | ((c & 0x00FF00FF00FF00FFull) << 8);
  c = ((c >> 16) & 0x0000FFFF0000FFFFull) | ((c & 0x0000FFFF0000FFFFull) << 16);
  c = (c >> 32) | (c << 32);
  return c >> (64 - 2 * k);
}
__host__ __device__ inline u64 kmer_ns(u64 b, int k) { u64 c = 0; for (int i = 0; i < k; i++) c = (c << 2) | b; return c; }

// util.IsLowComplexityDust (reference util/kmers.go:162-328): 3-mer windows i = 0..k-2 of (code >> 2i) & 63, score = sum_c c(c-1)/2 > 50.
// sum_c c(c-1)/2 == number of window pairs (i<j) with equal content; counted with bit tricks, no table.
__host__ __device__ inline bool dust_low_complexity(u64 code, int k) {
  const int W = k - 1;  // number of windows
  int score = 0;
  for (int d = 1; d < W; d++) {
    u64 x = code ^ (code >> (2 * d));
    u64 t = (x | (x >> 1)) & 0x5555555555555555ull;         // per-base "differs" flag at even bits
    u64 u = t | (t >> 2) | (t >> 4);                        // window i differs iff bit 2i set
    int nwin = W - d;                                       // windows i = 0..W-d-1 pair with i+d
    u64 valid = (nwin >= 32) ? 0x5555555555555555ull : (0x5555555555555555ull & ((1ull << (2 * nwin)) - 1));
#ifdef __CUDA_ARCH__
    score += __popcll(~u & valid);
#else
    score += __builtin_popcountll(~u & valid);
#endif
  }
  return score > 50;
}
__host__ __device__ inline bool kmer_low_complexity(u64 kmer, int k) {  // reference lib-index-search.go:1222-1238
  u64 ttt = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
  return kmer == (ttt / 3) || kmer == (ttt / 3 * 2) || kmer == ttt || dust_low_complexity(kmer, k);
}

// argmin over sorted a[lo,hi) of (a[i] XOR x): returns the equal-range [lo,hi) of the winning key
template <class Ptr>
__device__ __forceinline__ void xor_argmin_range(Ptr a, u32& lo, u32& hi, u64 x) {
  for (;;) {
    u64 al = a[lo], ah = a[hi - 1];
    if (al == ah) return;
    u64 b = 1ull << (63 - __clzll(al ^ ah));
    u32 l = lo, h = hi;  // first index with bit b set (keys share all higher bits, so sorted by this bit)
    while (l < h) { u32 m = (l + h) >> 1; if (a[m] & b) h = m; else l = m + 1; }
    if (x & b) lo = l; else hi = l;
  }
}
