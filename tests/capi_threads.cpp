// capi_threads.cpp — C-ABI client used by tests/test_gpu_boundary.py: calls lmg_search_batch on ONE index handle from two host threads at the
// same time (as goroutines locked to OS threads would, search.go:589-602) and checks that both get exactly the rows a serial call returns.
// usage: capi_threads <index.lmi> <queries.fasta>     exit 0 = identical
#include "lexicmap_gpu.h"
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

struct Batch { std::string seqs; std::vector<uint64_t> off{0}; int n = 0; };
struct Out { std::vector<lmg_hsp> rows; std::string pool; int rc = 0; std::string err; };

static void run(lmg_index* idx, const Batch& b, Out& o) {
  lmg_params p; lmg_default_params(&p); p.output_seq = 1; lmg_results* r = nullptr;
  o.rc = lmg_search_batch(idx, &p, (const uint8_t*)b.seqs.data(), b.off.data(), b.n, &r); if (o.rc) { o.err = lmg_last_error(); return; }
  const lmg_hsp* rows; uint64_t n, pl; const char* pool; lmg_results_rows(r, &rows, &n, &pool, &pl); o.rows.assign(rows, rows + n); o.pool.assign(pool ? pool : "", pl); lmg_results_free(r);
}
static bool same(const Out& a, const Out& b) { return a.rc == 0 && b.rc == 0 && a.rows.size() == b.rows.size() && a.pool == b.pool && (a.rows.empty() || !memcmp(a.rows.data(), b.rows.data(), a.rows.size() * sizeof(lmg_hsp))); }

int main(int argc, char** argv) {
  if (argc < 3) return 2; lmg_index* idx = nullptr; if (lmg_index_open(argv[1], 0, 0, 1, &idx)) { fprintf(stderr, "open: %s\n", lmg_last_error()); return 3; }
  Batch b[2]; { std::ifstream f(argv[2]); std::string line, seq; int q = 0; auto push = [&]() { if (seq.empty()) return; Batch& t = b[q++ & 1]; t.seqs += seq; t.off.push_back(t.seqs.size()); t.n++; seq.clear(); };
    while (std::getline(f, line)) { if (!line.empty() && line[0] == '>') push(); else seq += line; } push(); }
  for (auto& t : b) t.seqs.append(16, '\0');
  Out serial[2], par[2]; run(idx, b[0], serial[0]); run(idx, b[1], serial[1]);
  for (int round = 0; round < 3; round++) { std::thread t0(run, idx, std::cref(b[0]), std::ref(par[0])), t1(run, idx, std::cref(b[1]), std::ref(par[1])); t0.join(); t1.join();
    for (int i = 0; i < 2; i++) if (!same(serial[i], par[i])) { fprintf(stderr, "round %d batch %d differs (rc %d %s)\n", round, i, par[i].rc, par[i].err.c_str()); return 1; } }
  // results outlive the index (the id table is shared): close first, then read a sequence id
  lmg_params p; lmg_default_params(&p); lmg_results* r = nullptr; if (lmg_search_batch(idx, &p, (const uint8_t*)b[0].seqs.data(), b[0].off.data(), b[0].n, &r)) return 4;
  lmg_index_close(idx); const char* sid = nullptr; const lmg_hsp* rows; uint64_t n; lmg_results_rows(r, &rows, &n, nullptr, nullptr); if (n && (lmg_results_seq_id(r, 0, &sid) || !sid || !*sid)) return 5; lmg_results_free(r);
  printf("ok %zu + %zu rows\n", serial[0].rows.size(), serial[1].rows.size()); return 0;
}
