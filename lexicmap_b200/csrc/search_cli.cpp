// search_cli.cpp — `lexicmap-gpu search`: host driver with the reference's `lexicmap search` flag surface and TSV output
// (reference: lexicmap/cmd/search.go — flags :631-731, option checks :163-230, reader loop :554-605, header :426-430,
// printResult :437-533). The reference host is Go (no toolchain in this image), so this driver is C++ over the same C ABI a
// cgo binding would use. All search work happens in liblexicmap_gpu.so; there is no CPU search path here.
#include "lexicmap_gpu.h"
#include <zlib.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

static void die(const std::string& m) { fprintf(stderr, "[ERRO] %s\n", m.c_str()); exit(255); }  // checkError -> os.Exit(-1), util-cli.go:35-40

struct Opts { std::string index, out = "-"; std::vector<std::string> files; lmg_params p; bool all = false, show_idx = false, quiet = false; int device = 0; long batch_bases = 32 << 20; int batch_queries = 20000; };

static void usage() {
  fprintf(stderr,
          "lexicmap-gpu search: B200 implementation of `lexicmap search` (same flags, same TSV)\n\n"
          "Usage:\n  lexicmap-gpu search -d <index.lmi> [flags] <query.fasta[.gz]> ...\n\nFlags (defaults as the reference):\n"
          "  -d, --index string                 index directory created by `lexicmap index`\n  -o, --out-file string              out file (default \"-\")\n"
          "  -a, --all                          output more columns: cigar, qseq, sseq, align\n"
          "      --show-sseq-idx                prefix sseqid with c<chunk>/<chunks>:s<seq>/<seqs>:\n"
          "  -p, --seed-min-prefix int          (default 15)\n  -P, --seed-min-single-prefix int   (default 17)\n      --seed-max-gap int             (default 50)\n      --seed-max-dist int            (default 1000)\n"
          "  -n, --top-n-genomes int            (default 0)\n  -N, --top-n-chains int             (default 0)\n      --align-ext-len int            (default 1000)\n      --align-max-gap int            (default 20)\n"
          "      --align-band int               (default 100)\n  -l, --align-min-match-len int      (default 50)\n  -i, --align-min-match-pident float (default 70)\n  -q, --min-qcov-per-hsp float       (default 0)\n"
          "  -Q, --min-qcov-per-genome float    (default 0)\n  -e, --max-evalue float             (default 10)\n      --gpu int                      CUDA device (default 0)\n      --quiet\n");
}

int main(int argc, char** argv) {
  if (argc < 2 || strcmp(argv[1], "search")) { usage(); return 2; }
  Opts o; lmg_default_params(&o.p);
  auto need = [&](int& i) -> const char* { if (i + 1 >= argc) die(std::string("flag needs an argument: ") + argv[i]); return argv[++i]; };
  for (int i = 2; i < argc; i++) { std::string a = argv[i];
    if (a == "-d" || a == "--index") o.index = need(i); else if (a == "-o" || a == "--out-file") o.out = need(i); else if (a == "-a" || a == "--all") o.all = true; else if (a == "--show-sseq-idx") o.show_idx = true;
    else if (a == "-p" || a == "--seed-min-prefix") o.p.min_prefix = atoi(need(i)); else if (a == "-P" || a == "--seed-min-single-prefix") o.p.min_single_prefix = atoi(need(i));
    else if (a == "--seed-max-gap") o.p.max_gap = (float)atoi(need(i)); else if (a == "--seed-max-dist") o.p.max_distance = (float)atoi(need(i));
    else if (a == "-n" || a == "--top-n-genomes") o.p.top_n_genomes = atoi(need(i)); else if (a == "-N" || a == "--top-n-chains") o.p.top_n_chains = atoi(need(i));
    else if (a == "--align-ext-len") o.p.ext_len = atoi(need(i)); else if (a == "--align-max-gap") o.p.align_max_gap = atoi(need(i)); else if (a == "--align-band") o.p.align_band = atoi(need(i));
    else if (a == "-l" || a == "--align-min-match-len") o.p.align_min_len = atoi(need(i)); else if (a == "-i" || a == "--align-min-match-pident") o.p.min_pident = atof(need(i));
    else if (a == "-q" || a == "--min-qcov-per-hsp") o.p.min_qcov_hsp = atof(need(i)); else if (a == "-Q" || a == "--min-qcov-per-genome") o.p.min_qcov_genome = atof(need(i));
    else if (a == "-e" || a == "--max-evalue") o.p.max_evalue = atof(need(i)); else if (a == "--gpu") o.device = atoi(need(i)); else if (a == "--quiet") o.quiet = true;
    else if (a == "-j" || a == "--threads" || a == "-J" || a == "--max-query-conc" || a == "--max-open-files" || a == "--gc-interval" || a == "-S" || a == "--max-seed-matching-conc") need(i);   // accepted, meaningless on the GPU path
    else if (a == "-T" || a == "--taxdump" || a == "-G" || a == "--genome2taxid" || a == "-t" || a == "--taxids" || a == "--taxid-file" || a == "-k" || a == "--keep-genomes-without-taxid")
      die("taxonomy filtering (" + a + ") is not part of the GPU search path; filter the TSV by sgenome afterwards");   // search.go:236-330, out of scope
    else if (a == "-w" || a == "--load-whole-seeds") die("-w/--load-whole-seeds selects the reference's in-memory searcher, which tests the reversed flag of every seed value (kv-searcher2.go:302) where the default on-disk searcher tests the first value of a k-mer (kv-searcher.go:466-488); the GPU image always holds all seeds in memory and implements the default semantics only: run without -w");
    else if (a == "--debug") { if (!o.quiet) fprintf(stderr, "[INFO] --debug has no effect on the GPU path (LMG_DEBUG_TIMING=1 prints the host-side stage times)\n"); } else if (a == "-h" || a == "--help") { usage(); return 0; }
    else if (a[0] == '-' && a.size() > 1) die("unknown flag: " + a); else o.files.push_back(a); }
  // option checks, search.go:163-230
  if (o.index.empty()) die("flag -d/--index needed");
  if (o.p.min_prefix > 32 || o.p.min_prefix < 5) die("the value of flag -p/--seed-min-prefix should be in the range of [5, 32]");
  if (o.p.min_single_prefix > 32) die("the value of flag -P/--seed-min-single-prefix should be <= 32");
  if (o.p.min_single_prefix < o.p.min_prefix) die("the value of flag -P/--seed-min-single-prefix should be >= that of -p/--seed-min-prefix");
  if (o.p.align_min_len < o.p.min_single_prefix) die("the value of flag -l/--align-min-match-len should be >= that of -M/--seed-min-single-prefix");
  if (o.p.align_band < o.p.align_max_gap) die("the value of flag --align-band should not be smaller thant the value of --align-max-gap");
  if (o.p.min_qcov_genome > 100 || o.p.min_qcov_hsp > 100) die("query coverage should be in range of [0, 100]");
  if (o.p.min_pident < 60 || o.p.min_pident > 100) die("the value of flag -i/--align-min-match-pident should be in range of [60, 100]");
  if (o.files.empty()) o.files.push_back("-");
  o.p.output_seq = o.all;
  auto t0 = std::chrono::steady_clock::now();
  lmg_index* idx = nullptr; if (lmg_index_open(o.index.c_str(), o.device, 0, 1, &idx)) die(std::string("failed to load index: ") + lmg_last_error());
  if (!o.quiet) fprintf(stderr, "[INFO] index loaded to GPU %d in %.3fs\n", o.device, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
  FILE* out = o.out == "-" ? stdout : fopen(o.out.c_str(), "w"); if (!out) die("cannot write " + o.out);
  fprintf(out, "query\tqlen\thits\tsgenome\tsseqid\tqcovGnm\tcls\thsp\tqcovHSP\talenHSP\tpident\tgaps\tqstart\tqend\tsstart\tsend\tsstr\tslen\tevalue\tbitscore"); if (o.all) fprintf(out, "\tcigar\tqseq\tsseq\talign"); fprintf(out, "\n");
  std::vector<std::string> ids; std::string seqs; std::vector<uint64_t> off(1, 0); unsigned long long total = 0, matched = 0; auto t1 = std::chrono::steady_clock::now();
  auto flush = [&]() {
    if (ids.empty()) return; lmg_results* r = nullptr; seqs.append(16, '\0');
    if (lmg_search_batch(idx, &o.p, (const uint8_t*)seqs.data(), off.data(), (int32_t)ids.size(), &r)) die(std::string("search failed: ") + lmg_last_error());
    const lmg_hsp* rows; uint64_t n, pl; const char* pool; lmg_results_rows(r, &rows, &n, &pool, &pl); uint32_t lastq = 0xffffffffu;
    for (uint64_t i = 0; i < n; i++) { const lmg_hsp& h = rows[i]; if (h.query != lastq) { matched++; lastq = h.query; } const char *gname, *sid; lmg_genome_name(idx, h.genome, &gname); lmg_results_seq_id(r, i, &sid);
      fprintf(out, "%s\t%llu\t%u\t%s\t", ids[h.query].c_str(), (unsigned long long)(off[h.query + 1] - off[h.query]), h.hits, gname);
      if (o.show_idx) fprintf(out, "c%u/%u:s%u/%u:", h.chunk_idx + 1, h.n_chunks, h.seq_idx + 1, h.n_seqs);
      fprintf(out, "%s\t%.3f\t%d\t%d\t%.3f\t%d\t%.3f\t%d\t%d\t%d\t%d\t%d\t%c\t%d\t%.2e\t%d", sid, h.qcov_gnm, h.cls, h.hsp, h.qcov_hsp, h.alen, h.pident, h.gaps, h.qb + 1, h.qe + 1, h.tb + 1, h.te + 1, h.rc ? '-' : '+', h.seq_len, h.evalue, h.bitscore);
      if (o.all) { const char* t = pool + h.cigar_off + h.cigar_len; fprintf(out, "\t"); fwrite(pool + h.cigar_off, 1, h.cigar_len, out); for (int x = 0; x < 3; x++) { fprintf(out, "\t"); fwrite(t + (size_t)x * h.alen, 1, h.alen, out); } }   // pool entry: cigar | qseq | sseq | align
      fprintf(out, "\n"); }
    lmg_results_free(r); total += ids.size(); ids.clear(); seqs.clear(); off.assign(1, 0); fflush(out);
  };
  // whole lines of any length (a FASTQ record of an ONT read is one line of > 64 kb): gzgets chunks are joined until the newline
  auto read_line = [](gzFile g, std::string& line) -> bool { static char chunk[1 << 16]; line.clear(); bool any = false; while (gzgets(g, chunk, sizeof chunk)) { any = true; size_t l = strlen(chunk); line.append(chunk, l); if (l && chunk[l - 1] == '\n') break; }
    while (!line.empty() && (line.back() == '\n' || line.back() == '\r')) line.pop_back(); return any; };
  for (const std::string& f : o.files) { gzFile g = (f == "-") ? gzdopen(0, "rb") : gzopen(f.c_str(), "rb"); if (!g) die("cannot open " + f); gzbuffer(g, 1 << 20); std::string id, seq, line; bool have = false, fastq = false; int fqline = 0;
    auto push = [&]() { if (!have) return; if (seq.size() >= 31) { ids.push_back(id); seqs += seq; off.push_back(seqs.size()); } else total++;   // queries shorter than k are skipped, search.go:571-575
      have = false; seq.clear(); if ((long)seqs.size() >= o.batch_bases || (int)ids.size() >= o.batch_queries) flush(); };
    while (read_line(g, line)) { const char* buf = line.c_str();
      if (fastq) { fqline++; if (fqline == 1) seq += line; else if (fqline == 3) { fastq = false; push(); } continue; }
      if (buf[0] == '>' || buf[0] == '@') { push(); have = true; const char* e = buf + 1; while (*e && *e != ' ' && *e != '\t') e++; id.assign(buf + 1, e); if (buf[0] == '@') { fastq = true; fqline = 0; } }
      else if (have) seq += line; }
    push(); gzclose(g); }
  flush();
  if (!o.quiet) { double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count(); fprintf(stderr, "[INFO] processed queries: %llu, speed: %.3f queries per minute\n[INFO] %.4f%% (%llu/%llu) queries matched\n", total, total / s * 60, total ? 100.0 * matched / total : 0.0, matched, total); }
  if (out != stdout) fclose(out); lmg_index_close(idx); return 0;
}
