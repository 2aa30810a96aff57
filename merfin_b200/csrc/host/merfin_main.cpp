// merfin-compatible command line front-end over libmerfin_b200 (the C ABI in include/merfin_b200.h).
//
// Same flags, same output files and the same final stderr summary as the reference's main()
// (merfin.C:72-430); the per-base k-mer multiplicity evaluation runs on the GPU, everything
// written to files is produced from its integer results with the reference's double arithmetic.
// Extra flags: -device <n> (CUDA device, default 0), -nofuse (keep two per-database probes).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/merfin_b200.h"
#include "dump_driver.h"
#include "fasta.h"

enum { OP_NONE = 0, OP_HIST, OP_COMPL, OP_DUMP, OP_FILTER, OP_POLISH, OP_BETTER, OP_STRICT, OP_LOOSE };

struct Globals {
  const char *seqName = nullptr, *seqDBname = nullptr, *readDBname = nullptr, *prob = nullptr, *vcfName = nullptr, *outName = nullptr;
  double peak = 0;
  uint64_t minV = 0, maxV = ~0ull;
  uint32_t comb = 15;
  uint32_t threads = 0;          // host threads for the -dump formatter (0 = all cores)
  int report = OP_NONE, device = 0;
  uint32_t shard = 0, nshards = 1;   // -shard i/n: variant modes score cluster range i of n of every record
  std::vector<int> devices;          // -devices a,b,...: -hist / -completeness / variant modes shard their work over these GPUs
  bool nosplit = false, skipMissing = false, debug = false, nofuse = false;
  std::vector<uint32_t> probK;
  std::vector<double> probP;
};

// 'path' for /bin/sh: single quotes, with every embedded single quote written as '\''
static std::string shell_quote(const std::string &p) {
  std::string q = "'";
  for (char c : p) { if (c == '\'') q += "'\\''"; else q += c; }
  return q + "'";
}

// MFN_TIMELINE=1: seconds since the start of main at every phase, to stderr
static std::chrono::steady_clock::time_point g_t0;
static void mark(const char *what) {
  static const bool on = getenv("MFN_TIMELINE") != nullptr;
  if (on) fprintf(stderr, "[timeline] %8.3f s  %s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - g_t0).count(), what);
}

#define MFN_TRY_MAIN(call) do { int s_ = (call); if (s_ != MFN_OK) return s_; } while (0)

// Output files as the reference's compressedFileWriter opens them (files-compressed.C:95-141): a name ending in
// .gz / .bz2 / .xz is written through the matching compressor, anything else as it is.
struct OutFile {
  FILE *f = nullptr;
  bool piped = false;
  bool open(const std::string &path) {
    auto ends = [&](const char *e) { const size_t n = strlen(e); return path.size() > n && path.compare(path.size() - n, n, e) == 0; };
    const char *tool = ends(".gz") ? "gzip -1c" : ends(".bz2") ? "bzip2 -9c" : ends(".xz") ? "xz -1c" : nullptr;
    if (tool) { f = popen((std::string(tool) + " > " + shell_quote(path)).c_str(), "w"); piped = true; }
    else f = fopen(path.c_str(), "w");
    return f != nullptr;
  }
  void close() { if (f) { if (piped) pclose(f); else fclose(f); f = nullptr; } }
};

static void die(const char *what) {
  fprintf(stderr, "%s: %s\n", what, mfn_last_error());
  exit(1);
}

static void usage(const char *argv0, const std::vector<std::string> &err) {
  fprintf(stderr, "usage: %s <report-type>            \\\n", argv0);
  fprintf(stderr, "         -sequence <seq.fasta>     \\\n");
  fprintf(stderr, "         -readmers <read.meryl>    \\\n");
  fprintf(stderr, "         -peak     <haploid_peak>  \\\n");
  fprintf(stderr, "         -prob     <lookup_table>  \\\n");
  fprintf(stderr, "         -vcf      <input.vcf>     \\\n");
  fprintf(stderr, "         -output   <output>        \n\n");
  fprintf(stderr, "  Report types: -hist, -dump [-skipMissing], -completeness, -filter, -polish, -better, -strict, -loose\n");
  fprintf(stderr, "  Options:      -seqmers <seq.meryl>  -min m  -max m  -comb N  -nosplit  -debug\n");
  fprintf(stderr, "                -threads t  (host threads of the -dump formatter)   -memory m   (accepted for compatibility)\n");
  fprintf(stderr, "                -device d   CUDA device to use (default 0)\n");
  fprintf(stderr, "                -devices a,b,..  shard -hist / -completeness / the variant modes over several GPUs of this box (same output)\n");
  fprintf(stderr, "                -shard i/n  variant modes: score cluster range i of n of every record (records only in the output when i > 0)\n");
  fprintf(stderr, "  This is the B200 build of merfin's k-mer evaluation; flags and outputs follow merfin.\n\n");
  for (const std::string &e : err) fputs(e.c_str(), stderr);
  exit(1);
}

// -prob CSV: "readK,prob" per line, line i <-> multiplicity i (merfin-globals.C:21-62)
static void load_Kmetric(Globals &G) {
  if (!G.prob) return;
  fprintf(stderr, "-- Loading probability table '%s'.\n\n", G.prob);
  std::vector<char> data;
  std::string err;
  if (!slurp_file(G.prob, data, err)) {
    fprintf(stderr, "ERROR: Probability table (-prob) file '%s' doesn't exist!\n", G.prob);
    exit(1);
  }
  size_t p = 0;
  uint32_t lineNum = 0;
  while (p < data.size()) {
    size_t e = p;
    while (e < data.size() && data[e] != '\n') e++;
    std::string line(data.data() + p, e - p);
    p = e + 1;
    while (!line.empty() && line.back() == '\r') line.pop_back();
    std::vector<std::string> w;
    size_t i = 0;
    while (i < line.size()) {
      while (i < line.size() && line[i] == ',') i++;
      size_t j = i;
      while (j < line.size() && line[j] != ',') j++;
      if (j > i) w.emplace_back(line, i, j - i);
      i = j;
    }
    if (w.size() == 2) {
      uint32_t k = (uint32_t)strtoul(w[0].c_str(), nullptr, 10);
      double pr = strtod(w[1].c_str(), nullptr);
      G.probK.push_back(k); G.probP.push_back(pr);
      lineNum++;
      fprintf(stderr, "Copy-number: %u\t\tReadK: %u\tProbability: %f\n", lineNum, k, pr);
    } else {
      fprintf(stderr, "Copy-number: invalid line %u:  '%s'\n", lineNum, line.c_str());
    }
  }
}

static double histoQV(double kval, double ktot, uint32_t k) {          // merfin-histogram.C:22-31
  return -10.0 * log10(1.0 - pow(1.0 - kval / ktot, 1.0 / k));
}

// One GPU of a -devices run: its own ctx and its own replica of the tables.
struct Dev {
  int id = 0;
  mfn_ctx *ctx = nullptr;
  mfn_table *R = nullptr, *A = nullptr;
  mfn_accum *acc = nullptr;
  int status = MFN_OK;
  std::string error;
};

// run fn(d) for every device on its own host thread
template <class F> static void on_devices(std::vector<Dev> &devs, F fn) {
  std::vector<std::thread> th;
  for (size_t d = 1; d < devs.size(); d++) th.emplace_back([&, d]() { fn(devs[d], d); });
  fn(devs[0], 0);
  for (auto &t : th) t.join();
}

static bool devices_ok(std::vector<Dev> &devs, const char *what) {
  for (Dev &d : devs)
    if (d.status != MFN_OK) { fprintf(stderr, "%s (device %d): %s\n", what, d.id, d.error.c_str()); return false; }
  return true;
}

// Sum the accumulators of all devices into device 0's (device-to-device copies, NVLink between peers)
// and decode: per-record counters and the (v_read, v_asm) pairs, incl. the pairs outside the dense
// tier, which are gathered from every device's side table.
static int reduce_and_fetch(std::vector<Dev> &devs, uint32_t nseq, std::vector<mfn_seq_counts> *per, std::vector<mfn_pair_count> &pairs, uint64_t &np) {
  for (size_t d = 1; d < devs.size(); d++) {
    MFN_TRY_MAIN(mfn_ctx_sync(devs[d].ctx));
    MFN_TRY_MAIN(mfn_accum_add_from(devs[0].ctx, devs[0].acc, devs[d].acc));
  }
  uint64_t nwords = 0;
  uint32_t hw = 0, dr = 0, da = 0;
  MFN_TRY_MAIN(mfn_accum_layout(nseq, &hw, &dr, &da, &nwords));
  std::vector<uint64_t> buf(nwords);
  MFN_TRY_MAIN(mfn_accum_read(devs[0].ctx, devs[0].acc, buf.data()));
  std::vector<mfn_pair_count> sparse;
  if (buf[1]) {
    for (Dev &d : devs) {
      uint64_t n = 0;
      int st = mfn_accum_sparse_fetch(d.ctx, d.acc, nullptr, 0, &n);
      if (st != MFN_OK && st != MFN_ERR_CAPACITY) return st;
      if (!n) continue;
      const size_t at = sparse.size();
      sparse.resize(at + n);
      MFN_TRY_MAIN(mfn_accum_sparse_fetch(d.ctx, d.acc, sparse.data() + at, n, &n));
    }
  }
  if (per) per->assign(nseq + 1, mfn_seq_counts{0, 0});
  int st = mfn_accum_decode_gathered(buf.data(), nseq, sparse.data(), sparse.size(), per ? per->data() : nullptr, pairs.data(), pairs.size(), &np);
  if (st == MFN_ERR_CAPACITY && np > pairs.size()) {
    pairs.resize(np);
    st = mfn_accum_decode_gathered(buf.data(), nseq, sparse.data(), sparse.size(), per ? per->data() : nullptr, pairs.data(), pairs.size(), &np);
  }
  return st;
}

int main(int argc, char **argv) {
  g_t0 = std::chrono::steady_clock::now();
  Globals G;
  std::vector<std::string> err;
  for (int a = 1; a < argc; a++) {
    auto is = [&](const char *f) { return strcmp(argv[a], f) == 0; };
    auto next = [&]() -> const char * { return a + 1 < argc ? argv[++a] : ""; };
    if (is("-sequence")) G.seqName = next();
    else if (is("-seqmers")) G.seqDBname = next();
    else if (is("-readmers")) G.readDBname = next();
    else if (is("-peak")) G.peak = strtod(next(), nullptr);
    else if (is("-prob")) G.prob = next();
    else if (is("-vcf")) G.vcfName = next();
    else if (is("-output")) G.outName = next();
    else if (is("-min")) G.minV = strtoull(next(), nullptr, 10);
    else if (is("-max")) G.maxV = strtoull(next(), nullptr, 10);
    else if (is("-threads")) G.threads = (uint32_t)strtoul(next(), nullptr, 10);
    else if (is("-memory")) next();
    else if (is("-device")) G.device = atoi(next());
    else if (is("-devices")) {
      const char *v = next();
      for (const char *p = v; *p;) {
        char *e2 = nullptr;
        long d = strtol(p, &e2, 10);
        if (e2 == p || d < 0) { err.push_back(std::string("Bad -devices '") + v + "' (want a comma-separated list of device numbers).\n"); break; }
        G.devices.push_back((int)d);
        p = (*e2 == ',') ? e2 + 1 : e2;
        if (*e2 && *e2 != ',') { err.push_back(std::string("Bad -devices '") + v + "'.\n"); break; }
      }
    }
    else if (is("-nofuse")) G.nofuse = true;
    else if (is("-shard")) {
      const char *v = next();
      if (sscanf(v, "%u/%u", &G.shard, &G.nshards) != 2 || G.nshards == 0 || G.shard >= G.nshards) err.push_back(std::string("Bad -shard '") + v + "' (want i/n with i < n).\n");
    }
    else if (is("-nosplit")) G.nosplit = true;
    else if (is("-filter")) G.report = OP_FILTER;
    else if (is("-better")) G.report = OP_BETTER;
    else if (is("-strict")) G.report = OP_STRICT;
    else if (is("-loose")) { fprintf(stderr, "*EXPERIMENTAL* Running in -loose mode\n"); G.report = OP_LOOSE; }
    else if (is("-polish")) G.report = OP_POLISH;
    else if (is("-hist")) G.report = OP_HIST;
    else if (is("-dump")) G.report = OP_DUMP;
    else if (is("-skipMissing")) G.skipMissing = true;
    else if (is("-completeness")) G.report = OP_COMPL;
    else if (is("-comb")) G.comb = (uint32_t)strtoul(next(), nullptr, 10);
    else if (is("-debug")) G.debug = true;
    else err.push_back(std::string("Unknown option '") + argv[a] + "'.\n");
  }
  // the reference's checks (merfin.C:159-181)
  const bool variants = G.report == OP_POLISH || G.report == OP_FILTER || G.report == OP_BETTER || G.report == OP_STRICT || G.report == OP_LOOSE;
  if (G.report != OP_COMPL) {
    if (!G.seqName) err.push_back("No input sequences (-sequence) supplied.\n");
    if (!G.outName) err.push_back("No output (-output) supplied.\n");
  }
  if (variants && !G.vcfName) err.push_back("No variant call input (-vcf) supplied; mandatory for -filter or -polish.\n");
  if (G.report != OP_FILTER && G.peak == 0) err.push_back("No haploid peak (-peak) supplied.\n");
  if (G.report == OP_COMPL && !G.seqName && !G.seqDBname) err.push_back("No sequence meryl database (-seqmers) nor sequence (-sequence) supplied.\n");
  if (G.report == OP_NONE) err.push_back("No report type (-filter, -polish, -hist, -dump, -completeness) supplied.\n");
  if (!G.readDBname) err.push_back("No read meryl database (-readmers) supplied.\n");
  if (!err.empty()) usage(argv[0], err);

  load_Kmetric(G);
  mfn_kmodel km = {G.peak, (uint32_t)G.probK.size(), G.probK.data(), G.probP.data()};

  if (G.devices.empty()) G.devices.push_back(G.device);
  std::vector<Dev> devs(G.devices.size());
  for (size_t d = 0; d < devs.size(); d++) devs[d].id = G.devices[d];
  on_devices(devs, [&](Dev &D, size_t) { D.status = mfn_ctx_create(&D.ctx, D.id); if (D.status != MFN_OK) D.error = mfn_last_error(); });
  if (!devices_ok(devs, "cannot use the GPU")) return 1;
  mark("CUDA contexts created");
  mfn_ctx *ctx = devs[0].ctx;
  const uint32_t ND = (uint32_t)devs.size();

  // ---- -completeness streams both databases (merfin-completeness.C:48-144) ----------------------
  if (G.report == OP_COMPL) {
    fprintf(stderr, "-- Compute completeness.\n");
    std::vector<mfn_pair_count> pairs(1 << 20);
    uint64_t np = 0;
    int st;
    if (G.seqDBname && ND == 1) {
      st = mfn_completeness(ctx, G.readDBname, G.seqDBname, pairs.data(), pairs.size(), &np);
      if (st == MFN_ERR_CAPACITY && np > pairs.size()) { pairs.resize(np); st = mfn_completeness(ctx, G.readDBname, G.seqDBname, pairs.data(), pairs.size(), &np); }
    } else {
      // Every device joins its share of the read database's data files (files [64 d / N, 64 (d + 1) / N):
      // the grain of the reference's OpenMP loop, merfin-completeness.C:65) with its own replica of the raw
      // assembly table; the joint histograms are summed on device 0.  Without -seqmers the reference builds
      // <sequence>.meryl with "meryl count" first (merfin-globals.C:168-189); here the assembly k-mers are
      // counted on the GPU and joined in place.
      mfn_meryl_info_t mi;
      if (mfn_meryl_info(G.readDBname, &mi) != MFN_OK) die("-readmers");
      SeqSet S;
      if (!G.seqDBname) {
        fprintf(stderr, "-- No -seqmer given. Counting the %u-mers of '%s' on the GPU.\n", mi.k, G.seqName);
        std::string e;
        if (!load_sequences(G.seqName, S, e)) { fprintf(stderr, "ERROR: %s\n", e.c_str()); return 1; }
      }
      on_devices(devs, [&](Dev &D, size_t d) {
        int st2;
        if (G.seqDBname) st2 = mfn_table_load(D.ctx, G.seqDBname, 0, 0xffffffffu, &D.A);      // raw values: offset 0
        else {
          mfn_seq *seq = nullptr;
          st2 = mfn_seq_create(D.ctx, S.bases.data(), S.off.data(), (uint32_t)S.nseq(), &seq);
          if (st2 == MFN_OK) st2 = mfn_table_count_seq(D.ctx, mi.k, seq, &D.A);
          mfn_seq_free(seq);
        }
        if (st2 == MFN_OK) st2 = mfn_accum_create(D.ctx, 0, &D.acc);
        if (st2 == MFN_OK) st2 = mfn_completeness_shard(D.ctx, G.readDBname, D.A, (uint32_t)d, ND, D.acc);
        D.status = st2;
        if (st2 != MFN_OK) D.error = mfn_last_error();
      });
      if (!devices_ok(devs, "completeness")) return 1;
      st = reduce_and_fetch(devs, 0, nullptr, pairs, np);
      for (Dev &D : devs) { mfn_accum_free(D.acc); mfn_table_free(D.A); D.acc = nullptr; D.A = nullptr; }
    }
    if (st != MFN_OK) die("completeness");
    double total = 0, undr = 0;
    mfn_completeness_report(pairs.data(), np, &km, &total, &undr);
    fprintf(stderr, "\n");
    fprintf(stderr, "TOTAL readK:   %15.2f\n", total);
    fprintf(stderr, "TOTAL undrcpy:    %15.5f\n", undr);
    fprintf(stderr, "COMPLETENESS:             %0.5f\n", 1.0 - undr / total);
    for (Dev &D : devs) mfn_ctx_destroy(D.ctx);
    fprintf(stderr, "Bye!\n");
    return 0;
  }

  // ---- sequences and tables (merfin-globals.C:114-195) -------------------------------------------
  fprintf(stderr, "-- Opening sequences in '%s'.\n", G.seqName);
  SeqSet S;
  std::string e;
  if (!load_sequences(G.seqName, S, e)) { fprintf(stderr, "ERROR: %s\n", e.c_str()); return 1; }

  mark("sequences loaded");
  // every device loads its own replica of the tables (the databases are decoded on the device)
  mfn_meryl_info_t rmi;
  if (mfn_meryl_info(G.readDBname, &rmi) != MFN_OK) die("-readmers");
  const uint32_t K = rmi.k;
  // -hist does not need the read database resident: it is streamed through the device into the table
  // restricted to the k-mers of each device's share of the sequences (mfn_tables_restrict_stream);
  // the resident tables are loaded only if that is not possible (non-canonical database, odd batch)
  // (k > 32: tables of 128-bit k-mers, loaded whole and scanned directly; nothing to stream into or fuse)
  bool stream_reads = G.report == OP_HIST && !G.nofuse && !getenv("MFN_NO_STREAM") && S.nseq() > 0 && K <= 32;
  fprintf(stderr, "-- Loading kmers from '%s' into lookup table.\n", G.readDBname);
  if (G.seqDBname) fprintf(stderr, "-- Loading kmers from '%s' into lookup table.\n", G.seqDBname);
  else fprintf(stderr, "-- No -seqmer given. Counting the %u-mers of '%s' on the GPU.\n", K, G.seqName);
  auto load_reads = [&](Dev &D) -> int { return mfn_table_load(D.ctx, G.readDBname, (uint32_t)G.minV, (uint32_t)G.maxV, &D.R); };
  on_devices(devs, [&](Dev &D, size_t) {
    int st = stream_reads ? MFN_OK : load_reads(D);
    if (st == MFN_OK) {
      if (G.seqDBname) st = mfn_table_load(D.ctx, G.seqDBname, 0, 0xffffffffu, &D.A);
      else {
        // the reference shells out to "meryl count" here (merfin-globals.C:168-189); we count on the GPU
        mfn_seq *seq = nullptr;
        st = mfn_seq_create(D.ctx, S.bases.data(), S.off.data(), (uint32_t)S.nseq(), &seq);
        if (st == MFN_OK) st = mfn_table_count_seq(D.ctx, K, seq, &D.A);
        mfn_seq_free(seq);
      }
    }
    D.status = st;
    if (st != MFN_OK) D.error = mfn_last_error();
  });
  if (!devices_ok(devs, "loading the databases")) return 1;
  mark("tables loaded");
  mfn_table_info_t ri, ai;
  mfn_table_info(devs[0].A, &ai);
  if (ai.k != K) { fprintf(stderr, "mer size mismatch, can't process this set of files.\n"); return 1; }
  // the fused table of the resident pair (one probe per window instead of two)
  auto fuse_all = [&]() {
    mfn_table_info(devs[0].R, &ri);
    if (!G.nofuse && ri.canonical && ai.canonical) {
      on_devices(devs, [&](Dev &D, size_t) { D.status = mfn_tables_fuse(D.ctx, D.R, D.A); if (D.status != MFN_OK) D.error = mfn_last_error(); });
      for (Dev &D : devs) if (D.status != MFN_OK) { fprintf(stderr, "-- (fused table not built on device %d: %s)\n", D.id, D.error.c_str()); D.status = MFN_OK; }
    }
  };
  // The fused table costs one random insert per k-mer of both databases; a scan of a whole assembly earns that
  // back many times over, a few hundred thousand variant paths do not (decided below, once the VCF is read).
  const bool fuse_later = variants;
  if (!stream_reads && !fuse_later) fuse_all();
  mfn_table *R = devs[0].R, *A = devs[0].A;

  // ---- -hist (merfin-histogram.C) ----------------------------------------------------------------
  if (G.report == OP_HIST) {
    fprintf(stderr, "-- Generate histogram of the k* metric to '%s'.\n", G.outName);
    std::vector<mfn_seq_counts> per(S.nseq() + 1);
    std::vector<mfn_pair_count> pairs(1 << 20);
    uint64_t np = 0;
    int st = MFN_OK;
    bool done = false;
    if (stream_reads) {
      // device d: pack the sequences, stream the read database into the table of tiles [T d / N, T (d + 1) / N),
      // scan those tiles
      on_devices(devs, [&](Dev &D, size_t d) {
        mfn_seq *seq = nullptr;
        mfn_table *T = nullptr;
        int st2 = mfn_seq_create(D.ctx, S.bases.data(), S.off.data(), (uint32_t)S.nseq(), &seq);
        if (d == 0) mark("sequences packed on the device");
        if (st2 == MFN_OK) st2 = mfn_tables_restrict_stream(D.ctx, G.readDBname, (uint32_t)G.minV, (uint32_t)G.maxV, D.A, seq, (uint32_t)d, ND, &T, nullptr);
        if (d == 0) mark("read database streamed into the restricted table");
        if (st2 == MFN_OK) st2 = mfn_accum_create(D.ctx, (uint32_t)S.nseq(), &D.acc);
        if (st2 == MFN_OK) st2 = mfn_hist_resident(D.ctx, T, D.A, seq, &km, (uint32_t)d, ND, D.acc);
        uint64_t np2 = 0;
        if (st2 == MFN_OK) st2 = mfn_accum_fetch(D.ctx, D.acc, 0, 0, nullptr, nullptr, 0, &np2);     // a bin or the side list overflowed?
        if (d == 0) mark("scanned");
        D.status = st2;
        if (st2 != MFN_OK) D.error = mfn_last_error();
        mfn_table_free(T);
        mfn_seq_free(seq);
      });
      bool retry = false;
      for (Dev &D : devs) if (D.status == MFN_ERR_UNSUPPORTED || D.status == MFN_ERR_CAPACITY) retry = true;
      if (!retry) {
        if (!devices_ok(devs, "-hist")) return 1;
        st = reduce_and_fetch(devs, (uint32_t)S.nseq(), &per, pairs, np);
        done = true;
      } else {
        for (Dev &D : devs) if (D.status != MFN_OK) { fprintf(stderr, "-- (read database loaded instead of streamed: %s)\n", D.error.c_str()); break; }
      }
      for (Dev &D : devs) { mfn_accum_free(D.acc); D.acc = nullptr; D.status = MFN_OK; }
      if (!done) {
        on_devices(devs, [&](Dev &D, size_t) { D.status = load_reads(D); if (D.status != MFN_OK) D.error = mfn_last_error(); });
        if (!devices_ok(devs, "-readmers")) return 1;
        fuse_all();
        R = devs[0].R;
      }
    }
    if (!done && ND == 1) {
      st = mfn_hist(ctx, R, A, S.bases.data(), S.off.data(), (uint32_t)S.nseq(), &km, per.data(), pairs.data(), pairs.size(), &np);
      if (st == MFN_ERR_CAPACITY && np > pairs.size()) { pairs.resize(np); st = mfn_hist(ctx, R, A, S.bases.data(), S.off.data(), (uint32_t)S.nseq(), &km, per.data(), pairs.data(), pairs.size(), &np); }
    } else if (!done) {
      // device d scans tiles [T d / N, T (d + 1) / N) of the whole sequence set (uploading only those bases)
      // into its own accumulator; one reduction at the end
      on_devices(devs, [&](Dev &D, size_t d) {
        int st2 = mfn_accum_create(D.ctx, (uint32_t)S.nseq(), &D.acc);
        for (int attempt = 0; st2 == MFN_OK && attempt < 2; attempt++) {
          st2 = mfn_hist_host(D.ctx, D.R, D.A, S.bases.data(), S.off.data(), (uint32_t)S.nseq(), &km, (uint32_t)d, ND, D.acc);
          if (st2 != MFN_OK) break;
          // the range-partitioned scan reports a shard it could not hold through the accumulator: redo it with the direct scan
          uint64_t np2 = 0;
          int sf = mfn_accum_fetch(D.ctx, D.acc, 0, 0, nullptr, nullptr, 0, &np2);
          if (sf != MFN_ERR_CAPACITY || attempt == 1) { st2 = sf; break; }
          mfn_ctx_set_sweep(D.ctx, 0);
          st2 = mfn_accum_reset(D.ctx, D.acc);
        }
        mfn_ctx_set_sweep(D.ctx, -1);
        D.status = st2;
        if (st2 != MFN_OK) D.error = mfn_last_error();
      });
      if (!devices_ok(devs, "-hist")) return 1;
      st = reduce_and_fetch(devs, (uint32_t)S.nseq(), &per, pairs, np);
      for (Dev &D : devs) { mfn_accum_free(D.acc); D.acc = nullptr; }
    }
    if (st != MFN_OK) die("-hist");
    uint64_t cum = 0;
    for (size_t i = 0; i < S.nseq(); i++) {                        // the per-sequence lines of outputHistogram
      cum += per[i].kmissing;
      fprintf(stderr, "%s\t%lu\t%lu\t%lu\t%.2f\n", S.ident[i].c_str(), (unsigned long)per[i].kmissing, (unsigned long)cum,
              (unsigned long)per[i].kasm, histoQV((double)per[i].kmissing, (double)per[i].kasm, K));
    }
    if (S.nseq() > 0) {
      mfn_hist_summary sum;
      std::vector<uint64_t> undr(1 << 16), over(1 << 16);
      st = mfn_hist_report(pairs.data(), np, &km, K, undr.data(), over.data(), (uint32_t)undr.size(), &sum);
      if (st == MFN_ERR_CAPACITY) {
        undr.resize(sum.n_undr + 1); over.resize(sum.n_over + 1);
        st = mfn_hist_report(pairs.data(), np, &km, K, undr.data(), over.data(), (uint32_t)std::min(undr.size(), over.size()), &sum);
      }
      if (st != MFN_OK) die("-hist report");
      uint64_t need = 0;
      mfn_hist_format(undr.data(), sum.n_undr, over.data(), sum.n_over, nullptr, 0, &need);
      std::vector<char> text(need + 1);
      mfn_hist_format(undr.data(), sum.n_undr, over.data(), sum.n_over, text.data(), need + 1, &need);
      OutFile F;
      if (!F.open(G.outName)) { fprintf(stderr, "cannot write '%s'\n", G.outName); return 1; }
      fwrite(text.data(), 1, need, F.f);
      F.close();
      fprintf(stderr, "\n");
      fprintf(stderr, "K-mers not found in reads (missing) : %lu\n", (unsigned long)sum.kmissing);
      fprintf(stderr, "K-mers overly represented in assembly: %.2f\n", sum.koverCpy);
      fprintf(stderr, "K-mers found in the assembly: %lu\n", (unsigned long)sum.kasm);
      fprintf(stderr, "Missing QV: %.2f\n", sum.qv_missing);
      fprintf(stderr, "Merfin QV*: %.2f\n", sum.qv_star);
      fprintf(stderr, "*** Note this QV is valid only if -seqmer was generated with -sequence ***\n\n");
      fprintf(stderr, "*** Missing QV only considers missing kmers as errors. Merfin QV* includes overrepresented kmers. ***\n\n");
      fprintf(stderr, "*** When the lookup table is provided, missing QV includes weighted low frequency kmers, otherwise it is identical to Merqury QV. ***\n\n");
    }
  }

  // ---- -dump (merfin-dump.C) ---------------------------------------------------------------------
  if (G.report == OP_DUMP) {
    fprintf(stderr, "-- Dump per-base k* metric to '%s'.\n", G.outName);
    // records go through the GPU in batches while the previous batch is formatted on the host threads
    // and written in position order (host/dump_driver.cpp)
    OutFile OF;
    if (!G.skipMissing && S.nseq() > 0 && !OF.open(G.outName)) { fprintf(stderr, "cannot write '%s'\n", G.outName); return 1; }
    FILE *F = OF.f;
    DumpBatchFn pairs = [&](const char *bases, const uint64_t *off, uint32_t nrec, uint32_t *vr, uint32_t *va, uint8_t *ok) {
      return mfn_dump(ctx, R, A, bases, off, nrec, vr, va, ok);
    };
    std::string derr;
    if (dump_records(S, &km, F ? fileno(F) : -1, (int)G.threads, 0, pairs, stderr, &derr) != MFN_OK) { fprintf(stderr, "-dump: %s\n", derr.c_str()); return 1; }
    OF.close();
  }

  // ---- variants (merfin-variants.C) --------------------------------------------------------------
  if (variants) {
    fprintf(stderr, "-- Opening vcf file '%s'.\n", G.vcfName);
    mfn_vcf *vcf = nullptr;
    if (mfn_vcf_load(G.vcfName, &vcf) != MFN_OK) { fprintf(stderr, "ERROR: %s\n", mfn_last_error()); return 1; }
    mfn_vcf_info_t vi;
    mfn_vcf_info(vcf, &vi);
    fprintf(stderr, "   Collected %lu header lines.\n", (unsigned long)vi.headers);
    fprintf(stderr, "   Loaded %lu records:\n", (unsigned long)vi.records);
    fprintf(stderr, "      %-8lu unique contig%s\n", (unsigned long)vi.contigs, vi.contigs == 1 ? "" : "s");
    fprintf(stderr, "      %-8u contig IDs\n", (unsigned)vi.contig_ids);
    fprintf(stderr, "   Excluded %lu invalid records\n\n", (unsigned long)vi.excluded);
    fprintf(stderr, "Merge variants within %u-mer bases, splitting combinations greater than %u.\n", K, G.comb);
    if (mfn_vcf_cluster(vcf, K, G.comb, G.nosplit) != MFN_OK) die("-vcf");
    {
      // ~100 path bases (windows) per VCF record against one insert per k-mer of the two tables
      mfn_table_info_t tr, ta;
      mfn_table_info(devs[0].R, &tr);
      mfn_table_info(devs[0].A, &ta);
      if (vi.records * 100 > (tr.n_kmers + ta.n_kmers) / 2) fuse_all();
      mark("VCF loaded and clustered");
    }
    fprintf(stderr, "-- Generate variant mers and score them.\n");

    const int mode = G.report == OP_POLISH ? MFN_MODE_POLISH : G.report == OP_BETTER ? MFN_MODE_BETTER : G.report == OP_STRICT ? MFN_MODE_STRICT
                   : G.report == OP_LOOSE ? MFN_MODE_LOOSE : MFN_MODE_FILTER;
    std::string outPath = std::string(G.outName) + (G.report == OP_POLISH ? ".polish.vcf" : ".filter.vcf");   // merfin-variants.C:324-327 (Q7)
    OutFile OO;
    FILE *O = nullptr, *D = nullptr;
    // Records are independent: path enumeration and selection run on host threads (one record per
    // task, like the reference's sweatShop workers; mfn_variants_process batches the record's probes
    // through the device), and the results are written in record order, so the output does not
    // depend on the thread count.
    const size_t nrec = S.nseq();
    std::vector<char *> results(nrec, nullptr), dbgs(nrec, nullptr);
    std::vector<int> status(nrec, 0);
    unsigned nthreads = G.threads ? G.threads : std::max(1u, std::thread::hardware_concurrency());
    if (G.debug) nthreads = 1;                                   // debug records carry a running id
    if (G.debug && ND > 1) { fprintf(stderr, "-debug runs on one device.\n"); return 1; }
    uint64_t varMerId = 0;
    {
      // A task = one piece of one record's clusters: the pieces of a record are consecutive cluster ranges
      // (mfn_variants_process shard / nshards) whose texts concatenate to the record's result.  With -devices the
      // pieces of a record are dealt to the devices in turn; with -shard i/n only the pieces of range i are run.
      // Records are long and few (24 at the benchmark): one record per task would leave the threads waiting
      // for the biggest one.
      struct Task { size_t rec; uint32_t piece, npieces, shard, nshards, dev; uint64_t weight; char *text = nullptr; };
      std::vector<Task> tasks;
      for (size_t i = 0; i < nrec; i++) {
        const uint64_t nc = mfn_vcf_clusters_of(vcf, S.ident[i].c_str());
        if (nc == 0) continue;
        const uint32_t outer = ND > 1 ? ND : G.nshards;               // ranges that exist before the split into pieces
        const char *pe = getenv("MFN_VARIANT_PIECE");                 // clusters per piece (tests make it small)
        const uint64_t piece = pe && atoll(pe) > 0 ? (uint64_t)atoll(pe) : 50000;
        uint32_t per = G.debug ? 1u : (uint32_t)std::min<uint64_t>(4096, std::max<uint64_t>(1, (nc / outer + piece - 1) / piece));
        for (uint32_t o = 0; o < outer; o++) {
          if (ND == 1 && o != G.shard) continue;
          for (uint32_t p = 0; p < per; p++) {
            Task t;
            t.rec = i; t.piece = o * per + p; t.npieces = outer * per; t.shard = o * per + p; t.nshards = outer * per;
            t.dev = ND > 1 ? o : 0; t.weight = nc / (outer * per);
            tasks.push_back(t);
          }
        }
      }
      std::vector<size_t> order(tasks.size());
      for (size_t i = 0; i < order.size(); i++) order[i] = i;
      nthreads = (unsigned)std::max<size_t>(1, std::min<size_t>(nthreads, tasks.size()));
      if (nthreads > 1) std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return tasks[a].weight > tasks[b].weight; });
      // Scorer lanes: the device part of a scorer call (upload, pack, scan, D2H of the pairs) holds its ctx's lock.  A
      // second context per device over the SAME resident tables (they are only read) lets the calls of different
      // host threads overlap on the device (MFN_LANES, default 2; one with -debug or a single thread).
      size_t nlanes = 2;
      if (const char *le = getenv("MFN_LANES")) nlanes = (size_t)std::max(1, std::min(8, atoi(le)));
      if (nthreads == 1) nlanes = 1;
      std::vector<std::vector<mfn_ctx *>> lanes(devs.size());
      for (size_t d = 0; d < devs.size(); d++) {
        lanes[d].push_back(devs[d].ctx);
        for (size_t l = 1; l < nlanes; l++) {
          mfn_ctx *c = nullptr;
          if (mfn_ctx_create(&c, devs[d].id) != MFN_OK) break;          // (fewer lanes: still correct)
          lanes[d].push_back(c);
        }
      }
      std::atomic<size_t> next{0};
      std::vector<int> tstatus(tasks.size(), MFN_OK);
      auto worker = [&]() {
        for (size_t j; (j = next++) < order.size();) {
          Task &t = tasks[order[j]];
          uint64_t localId = 0;
          Dev &D = devs[t.dev];
          mfn_ctx *lane = lanes[t.dev][j % lanes[t.dev].size()];
          tstatus[order[j]] = mfn_variants_process(lane, D.R, D.A, vcf, S.ident[t.rec].c_str(), S.seq(t.rec), S.length(t.rec), mode, G.comb, &km,
                                                   t.shard, t.nshards, &t.text, nullptr, G.debug ? &dbgs[t.rec] : nullptr, nullptr,
                                                   nthreads == 1 ? &varMerId : &localId);
          if (tstatus[order[j]] != MFN_OK) fprintf(stderr, "scoring failed: %s\n", mfn_last_error());
        }
      };
      std::vector<std::thread> th;
      for (unsigned t = 1; t < nthreads; t++) th.emplace_back(worker);
      worker();
      for (auto &t : th) t.join();
      for (auto &dl : lanes) for (size_t l = 1; l < dl.size(); l++) mfn_ctx_destroy(dl[l]);
      // the pieces of a record, in order, are its result
      std::vector<std::string> joined(nrec);
      std::vector<char> has(nrec, 0);
      for (size_t j = 0; j < tasks.size(); j++) {                 // tasks were created in (record, piece) order
        Task &t = tasks[j];
        if (tstatus[j] != MFN_OK) status[t.rec] = tstatus[j];
        if (t.text) { joined[t.rec] += t.text; has[t.rec] = 1; mfn_text_free(t.text); }
      }
      for (size_t i = 0; i < nrec; i++)
        if (has[i]) { results[i] = (char *)malloc(joined[i].size() + 1); memcpy(results[i], joined[i].c_str(), joined[i].size() + 1); }
    }
    for (size_t i = 0; i < nrec; i++) {
      if (status[i] != 0) return 1;
      if (!results[i]) continue;
      if (!O) {                                                  // outputVariants: header once, then every record's result
        if (OO.open(outPath)) O = OO.f;
        if (!O) { fprintf(stderr, "cannot write '%s'\n", outPath.c_str()); return 1; }
        uint64_t need = 0;
        mfn_vcf_header(vcf, nullptr, 0, &need);
        std::vector<char> hdr(need + 1);
        mfn_vcf_header(vcf, hdr.data(), need, &need);
        if (G.shard == 0) fwrite(hdr.data(), 1, need, O);        // shards > 0 write records only: `cat` of the shard files is the whole output
      }
      fputs(results[i], O);
      mfn_text_free(results[i]);
      if (G.debug && dbgs[i] && dbgs[i][0]) {
        if (!D) { std::string cmd = "gzip -1c > " + shell_quote(std::string(G.outName) + ".00.debug.gz"); D = popen(cmd.c_str(), "w"); }
        if (D) fputs(dbgs[i], D);
      }
      mfn_text_free(dbgs[i]);
    }
    OO.close();
    if (D) pclose(D);
    mfn_vcf_free(vcf);
  }

  mark("report written");
  for (Dev &D : devs) {
    mfn_table_free(D.R);
    mfn_table_free(D.A);
    mfn_ctx_destroy(D.ctx);
  }
  mark("device memory released");
  fprintf(stderr, "Bye!\n");
  return 0;
}
