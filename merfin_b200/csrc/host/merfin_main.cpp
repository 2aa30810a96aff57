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

int main(int argc, char **argv) {
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

  mfn_ctx *ctx = nullptr;
  if (mfn_ctx_create(&ctx, G.device) != MFN_OK) die("cannot use the GPU");

  // ---- -completeness streams both databases (merfin-completeness.C:48-144) ----------------------
  if (G.report == OP_COMPL) {
    fprintf(stderr, "-- Compute completeness.\n");
    std::vector<mfn_pair_count> pairs(1 << 20);
    uint64_t np = 0;
    int st;
    if (G.seqDBname) {
      st = mfn_completeness(ctx, G.readDBname, G.seqDBname, pairs.data(), pairs.size(), &np);
      if (st == MFN_ERR_CAPACITY && np > pairs.size()) { pairs.resize(np); st = mfn_completeness(ctx, G.readDBname, G.seqDBname, pairs.data(), pairs.size(), &np); }
    } else {
      // the reference builds <sequence>.meryl with "meryl count" first (merfin-globals.C:168-189) and
      // streams that; here the assembly k-mers are counted on the GPU and joined in place
      mfn_meryl_info_t mi;
      if (mfn_meryl_info(G.readDBname, &mi) != MFN_OK) die("-readmers");
      fprintf(stderr, "-- No -seqmer given. Counting the %u-mers of '%s' on the GPU.\n", mi.k, G.seqName);
      SeqSet S;
      std::string e;
      if (!load_sequences(G.seqName, S, e)) { fprintf(stderr, "ERROR: %s\n", e.c_str()); return 1; }
      mfn_seq *seq = nullptr;
      mfn_table *A = nullptr;
      mfn_accum *acc = nullptr;
      if (mfn_seq_create(ctx, S.bases.data(), S.off.data(), (uint32_t)S.nseq(), &seq) != MFN_OK) die("-sequence");
      if (mfn_table_count_seq(ctx, mi.k, seq, &A) != MFN_OK) die("counting -sequence");
      mfn_seq_free(seq);
      st = mfn_accum_create(ctx, 0, &acc);
      if (st == MFN_OK) st = mfn_completeness_shard(ctx, G.readDBname, A, 0, 1, acc);
      if (st == MFN_OK) st = mfn_accum_fetch(ctx, acc, 1, 1, nullptr, pairs.data(), pairs.size(), &np);
      if (st == MFN_ERR_CAPACITY && np > pairs.size()) { pairs.resize(np); st = mfn_accum_fetch(ctx, acc, 1, 1, nullptr, pairs.data(), pairs.size(), &np); }
      mfn_accum_free(acc);
      mfn_table_free(A);
    }
    if (st != MFN_OK) die("completeness");
    double total = 0, undr = 0;
    mfn_completeness_report(pairs.data(), np, &km, &total, &undr);
    fprintf(stderr, "\n");
    fprintf(stderr, "TOTAL readK:   %15.2f\n", total);
    fprintf(stderr, "TOTAL undrcpy:    %15.5f\n", undr);
    fprintf(stderr, "COMPLETENESS:             %0.5f\n", 1.0 - undr / total);
    mfn_ctx_destroy(ctx);
    fprintf(stderr, "Bye!\n");
    return 0;
  }

  // ---- sequences and tables (merfin-globals.C:114-195) -------------------------------------------
  fprintf(stderr, "-- Opening sequences in '%s'.\n", G.seqName);
  SeqSet S;
  std::string e;
  if (!load_sequences(G.seqName, S, e)) { fprintf(stderr, "ERROR: %s\n", e.c_str()); return 1; }

  mfn_table *R = nullptr, *A = nullptr;
  fprintf(stderr, "-- Loading kmers from '%s' into lookup table.\n", G.readDBname);
  if (mfn_table_load(ctx, G.readDBname, (uint32_t)G.minV, (uint32_t)G.maxV, &R) != MFN_OK) die("-readmers");
  mfn_table_info_t ri;
  mfn_table_info(R, &ri);
  const uint32_t K = ri.k;
  if (G.seqDBname) {
    fprintf(stderr, "-- Loading kmers from '%s' into lookup table.\n", G.seqDBname);
    if (mfn_table_load(ctx, G.seqDBname, 0, 0xffffffffu, &A) != MFN_OK) die("-seqmers");
  } else {
    // the reference shells out to "meryl count" here (merfin-globals.C:168-189); we count on the GPU
    fprintf(stderr, "-- No -seqmer given. Counting the %u-mers of '%s' on the GPU.\n", K, G.seqName);
    mfn_seq *seq = nullptr;
    if (mfn_seq_create(ctx, S.bases.data(), S.off.data(), (uint32_t)S.nseq(), &seq) != MFN_OK) die("-sequence");
    if (mfn_table_count_seq(ctx, K, seq, &A) != MFN_OK) die("counting -sequence");
    mfn_seq_free(seq);
  }
  mfn_table_info_t ai;
  mfn_table_info(A, &ai);
  if (ai.k != K) { fprintf(stderr, "mer size mismatch, can't process this set of files.\n"); return 1; }
  if (!G.nofuse && ri.canonical && ai.canonical && mfn_tables_fuse(ctx, R, A) != MFN_OK)
    fprintf(stderr, "-- (fused table not built: %s)\n", mfn_last_error());

  // ---- -hist (merfin-histogram.C) ----------------------------------------------------------------
  if (G.report == OP_HIST) {
    fprintf(stderr, "-- Generate histogram of the k* metric to '%s'.\n", G.outName);
    std::vector<mfn_seq_counts> per(S.nseq() + 1);
    std::vector<mfn_pair_count> pairs(1 << 20);
    uint64_t np = 0;
    int st = mfn_hist(ctx, R, A, S.bases.data(), S.off.data(), (uint32_t)S.nseq(), &km, per.data(), pairs.data(), pairs.size(), &np);
    if (st == MFN_ERR_CAPACITY && np > pairs.size()) { pairs.resize(np); st = mfn_hist(ctx, R, A, S.bases.data(), S.off.data(), (uint32_t)S.nseq(), &km, per.data(), pairs.data(), pairs.size(), &np); }
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
      FILE *F = fopen(G.outName, "w");
      if (!F) { fprintf(stderr, "cannot write '%s'\n", G.outName); return 1; }
      fwrite(text.data(), 1, need, F);
      fclose(F);
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
    FILE *F = nullptr;
    if (!G.skipMissing && S.nseq() > 0) { F = fopen(G.outName, "w"); if (!F) { fprintf(stderr, "cannot write '%s'\n", G.outName); return 1; } }
    DumpBatchFn pairs = [&](const char *bases, const uint64_t *off, uint32_t nrec, uint32_t *vr, uint32_t *va, uint8_t *ok) {
      return mfn_dump(ctx, R, A, bases, off, nrec, vr, va, ok);
    };
    std::string derr;
    if (dump_records(S, &km, F ? fileno(F) : -1, (int)G.threads, 0, pairs, stderr, &derr) != MFN_OK) { fprintf(stderr, "-dump: %s\n", derr.c_str()); return 1; }
    if (F) fclose(F);
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
    fprintf(stderr, "-- Generate variant mers and score them.\n");

    const int mode = G.report == OP_POLISH ? MFN_MODE_POLISH : G.report == OP_BETTER ? MFN_MODE_BETTER : G.report == OP_STRICT ? MFN_MODE_STRICT
                   : G.report == OP_LOOSE ? MFN_MODE_LOOSE : MFN_MODE_FILTER;
    std::string outPath = std::string(G.outName) + (G.report == OP_POLISH ? ".polish.vcf" : ".filter.vcf");   // merfin-variants.C:324-327 (Q7)
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
    nthreads = (unsigned)std::max<size_t>(1, std::min<size_t>(nthreads, nrec));
    uint64_t varMerId = 0;
    {
      // biggest records first (they bound the makespan)
      std::vector<size_t> order(nrec);
      for (size_t i = 0; i < nrec; i++) order[i] = i;
      if (nthreads > 1) std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return S.length(a) > S.length(b); });
      std::atomic<size_t> next{0};
      auto worker = [&]() {
        for (size_t j; (j = next++) < nrec;) {
          const size_t i = order[j];
          if (mfn_vcf_clusters_of(vcf, S.ident[i].c_str()) == 0) continue;
          uint64_t localId = 0;
          status[i] = mfn_variants_process(ctx, R, A, vcf, S.ident[i].c_str(), S.seq(i), S.length(i), mode, G.comb, &km, G.shard, G.nshards,
                                           &results[i], nullptr, G.debug ? &dbgs[i] : nullptr, nullptr, nthreads == 1 ? &varMerId : &localId);
          if (status[i] != MFN_OK) fprintf(stderr, "scoring failed: %s\n", mfn_last_error());
        }
      };
      std::vector<std::thread> th;
      for (unsigned t = 1; t < nthreads; t++) th.emplace_back(worker);
      worker();
      for (auto &t : th) t.join();
    }
    for (size_t i = 0; i < nrec; i++) {
      if (status[i] != 0) return 1;
      if (!results[i]) continue;
      if (!O) {                                                  // outputVariants: header once, then every record's result
        O = fopen(outPath.c_str(), "w");
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
    if (O) fclose(O);
    if (D) pclose(D);
    mfn_vcf_free(vcf);
  }

  mfn_table_free(R);
  mfn_table_free(A);
  mfn_ctx_destroy(ctx);
  fprintf(stderr, "Bye!\n");
  return 0;
}
