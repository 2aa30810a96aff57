// Test harness: the product's variant front-end (merfin_b200/csrc/host/variants.cpp) driven with
// the CPU oracle as the path scorer, so the host-side control logic (VCF parsing, clustering,
// path enumeration, selectors) can be checked against the reference's golden VCFs without a GPU.
// Test infrastructure only.
//   variant_harness <asm.fa> <reads.meryl> <asm.meryl> <calls.vcf> <mode> <peak> [prob.csv] [--debug] [--comb N] [--nosplit] [--shards N] [--batch B]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../merfin_b200/csrc/host/fasta.h"
#include "../merfin_b200/csrc/host/variants.h"
#include "../oracle/mfn_oracle.h"

int main(int argc, char **argv) {
  if (argc < 7) return 2;
  SeqSet S; std::string err;
  if (!load_sequences(argv[1], S, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
  char eb[256];
  o_db *rd = o_db_load(argv[2], eb, sizeof eb), *ad = o_db_load(argv[3], eb, sizeof eb);
  if (!rd || !ad) { fprintf(stderr, "%s\n", eb); return 1; }
  o_lookup *R = o_lookup_build(rd, 0, 0xffffffffu), *A = o_lookup_build(ad, 0, 0xffffffffu);
  mfnv::VcfFile vcf;
  if (!vcf.load(argv[4], err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
  std::string mode = argv[5];
  std::vector<uint32_t> pk; std::vector<double> pp;
  bool debug = false, nosplit = false;
  uint32_t comb = 15;
  size_t shards = 1, batch = 0;      // --shards N: every record in N consecutive cluster ranges; --batch B: B clusters per scorer call
  for (int a = 7; a < argc; a++) {
    if (!strcmp(argv[a], "--debug")) { debug = true; continue; }
    if (!strcmp(argv[a], "--nosplit")) { nosplit = true; continue; }
    if (!strcmp(argv[a], "--comb") && a + 1 < argc) { comb = (uint32_t)atoi(argv[++a]); continue; }
    if (!strcmp(argv[a], "--shards") && a + 1 < argc) { shards = (size_t)atoi(argv[++a]); continue; }
    if (!strcmp(argv[a], "--batch") && a + 1 < argc) { batch = (size_t)atoi(argv[++a]); continue; }
    FILE *F = fopen(argv[a], "r"); char line[256];
    while (F && fgets(line, sizeof line, F)) {
      unsigned k; double p; char extra;
      if (sscanf(line, "%u,%lf%c", &k, &p, &extra) >= 2 && strchr(line, ',') == strrchr(line, ',')) { pk.push_back(k); pp.push_back(p); }
    }
    if (F) fclose(F);
  }
  mfnv::Options opt;
  opt.k = rd->k; opt.comb = comb; opt.debug = debug;
  opt.km = mfn_kmodel{atof(argv[6]), (uint32_t)pk.size(), pk.data(), pp.data()};
  opt.mode = mode == "polish" ? mfnv::POLISH : mode == "better" ? mfnv::BETTER : mode == "strict" ? mfnv::STRICT : mode == "loose" ? mfnv::LOOSE : mfnv::FILTER;
  vcf.mergeChrPosGT(opt.k, opt.comb, nosplit);
  mfnv::PathScorer scorer = [&](const std::string &paths, const std::vector<uint64_t> &off, mfnv::ScoreOut &out) -> int {
    out.vr_store.assign(paths.size() + 1, 0); out.va_store.assign(paths.size() + 1, 0); out.ok_store.assign(paths.size() + 1, 0);
    for (size_t i = 0; i + 1 < off.size(); i++)
      o_pairs(R, A, paths.data() + off[i], off[i + 1] - off[i], out.vr_store.data() + off[i], out.va_store.data() + off[i], out.ok_store.data() + off[i]);
    out.v_read = out.vr_store.data(); out.v_asm = out.va_store.data(); out.valid = out.ok_store.data();
    return 0;
  };
  for (const std::string &h : vcf.headers) printf("%s\n", h.c_str());
  uint64_t id = 0;
  std::string dbg_all;
  for (size_t i = 0; i < S.nseq(); i++) {
    auto it = vcf.byChr.find(S.ident[i]);
    if (it == vcf.byChr.end()) continue;
    std::string result, dbg;
    const size_t nc = it->second.size();
    for (size_t sh = 0; sh < shards; sh++)
      if (mfnv::process_sequence(S.ident[i], S.seq(i), S.length(i), it->second, opt, scorer, result, debug ? &dbg : nullptr, id,
                                 nc * sh / shards, nc * (sh + 1) / shards, batch)) return 1;
    fputs(result.c_str(), stdout);
    dbg_all += dbg;
  }
  if (debug) { FILE *D = fopen("harness.debug", "w"); fputs(dbg_all.c_str(), D); fclose(D); }
  return 0;
}
