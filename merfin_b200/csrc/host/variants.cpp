// See variants.h.  Host-only control logic; every behaviour that is visible in merfin's output is
// cited against the reference line it reproduces.
#include "variants.h"

#include <thread>

#include "fasta.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <list>

namespace mfnv {

// strings.C:188-249: runs of separators collapse, empty words do not exist
typedef std::pair<const char *, size_t> Span;
static void split_spans(const char *s, size_t n, const char *seps, std::vector<Span> &out) {
  out.clear();
  size_t i = 0;
  while (i < n) {
    while (i < n && strchr(seps, s[i])) i++;
    size_t j = i;
    while (j < n && !strchr(seps, s[j])) j++;
    if (j > i) out.emplace_back(s + i, j - i);
    i = j;
  }
}
std::string VcfRecord::save() const {
  char num[64];
  snprintf(num, sizeof num, "%.1f", qual);
  return chr + "\t" + std::to_string((int)pos) + "\t" + id + "\t" + ref + "\t" + alts + "\t" + num + "\t" + filter +
         "\t" + info + "\t" + formats + "\t" + samples + "\n";
}

VcfFile::~VcfFile() {
  for (auto *r : records) delete r;
  for (auto *a : alleles) delete a;
  for (auto &kv : byChr) for (auto *p : kv.second) delete p;
}

// vcf.C:23-87
static GtAllele *make_allele(const VcfRecord *r) {
  GtAllele *g = new GtAllele();
  g->record = r;
  g->pos = r->pos - 1;
  g->refLen = (uint32_t)r->ref.size();
  g->qual = r->qual;
  const std::string &gt = r->sample_fields.empty() ? std::string() : r->sample_fields[0];
  if (gt.compare(0, 3, "./.") == 0 || gt.compare(0, 3, "0/0") == 0) return g;      // no alleles at all (Q10)
  g->alleles.reserve(3);
  g->alleles.push_back(r->ref);
  int used[8];                                        // ALT column entries taken so far (a sample names a handful at most)
  size_t nused = 0;
  std::vector<int> used_more;
  const char *p = gt.data(), *const e = p + gt.size();
  while (p < e) {
    while (p < e && (*p == '|' || *p == '/')) p++;
    const char *q = p;
    while (q < e && *q != '|' && *q != '/') q++;
    if (q == p) break;
    char tok[24];
    const size_t tl = std::min<size_t>((size_t)(q - p), sizeof tok - 1);
    memcpy(tok, p, tl);
    tok[tl] = 0;
    p = q;
    long idx = strtol(tok, nullptr, 10);
    if (idx <= 0) continue;
    if ((size_t)idx > r->alt_list.size()) continue;
    bool seen = std::find(used, used + nused, (int)idx) != used + nused ||
                std::find(used_more.begin(), used_more.end(), (int)idx) != used_more.end();
    if (seen) continue;                                                             // same ALT column entry again
    if (nused < 8) used[nused++] = (int)idx; else used_more.push_back((int)idx);
    const std::string &hap = r->alt_list[idx - 1];
    if (hap == r->ref) continue;
    g->alleles.push_back(hap);
  }
  return g;
}

// vcf.C:93-151.  Lines are independent, so they are parsed on host threads (contiguous ranges of
// lines per thread) and appended in file order afterwards: the order inside byChr is what the
// reference's (unstable) sort starts from, so it must be the file order.
bool VcfFile::load(const std::string &path, std::string &err) {
  std::vector<char> data;
  if (!slurp_file(path, data, err)) return false;
  const size_t N = data.size();
  std::vector<size_t> starts;                       // first byte of every line
  for (size_t p = 0; p < N;) {
    starts.push_back(p);
    const char *nl = (const char *)memchr(data.data() + p, '\n', N - p);
    p = nl ? (size_t)(nl - data.data()) + 1 : N;
  }
  starts.push_back(N);
  const size_t nlines = starts.size() - 1;
  struct Item { std::string header; VcfRecord *r = nullptr; GtAllele *g = nullptr; PosGT *pg = nullptr; bool excluded = false; };
  unsigned T = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  if (nlines < 20000) T = 1;
  if (const char *env = getenv("MFN_VCF_THREADS")) T = (unsigned)std::max(1, atoi(env));     // (tests force the threaded path on small files)
  T = (unsigned)std::max<size_t>(1, std::min<size_t>(T, nlines));
  std::vector<std::vector<Item>> parts(T);
  auto parse = [&](unsigned t) {
    const size_t l0 = nlines * t / T, l1 = nlines * (t + 1) / T;
    std::vector<Item> &out = parts[t];
    out.reserve(l1 - l0);
    std::vector<Span> w, sub;                        // reused for every line of this thread
    w.reserve(16); sub.reserve(8);
    for (size_t l = l0; l < l1; l++) {
      size_t p = starts[l], e = starts[l + 1];
      while (e > p && (data[e - 1] == '\r' || data[e - 1] == '\n')) e--;
      if (e == p) continue;
      const char *line = data.data() + p;
      const size_t len = e - p;
      Item it;
      if (line[0] == '#') { it.header.assign(line, len); out.push_back(std::move(it)); continue; }
      split_spans(line, len, "\t", w);
      if (w.size() < 10) { it.excluded = true; out.push_back(std::move(it)); continue; }   // vcfRecord.H:53-56
      auto str = [&](int i) { return std::string(w[i].first, w[i].second); };
      VcfRecord *r = new VcfRecord();
      char num[64];
      auto cstr = [&](int i) { const size_t n = std::min<size_t>(w[i].second, sizeof num - 1); memcpy(num, w[i].first, n); num[n] = 0; return num; };
      r->chr = str(0); r->pos = (uint32_t)strtoul(cstr(1), nullptr, 10); r->id = str(2); r->ref = str(3); r->alts = str(4);
      r->qual = strtod(cstr(5), nullptr); r->filter = str(6); r->info = str(7); r->formats = str(8); r->samples = str(9);
      split_spans(r->alts.data(), r->alts.size(), ",", sub);
      r->alt_list.reserve(sub.size());
      for (const Span &x : sub) r->alt_list.emplace_back(x.first, x.second);
      split_spans(r->samples.data(), r->samples.size(), ":", sub);
      r->sample_fields.reserve(sub.size());
      for (const Span &x : sub) r->sample_fields.emplace_back(x.first, x.second);
      GtAllele *g = make_allele(r);
      PosGT *pg = new PosGT();
      pg->chr = r->chr; pg->rStart = g->pos; pg->rEnd = g->pos + g->refLen; pg->gts.push_back(g);
      it.r = r; it.g = g; it.pg = pg;
      out.push_back(std::move(it));
    }
  };
  {
    std::vector<std::thread> th;
    for (unsigned t = 1; t < T; t++) th.emplace_back(parse, t);
    parse(0);
    for (auto &x : th) x.join();
  }
  for (auto &part : parts)
    for (Item &it : part) {
      if (it.excluded) { excluded++; continue; }
      if (!it.r) {
        if (it.header.compare(0, 12, "##contig=<ID") == 0) numChr++;
        headers.push_back(std::move(it.header));
        continue;
      }
      records.push_back(it.r);
      alleles.push_back(it.g);
      byChr[it.r->chr].push_back(it.pg);
    }
  return true;
}

// vcf.C:156-246.  std::sort with the reference's comparator on the reference's initial (file) order.
void VcfFile::mergeChrPosGT(uint32_t ksize, uint32_t comb, bool nosplit) {
  const uint32_t K_OFFSET = 2 * ksize;
  for (auto &kv : byChr) {
    std::vector<PosGT *> &in = kv.second;
    std::vector<PosGT *> out;
    uint32_t removed = 0, nsplit = 0, merged = 0;
    std::sort(in.begin(), in.end(), [](PosGT *const &A, PosGT *const &B) { return A->rStart < B->rStart; });
    out.push_back(in[0]);
    std::vector<PosGT *> dead;
    for (size_t ii = 1; ii < in.size(); ii++) {
      if (in[ii]->gts.empty()) { removed++; dead.push_back(in[ii]); continue; }
      const bool overlapping = in[ii]->rStart < out.back()->rEnd + K_OFFSET;
      const bool toomany = out.back()->gts.size() >= comb;
      if (!overlapping) { out.push_back(in[ii]); continue; }
      if (toomany && !nosplit) { out.push_back(in[ii]); nsplit++; continue; }
      GtAllele *g = in[ii]->gts[0];
      out.back()->gts.push_back(g);
      out.back()->rStart = std::min(out.back()->rStart, g->pos);
      out.back()->rEnd = std::max(out.back()->rEnd, g->pos + g->refLen);
      merged++;
      dead.push_back(in[ii]);
    }
    if (!getenv("MFN_QUIET")) {
      fprintf(stderr, "%s : Reduced %lu variants down to %lu combinations for evaluation:\n", kv.first.c_str(),
              (unsigned long)in.size(), (unsigned long)out.size());
      if (removed > 0) fprintf(stderr, "%s :   Removed %u empty alleles.\n", kv.first.c_str(), removed);
      if (nsplit > 0) fprintf(stderr, "%s :   Split   %u complicated combinations.\n", kv.first.c_str(), nsplit);
      if (merged > 0) fprintf(stderr, "%s :   Merged  %u variants into combinations.\n", kv.first.c_str(), merged);
    }
    for (auto *d : dead) delete d;
    in.swap(out);
  }
}

// ---------------------------------------------------------------------------------------------
// One cluster: its candidate paths and their scores (the reference's varMer object)
// ---------------------------------------------------------------------------------------------
struct VarMer {
  PosGT *posGt = nullptr;
  std::vector<std::string> seqs;
  std::vector<std::vector<int>> gtPaths;
  std::vector<std::vector<uint32_t>> idxPaths, lenPaths;
  std::vector<uint32_t> numMs;
  std::vector<std::vector<double>> kstrs, dkstrs;   // per window; kept only for the -debug statistics
  std::vector<double> totdKs;                        // sum of a path's dkstrs, added up in window order

  void addSeqPath(const std::string &seq, const std::vector<int> &path, const std::vector<uint32_t> &idx,
                  const std::vector<uint32_t> &len) {                                // varMer.C:37-45
    if (std::find(seqs.begin(), seqs.end(), seq) != seqs.end()) return;
    seqs.push_back(seq); gtPaths.push_back(path); idxPaths.push_back(idx); lenPaths.push_back(len);
  }
  double totdK(int i) const { return totdKs[i]; }                                                    // :646-659 (same additions, same order)
  double minAbsK(int i) const {                                                                      // :553-569
    double m = DBL_MAX;
    for (double a : kstrs[i]) { if (a < 0) continue; if (a < m) m = a; }
    return m == DBL_MAX ? -1 : m;
  }
  double maxAbsK(int i) const { double m = -2; for (double a : kstrs[i]) if (a > m) m = a; return m; }  // :572-584
  double avgAbsK(int i) const {                                                                      // :586-604
    double s = 0;
    for (double a : kstrs[i]) if (a >= 0) s += a;
    if (kstrs[i].size() == numMs[i]) return -1;
    return s / (kstrs[i].size() - numMs[i]);
  }
  double medAbsK(int i) const {                                                                      // :606-622
    std::vector<double> k = kstrs[i];
    std::sort(k.begin(), k.end());
    size_t j = 0;
    for (; j < k.size(); j++) if (k[j] >= 0) break;
    if (j == k.size()) return -1;
    return k[j + (k.size() - j) / 2];
  }
  std::string homRecord(int idx) const;
  std::string hetRecord(int idx1, int idx2) const;
};

static std::string qual_str(double q) { return std::to_string((int)q); }                // varMer.C:487,534

std::string VarMer::homRecord(int idx) const {                                          // varMer.C:531-550
  std::string out;
  for (size_t i = 0; i < gtPaths[idx].size(); i++) {
    int alt = gtPaths[idx][i];
    if (alt <= 0) continue;
    const GtAllele *g = posGt->gts[i];
    out += posGt->chr + "\t" + std::to_string(g->pos + 1) + "\t.\t" + g->alleles[0] + "\t" + g->alleles[alt] + "\t" +
           qual_str(g->qual) + "\tPASS\t.\tGT\t1/1\n";
  }
  return out;
}

std::string VarMer::hetRecord(int idx1, int idx2) const {                               // varMer.C:472-529
  std::string out;
  for (size_t i = 0; i < gtPaths[idx1].size(); i++) {
    int a1 = gtPaths[idx1][i], a2 = gtPaths[idx2][i];
    if (a1 + a2 <= 0) continue;
    const GtAllele *g = posGt->gts[i];
    out += posGt->chr + "\t" + std::to_string(g->pos + 1) + "\t.\t" + g->alleles[0] + "\t";
    if (a1 == a2)               out += g->alleles[a1] + "\t" + qual_str(g->qual) + "\tPASS\t.\tGT\t1/1\n";
    else if (a1 == 0 && a2 > 0) out += g->alleles[a2] + "\t" + qual_str(g->qual) + "\tPASS\t.\tGT\t0/1\n";
    else if (a1 > 0 && a2 > 0)  out += g->alleles[a1] + "," + g->alleles[a2] + "\t" + qual_str(g->qual) + "\tPASS\t.\tGT\t1/2\n";
    else if (a1 > 0 && a2 == 0) out += g->alleles[a1] + "\t" + qual_str(g->qual) + "\tPASS\t.\tGT\t1/0\n";
  }
  return out;
}

// merfin-variants.C:22-126.  idxList is shared down the recursion (shifted by the length change of
// every ALT taken so far and restored afterwards); lenList is private to each level.
static void traverse(uint32_t idx, std::vector<uint32_t> &idxList, std::vector<uint32_t> lenList,
                     const std::vector<const std::vector<std::string> *> &haps, const std::string &candidate,
                     std::vector<int> &path, VarMer &vm) {
  const std::vector<std::string> &alts = *haps[idx];
  const uint32_t refLen = lenList[idx];
  const uint32_t last = (uint32_t)idxList.size() - 1;
  for (size_t j = 0; j < alts.size(); j++) {
    path.push_back((int)j);
    std::string changed;                                          // the candidate with this ALT put in (the REF allele keeps `candidate`)
    const std::string &replaced = j > 0 ? changed : candidate;
    int skipped = 0, delta = 0;
    bool overlaps = false, done = false;
    if (j > 0) {
      lenList[idx] = refLen;
      const std::string &hap = alts[j];
      changed = candidate;
      changed.replace(idxList[idx], lenList[idx], hap);           // throws like std::string::replace if out of range
      delta = (int)hap.size() - (int)lenList[idx];
      const uint32_t affected = idxList[idx] + lenList[idx];
      lenList[idx] = (uint32_t)hap.size();
      for (uint32_t i = idx + 1; i < idxList.size(); i++) {       // later variants inside the replaced span keep REF
        if (idxList[i] >= affected) break;
        overlaps = true; idx++; path.push_back(0); skipped++;
      }
      if (overlaps && idx == last) {
        vm.addSeqPath(replaced, path, idxList, lenList);
        done = true;
      } else {
        for (uint32_t i = idx + 1; i < idxList.size(); i++) idxList[i] += delta;
      }
    }
    if (!done) {
      if (idx + 1 < idxList.size()) traverse(idx + 1, idxList, lenList, haps, replaced, path, vm);
      if (idx == last) vm.addSeqPath(replaced, path, idxList, lenList);
      for (uint32_t i = idx + 1; i < idxList.size(); i++) idxList[i] -= delta;
    }
    for (int s = 0; s < skipped; s++) { path.pop_back(); idx--; }
    path.pop_back();
  }
}

// merfinGlobal::getK (merfin-globals.C:66-98) per read multiplicity, tabulated once per call of
// process_sequence: the replay below asks for it once per window of every path
struct KTable {
  static constexpr uint32_t N = 4096;
  std::vector<double> rk, pr;
  const mfn_kmodel *km;
  explicit KTable(const mfn_kmodel *m) : rk(N), pr(N), km(m) {
    for (uint32_t v = 0; v < N; v++) { double a; mfn_getK(m, v, 0, &rk[v], &a, &pr[v]); }
  }
  inline void get(uint32_t vr, double &readK, double &prob) const {
    if (vr < N) { readK = rk[vr]; prob = pr[vr]; }
    else { double a; mfn_getK(km, vr, 0, &readK, &a, &prob); }
  }
};

// varMer.C:48-145 from the integer pairs of every window of every path
static void score(VarMer &vm, const Options &opt, const KTable &kt, const uint32_t *vr, const uint32_t *va, const uint8_t *ok,
                  const std::vector<uint64_t> &off, size_t first_path, double &prob) {
  const uint32_t k = opt.k;
  for (size_t ii = 0; ii < vm.seqs.size(); ii++) {
    const size_t L = vm.seqs[ii].size();
    const uint64_t base = off[first_path + ii];
    uint32_t numM = 0, idx = 0;
    std::vector<double> ks, dks;
    double totd = 0;
    const size_t nwin = L >= k ? L - k + 1 : 0;
    const bool keep = opt.debug;                     // the per-window values are only read by the -debug statistics
    if (opt.mode != FILTER && keep) { ks.reserve(nwin); dks.reserve(nwin); }
    const size_t nvar = vm.idxPaths[ii].size();
    const uint32_t *ipath = vm.idxPaths[ii].data(), *lpath = vm.lenPaths[ii].data();
    const int *gpath = vm.gtPaths[ii].data();
    for (size_t w = 0; w < nwin; w++) {
      double readK = 0, asmK = 0;
      if (ok[base + w]) { kt.get(vr[base + w], readK, prob); asmK = (double)va[base + w]; }
      if (readK == 0) numM++;
      if (opt.mode == FILTER) { idx++; continue; }
      const double oD = fabs(readK - asmK) * prob;
      for (size_t jj = 0; jj < nvar; jj++) {
        const uint32_t ip = ipath[jj], lp = lpath[jj];
        // unsigned arithmetic as in the reference: ip + 1 - k wraps when ip < k - 1 (Q5)
        if (gpath[jj] > 0 && (uint32_t)(ip + 1 - k) <= idx && idx < (uint32_t)(ip + lp + k)) { asmK++; break; }
      }
      double kMetric;
      if (readK == 0) kMetric = -1;
      else if (readK > asmK) kMetric = readK / asmK - 1;
      else kMetric = asmK / readK - 1;
      const double nD = fabs(readK - asmK) * prob;
      totd += oD - nD;
      if (keep) { ks.push_back(kMetric); dks.push_back(oD - nD); }
      idx++;
    }
    vm.numMs.push_back(numM); vm.totdKs.push_back(totd);
    if (keep) { vm.kstrs.push_back(std::move(ks)); vm.dkstrs.push_back(std::move(dks)); }
  }
}

// ---- selectors (varMer.C:150-467) -------------------------------------------------------------
static void pick_min_missing(const VarMer &vm, bool le_ref, std::vector<int> &idxs) {   // shared by better/strict/loose
  const uint32_t refMissing = vm.numMs[0];
  uint32_t numMissing = refMissing;
  for (size_t ii = 0; ii < vm.numMs.size(); ii++) {
    const uint32_t m = vm.numMs[ii];
    if (m < numMissing) { numMissing = m; idxs.clear(); idxs.push_back((int)ii); }
    else if (m == numMissing && (le_ref ? m <= refMissing : m < refMissing)) idxs.push_back((int)ii);
  }
}

static std::string longest_of(const VarMer &vm, const std::vector<int> &idxs) {         // varMer.C:229-246,288-308
  int idx = idxs[0];
  if (idxs.size() > 1) {
    uint32_t best = (uint32_t)vm.seqs[idx].size();
    for (size_t ii = 1; ii < idxs.size(); ii++) {
      uint32_t len = (uint32_t)vm.seqs[idxs[ii]].size();
      if (len > best) { best = len; idx = idxs[ii]; }
    }
  }
  return vm.homRecord(idx);
}

static std::string better_or_strict(const VarMer &vm) {                                 // varMer.C:199-309
  if (vm.numMs.empty()) return "";
  std::vector<int> idxs;
  pick_min_missing(vm, false, idxs);
  if (idxs.empty()) return "";
  return longest_of(vm, idxs);
}

static std::string loose(const VarMer &vm) {                                            // varMer.C:311-391
  if (vm.numMs.empty()) return "";
  std::vector<int> idxs;
  pick_min_missing(vm, true, idxs);
  if (idxs.empty()) return "";
  if (idxs.size() == 1) return vm.homRecord(idxs[0]);
  if (idxs[0] == 0 && idxs.size() == 2) return vm.homRecord(idxs[1]);
  int maxVars = 0, maxIdx = idxs[0];
  for (size_t ii = 1; ii < idxs.size(); ii++) {
    int count = 0;
    for (int a : vm.gtPaths[idxs[ii]]) if (a > 0) count++;
    if (count > maxVars) { maxVars = count; maxIdx = idxs[ii]; }
  }
  fprintf(stderr, "[ WARNING ] :: Multiple (%lu) alternate pathes detected in a path beginning with variant : %s",
          (unsigned long)idxs.size(), vm.posGt->gts[0]->record->save().c_str());
  fprintf(stderr, "[ WARNING ] :: Max. %d ALT variants selected\n", maxVars);
  return vm.homRecord(maxIdx);
}

// the tie-break container is declared multimap<double,int,greater<int>> (varMer.H:72): keys are
// compared after conversion to int, descending (Q6)
struct AsIntDescending { bool operator()(double a, double b) const { return (int)a > (int)b; } };

static std::string best_variant(const VarMer &vm, uint32_t k) {                         // varMer.C:396-467
  uint32_t numMissing = UINT32_MAX;
  std::vector<int> idxs;
  for (size_t ii = 0; ii < vm.numMs.size(); ii++) {
    if (vm.numMs[ii] == vm.seqs[ii].size() - k + 1) continue;
    if (vm.numMs[ii] < numMissing) { numMissing = vm.numMs[ii]; idxs.clear(); idxs.push_back((int)ii); }
    else if (vm.numMs[ii] == numMissing) idxs.push_back((int)ii);
  }
  if (numMissing == UINT32_MAX) return "";
  if (idxs.size() == 1) return vm.homRecord(idxs[0]);
  if (idxs.size() > 1) {
    std::multimap<double, int, AsIntDescending> byK;
    for (int idx : idxs) byK.insert(std::make_pair(vm.totdK(idx), idx));
    auto it = byK.begin();
    const double k1 = it->first; const int i1 = it->second;
    ++it;
    const double k2 = it->first; const int i2 = it->second;
    if (k1 == k2) {
      if (vm.seqs[i1].size() >= vm.seqs[i2].size()) return vm.hetRecord(i1, i2);
      return vm.hetRecord(i2, i1);
    }
    return vm.homRecord(i1);
  }
  return "";
}

static std::string best_filter(const VarMer &vm, uint32_t k) {                          // varMer.C:150-194
  uint32_t numMissing = UINT32_MAX;
  std::vector<int> idxs;
  for (size_t ii = 0; ii < vm.numMs.size(); ii++) {
    if (vm.numMs[ii] == vm.seqs[ii].size() - k + 1) continue;
    if (vm.numMs[ii] == 0) { idxs.push_back((int)ii); numMissing = 0; }
    if (vm.numMs[ii] < numMissing) { numMissing = vm.numMs[ii]; idxs.clear(); idxs.push_back((int)ii); }
    else if (vm.numMs[ii] == numMissing) idxs.push_back((int)ii);
  }
  std::string out;
  if (idxs.empty()) return out;
  std::list<int> gtIdxs;
  for (int idx : idxs)
    for (size_t i = 0; i < vm.gtPaths[idx].size(); i++)
      if (vm.gtPaths[idx][i] > 0) gtIdxs.push_back((int)i);
  gtIdxs.sort();
  gtIdxs.unique();
  for (int i : gtIdxs) out += vm.posGt->gts[i]->record->save();
  return out;
}

// ---------------------------------------------------------------------------------------------
// merfin-variants.C:131-310 for one sequence record, with the probes of all its clusters batched
// ---------------------------------------------------------------------------------------------
int process_sequence(const std::string &ident, const char *seq, uint64_t seqLen, std::vector<PosGT *> &clusters,
                     const Options &opt, const PathScorer &scorer, std::string &result, std::string *debug,
                     uint64_t &varMerId, size_t cluster_begin, size_t cluster_end, size_t batch_clusters) {
  if (!getenv("MFN_QUIET")) fprintf(stderr, "Processing sequence %s for variants\n", ident.c_str());
  cluster_end = std::min(cluster_end, clusters.size());
  if (cluster_begin >= cluster_end) return 0;
  if (batch_clusters == 0) batch_clusters = cluster_end - cluster_begin;
  // `prob` is carried from window to window (and cluster to cluster) like the reference's local: it
  // is left uninitialised there until the first valid window (Q5).  A range that does not start at
  // the record's first cluster scores a few clusters before it first (results dropped) to enter
  // with the value a whole-record run would carry in.
  double prob = 1.0;
  const size_t PREROLL = 4;
  const size_t first = cluster_begin > PREROLL ? cluster_begin - PREROLL : 0;

  struct Work { VarMer vm; uint32_t rStart, rEnd; size_t first_path; bool keep; };
  std::vector<Work> work;
  std::string paths;
  std::vector<uint64_t> off;
  ScoreOut sc;
  Options sopt = opt;
  sopt.debug = debug != nullptr;                     // the per-window values are kept exactly when the debug records are written
  const KTable kt(&opt.km);
  for (size_t b0 = first; b0 < cluster_end;) {
    // the pre-roll clusters ride along with the first batch
    const size_t b1 = std::min(cluster_end, (b0 < cluster_begin ? cluster_begin : b0) + batch_clusters);
    work.clear();
    work.reserve(b1 - b0);
    paths.clear();
    off.assign(1, 0);
    for (size_t ci = b0; ci < b1; ci++) {
      PosGT *pg = clusters[ci];
      uint32_t rStart = pg->rStart, rEnd = pg->rEnd;
      const uint32_t K_PADD = opt.k - 1;
      if (rStart > K_PADD) rStart -= K_PADD; else rStart = 0;
      if (rEnd < seqLen - K_PADD) rEnd += K_PADD; else rEnd = (uint32_t)seqLen;        // (unsigned, as the reference)
      std::vector<uint32_t> idxList, lenList;
      std::vector<const std::vector<std::string> *> haps;
      for (GtAllele *g : pg->gts) { idxList.push_back(g->pos - rStart); lenList.push_back(g->refLen); haps.push_back(&g->alleles); }
      const bool keep = ci >= cluster_begin;
      if (rEnd < rStart || seqLen < rEnd) {                                            // dnaSeq::copy fails (sequence.C:521)
        if (keep) fprintf(stderr, "PANIC : Invalid region specified: %s : %u - %u\n", ident.c_str(), rStart, rEnd);
        continue;
      }
      if (idxList.size() > opt.comb) {
        if (keep) fprintf(stderr, "PANIC : Combination %s:%u-%u has too many variants ( found %lu > %u ) to evaluate. Consider filtering the vcf upfront. Skipping...\n",
                          ident.c_str(), rStart, rEnd, (unsigned long)pg->gts.size(), opt.comb);
        continue;
      }
      work.emplace_back();
      Work &w = work.back();
      w.vm.posGt = pg; w.rStart = rStart; w.rEnd = rEnd; w.first_path = off.size() - 1; w.keep = keep;
      {
        // room for the usual two paths (REF and one ALT) up front: the vectors would otherwise grow 1 -> 2 and move
        VarMer &v = w.vm;
        v.seqs.reserve(2); v.gtPaths.reserve(2); v.idxPaths.reserve(2); v.lenPaths.reserve(2); v.numMs.reserve(2); v.totdKs.reserve(2);
      }
      std::vector<int> path;
      traverse(0, idxList, lenList, haps, std::string(seq + rStart, rEnd - rStart), path, w.vm);
      for (const std::string &s : w.vm.seqs) { paths += s; off.push_back(paths.size()); }
    }

    if (!paths.empty() || off.size() > 1) {
      int st = scorer(paths, off, sc);
      if (st != 0) return st;
    }

    for (Work &w : work) {
      VarMer &vm = w.vm;
      score(vm, sopt, kt, sc.v_read, sc.v_asm, sc.valid, off, w.first_path, prob);
      if (!w.keep) continue;
      if (debug) {
        char buf[512];
        for (size_t idx = 0; idx < vm.seqs.size(); idx++) {
          snprintf(buf, sizeof buf, "%lu\t%s:%u-%u\t", (unsigned long)varMerId++, ident.c_str(), w.rStart, w.rEnd);
          *debug += buf; *debug += vm.seqs[idx];
          snprintf(buf, sizeof buf, "\t%u\t%.5f\t%.5f\t%.5f\t%.5f\t%.5f\t", vm.numMs[idx], vm.minAbsK((int)idx), vm.maxAbsK((int)idx),
                   vm.medAbsK((int)idx), vm.avgAbsK((int)idx), vm.totdK((int)idx));
          *debug += buf;
          for (size_t i = 0; i < vm.gtPaths[idx].size(); i++) {
            int alt = vm.gtPaths[idx][i];
            if (alt > 0) {
              const GtAllele *g = vm.posGt->gts[i];
              *debug += ident + " " + std::to_string(g->pos + 1) + " . " + g->alleles[0] + " " + g->alleles[alt] + " . PASS . GT 1/1  ";
            }
          }
          *debug += "\n";
        }
      }
      switch (opt.mode) {
        case POLISH: result += best_variant(vm, opt.k); break;
        case BETTER: result += better_or_strict(vm); break;
        case STRICT: result += better_or_strict(vm); break;
        case LOOSE:  result += loose(vm); break;
        default:     result += best_filter(vm, opt.k); break;
      }
    }
    b0 = b1;
  }
  return 0;
}

}  // namespace mfnv
