// See variants.h.  Host-only control logic; every behaviour that is visible in merfin's output is
// cited against the reference line it reproduces.
#include "variants.h"

#include <atomic>
#include <thread>

#include "fasta.h"

#include <algorithm>
#include <cfloat>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace mfnv {

// strings.C:188-249: runs of separators collapse, empty words do not exist
typedef std::pair<const char *, size_t> Span;
static void split_spans(const char *s, size_t n, char sep, std::vector<Span> &out) {
  out.clear();
  const char *p = s, *const e = s + n;
  while (p < e) {
    while (p < e && *p == sep) p++;
    if (p == e) break;
    const char *q = (const char *)memchr(p, sep, (size_t)(e - p));
    if (!q) q = e;
    out.emplace_back(p, (size_t)(q - p));
    p = q;
  }
}
template <class I> static inline void put_int(std::string &out, I v) {
  char num[24];
  out.append(num, std::to_chars(num, num + sizeof num, v).ptr);
}

std::string VcfRecord::save() const {
  char num[64];
  const int nn = std::min<int>(snprintf(num, sizeof num, "%.1f", qual), (int)sizeof num - 1);
  std::string out;
  size_t need = 32 + (size_t)nn;
  for (int i = 0; i < NCOL; i++) need += col_len[i];
  out.reserve(need);
  out += col(CHROM); out += '\t'; put_int(out, (int)pos); out += '\t'; out += col(ID); out += '\t'; out += col(REF); out += '\t';
  out += col(ALT); out += '\t'; out.append(num, (size_t)nn);
  for (int i = FILTER; i < NCOL; i++) { out += '\t'; out += col(i); }
  out += '\n';
  return out;
}

// vcf.C:23-87: the alleles the sample's genotype names.  alts: the ALT column split at ','; gt: the first
// ':'-separated field of the sample column.
static void make_allele(GtAllele *g, const VcfRecord *r, const std::vector<Span> &alts, std::string_view gt) {
  const std::string_view ref = r->col(VcfRecord::REF);
  g->record = r;
  g->pos = r->pos - 1;
  g->refLen = (uint32_t)ref.size();
  g->qual = r->qual;
  if (gt.compare(0, 3, "./.") == 0 || gt.compare(0, 3, "0/0") == 0) return;            // no alleles at all (Q10)
  g->alleles.push_back(ref);
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
    if ((size_t)idx > alts.size()) continue;
    bool seen = std::find(used, used + nused, (int)idx) != used + nused ||
                std::find(used_more.begin(), used_more.end(), (int)idx) != used_more.end();
    if (seen) continue;                                                             // same ALT column entry again
    if (nused < 8) used[nused++] = (int)idx; else used_more.push_back((int)idx);
    const std::string_view hap(alts[idx - 1].first, alts[idx - 1].second);
    if (hap == ref) continue;
    g->alleles.push_back(hap);
  }
}

// vcf.C:93-151.  Lines are independent, so they are parsed on host threads (contiguous ranges of
// lines per thread) and appended in file order afterwards: the order inside byChr is what the
// reference's (unstable) sort starts from, so it must be the file order.  Nothing is allocated per
// line: a thread's records, alleles and clusters-to-be are three arrays sized for its share of the
// lines (they never move, so the pointers between them hold), and all text is spans of text_.
FileText::~FileText() { if (map_) munmap(map_, map_len_); }

bool FileText::load(const std::string &path, std::string &err) {
  auto ends = [&](const char *suf) { const size_t m = strlen(suf); return path.size() >= m && path.compare(path.size() - m, m, suf) == 0; };
  struct stat st;
  if (!ends(".gz") && !ends(".bz2") && !ends(".xz") && stat(path.c_str(), &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) { err = "cannot open '" + path + "'"; return false; }
    void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
    close(fd);
    if (m != MAP_FAILED) { map_ = m; map_len_ = (size_t)st.st_size; p = (const char *)m; n = map_len_; return true; }
  }
  if (!slurp_file(path, owned_, err)) return false;      // compressed files, pipes, empty files
  p = owned_.data(); n = owned_.size();
  return true;
}

bool VcfFile::load(const std::string &path, std::string &err) {
  if (!text_.load(path, err)) return false;
  const char *const data = text_.p;
  const size_t N = text_.n;
  unsigned T = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  if (N < (1u << 20)) T = 1;
  if (const char *env = getenv("MFN_VCF_THREADS")) T = (unsigned)std::max(1, atoi(env));     // (tests force the threaded path on small files)
  T = (unsigned)std::max<size_t>(1, std::min<size_t>(T, N / 64 + 1));
  struct Run { std::string_view chr; size_t first, count; };      // consecutive records of one sequence (indices into the thread's arrays)
  struct Part { std::vector<std::string> headers; uint64_t excluded = 0; std::vector<Run> runs; };
  std::vector<Part> parts(T);
  records_.assign(T, {}); alleles_.assign(T, {}); posgts_.assign(T, {});
  // thread t takes the lines that START in bytes [N t / T, N (t + 1) / T)
  auto line_start_at_or_after = [&](size_t p) -> size_t {
    if (p == 0 || p >= N) return std::min(p, N);
    if (data[p - 1] == '\n') return p;
    const char *nl = (const char *)memchr(data + p, '\n', N - p);
    return nl ? (size_t)(nl - data) + 1 : N;
  };
  auto parse = [&](unsigned t) {
    const size_t b0 = line_start_at_or_after(N * t / T), b1 = line_start_at_or_after(N * (t + 1) / T);
    size_t nl_count = 0;
    for (size_t p = b0; p < b1;) {
      const char *nl = (const char *)memchr(data + p, '\n', b1 - p);
      nl_count++;
      p = nl ? (size_t)(nl - data) + 1 : b1;
    }
    std::vector<VcfRecord> &recs = records_[t];
    std::vector<GtAllele> &als = alleles_[t];
    std::vector<PosGT> &pgs = posgts_[t];
    recs.reserve(nl_count); als.reserve(nl_count); pgs.reserve(nl_count);     // (never grown: the elements must not move)
    std::vector<Span> w, alts, fields;               // reused for every line of this thread
    w.reserve(16); alts.reserve(8); fields.reserve(8);
    for (size_t next = b0; next < b1;) {
      size_t p = next, e;
      {
        const char *nl = (const char *)memchr(data + p, '\n', b1 - p);
        e = next = nl ? (size_t)(nl - data) + 1 : b1;
      }
      while (e > p && (data[e - 1] == '\r' || data[e - 1] == '\n')) e--;
      if (e == p) continue;
      const char *line = data + p;
      const size_t len = e - p;
      if (line[0] == '#') { parts[t].headers.emplace_back(line, len); continue; }
      split_spans(line, len, '\t', w);
      if (w.size() < 10) { parts[t].excluded++; continue; }                            // vcfRecord.H:53-56
      recs.emplace_back();
      VcfRecord *r = &recs.back();
      r->line = line;
      for (int i = 0; i < VcfRecord::NCOL; i++) { r->col_off[i] = (uint32_t)(w[i].first - line); r->col_len[i] = (uint32_t)w[i].second; }
      char num[64];
      auto cstr = [&](int i) { const size_t n = std::min<size_t>(w[i].second, sizeof num - 1); memcpy(num, w[i].first, n); num[n] = 0; return num; };
      // POS and QUAL as strtoul / strtod read them; plain decimal numbers of a few digits take the exact short cut
      // (an integer below 2^53 divided by a power of ten up to 10^22 is correctly rounded, i.e. what strtod returns)
      {
        const char *q = w[VcfRecord::POS].first;
        const size_t n = w[VcfRecord::POS].second;
        uint64_t v = 0;
        size_t i = 0;
        for (; i < n && i < 18 && q[i] >= '0' && q[i] <= '9'; i++) v = v * 10 + (uint64_t)(q[i] - '0');
        r->pos = (i == n && n > 0) ? (uint32_t)v : (uint32_t)strtoul(cstr(VcfRecord::POS), nullptr, 10);
      }
      {
        static const double P10[16] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15};
        const char *q = w[VcfRecord::QUAL].first;
        const size_t n = w[VcfRecord::QUAL].second;
        uint64_t v = 0;
        size_t i = 0, digits = 0, frac = 0;
        bool dot = false, plain = n > 0 && n <= 16;
        for (; plain && i < n; i++) {
          if (q[i] >= '0' && q[i] <= '9') { v = v * 10 + (uint64_t)(q[i] - '0'); digits++; if (dot) frac++; }
          else if (q[i] == '.' && !dot) dot = true;
          else plain = false;
        }
        r->qual = (plain && digits > 0 && digits <= 15) ? (double)v / P10[frac] : strtod(cstr(VcfRecord::QUAL), nullptr);
      }
      split_spans(w[VcfRecord::ALT].first, w[VcfRecord::ALT].second, ',', alts);
      split_spans(w[VcfRecord::SAMPLE].first, w[VcfRecord::SAMPLE].second, ':', fields);
      als.emplace_back();
      GtAllele *g = &als.back();
      make_allele(g, r, alts, fields.empty() ? std::string_view() : std::string_view(fields[0].first, fields[0].second));
      pgs.emplace_back();
      PosGT *pg = &pgs.back();
      pg->rStart = g->pos; pg->rEnd = g->pos + g->refLen; pg->gts.push_back(g);
      const std::string_view chr = r->col(VcfRecord::CHROM);
      std::vector<Run> &runs = parts[t].runs;
      if (runs.empty() || runs.back().chr != chr) runs.push_back(Run{chr, recs.size() - 1, 1});
      else runs.back().count++;
    }
  };
  {
    std::vector<std::thread> th;
    for (unsigned t = 1; t < T; t++) th.emplace_back(parse, t);
    parse(0);
    for (auto &x : th) x.join();
  }
  // Every sequence's list of clusters-to-be, in file order: the runs are placed serially (where each run of a thread
  // goes in its sequence's list), the lists sized once, and then every thread fills in its own runs.
  typedef std::pair<const std::string, std::vector<PosGT *>> ChrNode;
  struct Job { size_t first, count, dst; ChrNode *node; };
  std::vector<std::vector<Job>> jobs(T);
  std::map<ChrNode *, size_t> fill;
  for (unsigned t = 0; t < T; t++) {
    excluded += parts[t].excluded;
    for (std::string &h : parts[t].headers) {
      if (h.compare(0, 12, "##contig=<ID") == 0) numChr++;
      headers.push_back(std::move(h));
    }
    num_records += records_[t].size();
    for (const Run &r : parts[t].runs) {
      ChrNode *node = &*byChr.try_emplace(std::string(r.chr)).first;
      size_t &at = fill[node];
      jobs[t].push_back(Job{r.first, r.count, at, node});
      at += r.count;
    }
  }
  for (auto &kv : fill) kv.first->second.resize(kv.second);
  auto place = [&](unsigned t) {
    for (const Job &j : jobs[t])
      for (size_t i = 0; i < j.count; i++) {
        PosGT *pg = &posgts_[t][j.first + i];
        pg->chr = &j.node->first;
        j.node->second[j.dst + i] = pg;
      }
  };
  {
    std::vector<std::thread> th;
    for (unsigned t = 1; t < T; t++) th.emplace_back(place, t);
    place(0);
    for (auto &x : th) x.join();
  }
  return true;
}

// vcf.C:156-246.  std::sort with the reference's comparator on the reference's initial (file) order.
void VcfFile::mergeChrPosGT(uint32_t ksize, uint32_t comb, bool nosplit) {
  const uint32_t K_OFFSET = 2 * ksize;
  const bool quiet = getenv("MFN_QUIET") != nullptr;
  // the sequences are independent of each other: one at a time per host thread, messages printed in map order afterwards
  std::vector<std::pair<const std::string, std::vector<PosGT *>> *> chrs;
  for (auto &kv : byChr) chrs.push_back(&kv);
  std::vector<std::string> msgs(chrs.size());
  auto one = [&](size_t ci) {
    const std::string &name = chrs[ci]->first;
    std::vector<PosGT *> &in = chrs[ci]->second;
    std::vector<PosGT *> out;
    uint32_t removed = 0, nsplit = 0, merged = 0;
    std::sort(in.begin(), in.end(), [](PosGT *const &A, PosGT *const &B) { return A->rStart < B->rStart; });
    out.push_back(in[0]);
    for (size_t ii = 1; ii < in.size(); ii++) {
      if (in[ii]->gts.empty()) { removed++; continue; }
      const bool overlapping = in[ii]->rStart < out.back()->rEnd + K_OFFSET;
      const bool toomany = out.back()->gts.size() >= comb;
      if (!overlapping) { out.push_back(in[ii]); continue; }
      if (toomany && !nosplit) { out.push_back(in[ii]); nsplit++; continue; }
      GtAllele *g = in[ii]->gts[0];
      out.back()->gts.push_back(g);
      out.back()->rStart = std::min(out.back()->rStart, g->pos);
      out.back()->rEnd = std::max(out.back()->rEnd, g->pos + g->refLen);
      merged++;
    }
    if (!quiet) {
      char buf[256];
      std::string &m = msgs[ci];
      snprintf(buf, sizeof buf, " : Reduced %lu variants down to %lu combinations for evaluation:\n", (unsigned long)in.size(), (unsigned long)out.size());
      m += name; m += buf;
      if (removed > 0) { snprintf(buf, sizeof buf, " :   Removed %u empty alleles.\n", removed); m += name; m += buf; }
      if (nsplit > 0) { snprintf(buf, sizeof buf, " :   Split   %u complicated combinations.\n", nsplit); m += name; m += buf; }
      if (merged > 0) { snprintf(buf, sizeof buf, " :   Merged  %u variants into combinations.\n", merged); m += name; m += buf; }
    }
    in.swap(out);
  };
  unsigned T = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  if (num_records < 200000) T = 1;
  T = (unsigned)std::max<size_t>(1, std::min<size_t>(T, chrs.size()));
  std::atomic<size_t> next{0};
  auto worker = [&]() { for (size_t ci; (ci = next.fetch_add(1)) < chrs.size();) one(ci); };
  std::vector<std::thread> th;
  for (unsigned t = 1; t < T; t++) th.emplace_back(worker);
  worker();
  for (auto &x : th) x.join();
  for (const std::string &m : msgs) fputs(m.c_str(), stderr);
}

// ---------------------------------------------------------------------------------------------
// One batch of clusters.  What the reference keeps in one varMer object per cluster (the candidate
// paths, their genotype / index / length lists and their scores) lies side by side in flat arrays for
// the whole batch: `paths` + `off` are exactly what the scorer takes, nothing is allocated per
// cluster or per path, and the batch's storage is reused by the next one.
// ---------------------------------------------------------------------------------------------
struct Cluster {
  PosGT *posGt = nullptr;
  uint32_t rStart = 0, rEnd = 0;
  uint32_t first_path = 0, npaths = 0;   // paths [first_path, first_path + npaths) of the batch
  uint32_t nvar = 0;                     // variants of the cluster = entries per path in gt / idx / len
  size_t first_var = 0;                  // the entries of the cluster's path i start at first_var + i * nvar
  bool keep = false;
};

struct Batch {
  std::string paths;                     // all candidate sequences of the batch, concatenated
  std::vector<uint64_t> off;             // path p is paths[off[p], off[p + 1])
  std::vector<int> gt;                   // varMer::gtPaths
  std::vector<uint32_t> idx, len;        // varMer::idxPaths / lenPaths
  std::vector<uint32_t> numM;            // per path
  std::vector<double> totdK;             // per path: sum of the path's dkstrs, added up in window order
  std::vector<std::vector<double>> kstrs, dkstrs;   // per path and window; kept only for the -debug statistics
  std::vector<Cluster> clusters;
  void clear() {
    paths.clear(); off.assign(1, 0); gt.clear(); idx.clear(); len.clear(); numM.clear(); totdK.clear();
    kstrs.clear(); dkstrs.clear(); clusters.clear();
  }
};

// One cluster of a batch as the selectors and the output records read it (the reference's varMer)
struct VarMer {
  const Batch &b;
  const Cluster &c;
  size_t npaths() const { return c.npaths; }
  size_t seqLen(size_t i) const { return (size_t)(b.off[c.first_path + i + 1] - b.off[c.first_path + i]); }
  const char *seq(size_t i) const { return b.paths.data() + b.off[c.first_path + i]; }
  const int *gtPath(size_t i) const { return b.gt.data() + c.first_var + i * c.nvar; }
  uint32_t numM(size_t i) const { return b.numM[c.first_path + i]; }
  double totdK(size_t i) const { return b.totdK[c.first_path + i]; }                                  // varMer.C:646-659 (same additions, same order)
  const std::vector<double> &kstr(size_t i) const { return b.kstrs[c.first_path + i]; }
  double minAbsK(size_t i) const {                                                                    // :553-569
    double m = DBL_MAX;
    for (double a : kstr(i)) { if (a < 0) continue; if (a < m) m = a; }
    return m == DBL_MAX ? -1 : m;
  }
  double maxAbsK(size_t i) const { double m = -2; for (double a : kstr(i)) if (a > m) m = a; return m; }  // :572-584
  double avgAbsK(size_t i) const {                                                                    // :586-604
    double s = 0;
    for (double a : kstr(i)) if (a >= 0) s += a;
    if (kstr(i).size() == numM(i)) return -1;
    return s / (kstr(i).size() - numM(i));
  }
  double medAbsK(size_t i) const {                                                                    // :606-622
    std::vector<double> k = kstr(i);
    std::sort(k.begin(), k.end());
    size_t j = 0;
    for (; j < k.size(); j++) if (k[j] >= 0) break;
    if (j == k.size()) return -1;
    return k[j + (k.size() - j) / 2];
  }
  void homRecord(size_t idx, std::string &out) const;
  void hetRecord(size_t idx1, size_t idx2, std::string &out) const;
};

// chr <tab> pos <tab> . <tab> REF <tab> ALT[,ALT2] <tab> (int)QUAL <tab> PASS <tab> . <tab> GT <tab> gt   (varMer.C:487,534)
static inline void put_record(std::string &out, const PosGT *pg, const GtAllele *g, std::string_view alt, const std::string_view *alt2,
                              const char *gt) {
  out += *pg->chr; out += '\t'; put_int(out, (uint32_t)(g->pos + 1)); out += "\t.\t"; out += g->alleles[0]; out += '\t'; out += alt;
  if (alt2) { out += ','; out += *alt2; }
  out += '\t'; put_int(out, (int)g->qual); out += "\tPASS\t.\tGT\t"; out += gt; out += '\n';
}

void VarMer::homRecord(size_t idx, std::string &out) const {                            // varMer.C:531-550
  const int *gp = gtPath(idx);
  for (size_t i = 0; i < c.nvar; i++) {
    const int alt = gp[i];
    if (alt <= 0) continue;
    const GtAllele *g = c.posGt->gts[i];
    put_record(out, c.posGt, g, g->alleles[alt], nullptr, "1/1");
  }
}

void VarMer::hetRecord(size_t idx1, size_t idx2, std::string &out) const {              // varMer.C:472-529
  const int *g1 = gtPath(idx1), *g2 = gtPath(idx2);
  for (size_t i = 0; i < c.nvar; i++) {
    const int a1 = g1[i], a2 = g2[i];
    if (a1 + a2 <= 0) continue;
    const GtAllele *g = c.posGt->gts[i];
    if (a1 == a2)               put_record(out, c.posGt, g, g->alleles[a1], nullptr, "1/1");
    else if (a1 == 0 && a2 > 0) put_record(out, c.posGt, g, g->alleles[a2], nullptr, "0/1");
    else if (a1 > 0 && a2 > 0)  { const std::string_view second = g->alleles[a2]; put_record(out, c.posGt, g, g->alleles[a1], &second, "1/2"); }
    else if (a1 > 0 && a2 == 0) put_record(out, c.posGt, g, g->alleles[a1], nullptr, "1/0");
    else { out += *c.posGt->chr; out += '\t'; put_int(out, (uint32_t)(g->pos + 1)); out += "\t.\t"; out += g->alleles[0]; out += '\t'; }   // (negative index: the reference leaves the line unfinished)
  }
}

// ---------------------------------------------------------------------------------------------
// merfin-variants.C:22-126: every combination of the alleles of a cluster's variants.  idxList is
// shared down the recursion (shifted by the length change of every ALT taken so far and restored
// afterwards); lenList is private to each level (a row of lenStack); the candidate of a level that
// goes deeper is built in that level's scratch string, the candidate of a leaf straight at the end
// of the batch's path buffer.
// ---------------------------------------------------------------------------------------------
struct Enumerator {
  Batch &b;
  Cluster *cur = nullptr;
  uint32_t n = 0;                                              // variants of the current cluster
  std::vector<uint32_t> idxList, lenStack;
  std::vector<const SmallVec<std::string_view, 3> *> haps;
  std::vector<int> path;
  std::vector<std::string> scratch;                            // one per recursion depth, capacity kept

  explicit Enumerator(Batch &batch) : b(batch) {}

  // varMer::addSeqPath (varMer.C:37-45): a sequence already among the cluster's paths is dropped
  void add(const char *s0, size_t n0, const char *s1, size_t n1, const char *s2, size_t n2, const uint32_t *lenList) {
    const size_t start = b.paths.size(), L = n0 + n1 + n2;
    if (n0) b.paths.append(s0, n0);
    if (n1) b.paths.append(s1, n1);
    if (n2) b.paths.append(s2, n2);
    const char *base = b.paths.data();
    for (uint32_t p = cur->first_path; p < cur->first_path + cur->npaths; p++)
      if (b.off[p + 1] - b.off[p] == L && memcmp(base + b.off[p], base + start, L) == 0) { b.paths.resize(start); return; }
    b.off.push_back(b.paths.size());
    b.gt.insert(b.gt.end(), path.begin(), path.end());
    b.idx.insert(b.idx.end(), idxList.begin(), idxList.end());
    b.len.insert(b.len.end(), lenList, lenList + n);
    cur->npaths++;
  }

  void go(uint32_t idx, uint32_t depth, const char *cand, size_t candLen) {
    uint32_t *lenList = lenStack.data() + (size_t)depth * n;
    const SmallVec<std::string_view, 3> &alts = *haps[idx];
    const uint32_t refLen = lenList[idx];
    const uint32_t last = n - 1;
    for (size_t j = 0; j < alts.size(); j++) {
      path.push_back((int)j);
      int skipped = 0, delta = 0;
      if (j == 0) {                                              // the REF allele keeps the candidate
        if (idx == last) add(cand, candLen, nullptr, 0, nullptr, 0, lenList);
        else { memcpy(lenList + n, lenList, n * sizeof(uint32_t)); go(idx + 1, depth + 1, cand, candLen); }
      } else {
        lenList[idx] = refLen;
        const std::string_view hap = alts[j];
        const size_t at = idxList[idx];
        if (at > candLen) throw std::out_of_range("basic_string::replace");      // as the reference's std::string::replace
        const size_t cut = std::min<size_t>(lenList[idx], candLen - at);
        delta = (int)hap.size() - (int)lenList[idx];
        const uint32_t affected = idxList[idx] + lenList[idx];
        lenList[idx] = (uint32_t)hap.size();
        for (uint32_t i = idx + 1; i < n; i++) {                  // later variants inside the replaced span keep REF
          if (idxList[i] >= affected) break;
          idx++; path.push_back(0); skipped++;
        }
        if (idx == last) {
          add(cand, at, hap.data(), hap.size(), cand + at + cut, candLen - at - cut, lenList);
        } else {
          for (uint32_t i = idx + 1; i < n; i++) idxList[i] += delta;
          std::string &changed = scratch[depth];
          changed.assign(cand, at); changed.append(hap.data(), hap.size()); changed.append(cand + at + cut, candLen - at - cut);
          memcpy(lenList + n, lenList, n * sizeof(uint32_t));
          go(idx + 1, depth + 1, changed.data(), changed.size());
          for (uint32_t i = idx + 1; i < n; i++) idxList[i] -= delta;
        }
      }
      for (int s = 0; s < skipped; s++) { path.pop_back(); idx--; }
      path.pop_back();
    }
  }

  // the paths of cluster c (its variants' alleles over the region [c.rStart, c.rEnd) of the record)
  void run(Cluster &c, const char *seq) {
    cur = &c;
    const SmallVec<GtAllele *, 2> &gts = c.posGt->gts;
    n = (uint32_t)gts.size();
    c.nvar = n; c.first_path = (uint32_t)(b.off.size() - 1); c.npaths = 0; c.first_var = b.gt.size();
    if (n == 0) return;
    idxList.resize(n); haps.resize(n);
    lenStack.resize((size_t)(n + 1) * n);
    if (scratch.size() < n + 1) scratch.resize(n + 1);
    for (uint32_t i = 0; i < n; i++) { idxList[i] = gts[i]->pos - c.rStart; lenStack[i] = gts[i]->refLen; haps[i] = &gts[i]->alleles; }
    path.clear();
    go(0, 0, seq + c.rStart, c.rEnd - c.rStart);
  }
};

// merfinGlobal::getK (merfin-globals.C:66-98) per read multiplicity, tabulated once per call of
// process_sequence: the replay below asks for it once per window of every path
struct KTable {
  static constexpr uint32_t N = 4096;
  struct Entry { double readK, prob; };
  std::vector<Entry> e;
  const mfn_kmodel *km;
  bool tame = true;                                  // every tabulated value is finite and far from overflow
  explicit KTable(const mfn_kmodel *m) : e(N), km(m) {
    for (uint32_t v = 0; v < N; v++) {
      double a;
      mfn_getK(m, v, 0, &e[v].readK, &a, &e[v].prob);
      if (!(fabs(e[v].readK) < 1e100 && fabs(e[v].prob) < 1e100)) tame = false;
    }
  }
  inline void get(uint32_t vr, double &readK, double &prob) const {
    if (vr < N) { readK = e[vr].readK; prob = e[vr].prob; }
    else { double a; mfn_getK(km, vr, 0, &readK, &a, &prob); }
  }
};

// varMer.C:48-145 from the integer pairs of every window of every path of one cluster.
//
// Per window the reference computes oD = |readK - asmK| * prob, adds one to asmK if the window touches
// an ALT allele of the path, computes nD the same way and adds oD - nD to the path's delta sum.  A path
// without ALT alleles (the REF path) adds oD - oD, which is +0.0 whenever oD is finite, so for those
// paths -- and for -filter, which only counts the missing windows -- only the count is done; a window
// whose K values are not comfortably finite (a zero -peak, odd -prob tables) sends the path through
// the general loop instead.
static void score(Batch &b, const Cluster &c, const Options &opt, const KTable &kt, const uint32_t *vr, const uint32_t *va,
                  const uint8_t *ok, double &prob_io, std::vector<uint32_t> &alt_lo, std::vector<uint32_t> &alt_hi) {
  const uint32_t k = opt.k;
  double prob = prob_io;
  const bool keep = opt.debug;                         // the per-window values are only read by the -debug statistics
  const size_t nvar = c.nvar;
  const KTable::Entry *tab = kt.e.data();
  for (size_t ii = 0; ii < c.npaths; ii++) {
    const size_t p = c.first_path + ii;
    const uint64_t base = b.off[p];
    const size_t L = (size_t)(b.off[p + 1] - base);
    const size_t nwin = L >= k ? L - k + 1 : 0;
    const uint32_t *ipath = b.idx.data() + c.first_var + ii * nvar, *lpath = b.len.data() + c.first_var + ii * nvar;
    const int *gpath = b.gt.data() + c.first_var + ii * nvar;
    const uint32_t *pvr = vr + base, *pva = va + base;
    const uint8_t *pok = ok + base;
    // the windows an ALT allele of this path reaches: lo <= window < hi in unsigned arithmetic, as the
    // reference writes it: ip + 1 - k wraps when ip < k - 1 and the test is then never true (Q5)
    alt_lo.clear(); alt_hi.clear();
    for (size_t jj = 0; jj < nvar; jj++)
      if (gpath[jj] > 0) { alt_lo.push_back((uint32_t)(ipath[jj] + 1 - k)); alt_hi.push_back((uint32_t)(ipath[jj] + lpath[jj] + k)); }
    const size_t nalt = alt_lo.size();

    if (opt.mode == FILTER || (nalt == 0 && !keep)) {
      const double prob_in = prob;
      uint32_t numM = 0;
      bool exact = true;
      for (size_t w = 0; w < nwin; w++) {
        double readK = 0;
        if (pok[w]) {
          const uint32_t v = pvr[w];
          if (v < KTable::N) { readK = tab[v].readK; prob = tab[v].prob; }
          else { kt.get(v, readK, prob); if (!(fabs(readK) < 1e100 && fabs(prob) < 1e100)) exact = false; }
        }
        if (readK == 0) numM++;
      }
      if (opt.mode == FILTER || (exact && kt.tame)) { b.numM[p] = numM; b.totdK[p] = 0.0; continue; }
      prob = prob_in;
    }

    uint32_t numM = 0;
    std::vector<double> ks, dks;
    double totd = 0;
    if (keep) { ks.reserve(nwin); dks.reserve(nwin); }
    const uint32_t lo0 = nalt ? alt_lo[0] : 1, hi0 = nalt ? alt_hi[0] : 0;
    for (size_t w = 0; w < nwin; w++) {
      const uint32_t idx = (uint32_t)w;
      double readK = 0, asmK = 0;
      if (pok[w]) { kt.get(pvr[w], readK, prob); asmK = (double)pva[w]; }
      if (readK == 0) numM++;
      const double oD = fabs(readK - asmK) * prob;
      bool hit = lo0 <= idx && idx < hi0;
      for (size_t jj = 1; jj < nalt && !hit; jj++) hit = alt_lo[jj] <= idx && idx < alt_hi[jj];
      if (hit) asmK++;
      const double nD = fabs(readK - asmK) * prob;
      totd += oD - nD;
      if (keep) {
        double kMetric;
        if (readK == 0) kMetric = -1;
        else if (readK > asmK) kMetric = readK / asmK - 1;
        else kMetric = asmK / readK - 1;
        ks.push_back(kMetric); dks.push_back(oD - nD);
      }
    }
    b.numM[p] = numM; b.totdK[p] = totd;
    if (keep) { b.kstrs[p] = std::move(ks); b.dkstrs[p] = std::move(dks); }
  }
  prob_io = prob;
}

// ---- selectors (varMer.C:150-467) -------------------------------------------------------------
static void pick_min_missing(const VarMer &vm, bool le_ref, std::vector<int> &idxs) {   // shared by better/strict/loose
  const uint32_t refMissing = vm.numM(0);
  uint32_t numMissing = refMissing;
  for (size_t ii = 0; ii < vm.npaths(); ii++) {
    const uint32_t m = vm.numM(ii);
    if (m < numMissing) { numMissing = m; idxs.clear(); idxs.push_back((int)ii); }
    else if (m == numMissing && (le_ref ? m <= refMissing : m < refMissing)) idxs.push_back((int)ii);
  }
}

static void longest_of(const VarMer &vm, const std::vector<int> &idxs, std::string &out) {   // varMer.C:229-246,288-308
  int idx = idxs[0];
  if (idxs.size() > 1) {
    uint32_t best = (uint32_t)vm.seqLen(idx);
    for (size_t ii = 1; ii < idxs.size(); ii++) {
      uint32_t len = (uint32_t)vm.seqLen(idxs[ii]);
      if (len > best) { best = len; idx = idxs[ii]; }
    }
  }
  vm.homRecord(idx, out);
}

static void better_or_strict(const VarMer &vm, std::vector<int> &idxs, std::string &out) {   // varMer.C:199-309
  if (vm.npaths() == 0) return;
  idxs.clear();
  pick_min_missing(vm, false, idxs);
  if (idxs.empty()) return;
  longest_of(vm, idxs, out);
}

static void loose(const VarMer &vm, std::vector<int> &idxs, std::string &out) {              // varMer.C:311-391
  if (vm.npaths() == 0) return;
  idxs.clear();
  pick_min_missing(vm, true, idxs);
  if (idxs.empty()) return;
  if (idxs.size() == 1) { vm.homRecord(idxs[0], out); return; }
  if (idxs[0] == 0 && idxs.size() == 2) { vm.homRecord(idxs[1], out); return; }
  int maxVars = 0, maxIdx = idxs[0];
  for (size_t ii = 1; ii < idxs.size(); ii++) {
    int count = 0;
    const int *gp = vm.gtPath(idxs[ii]);
    for (size_t i = 0; i < vm.c.nvar; i++) if (gp[i] > 0) count++;
    if (count > maxVars) { maxVars = count; maxIdx = idxs[ii]; }
  }
  fprintf(stderr, "[ WARNING ] :: Multiple (%lu) alternate pathes detected in a path beginning with variant : %s",
          (unsigned long)idxs.size(), vm.c.posGt->gts[0]->record->save().c_str());
  fprintf(stderr, "[ WARNING ] :: Max. %d ALT variants selected\n", maxVars);
  vm.homRecord(maxIdx, out);
}

// The reference's tie-break container is declared multimap<double,int,greater<int>> (varMer.H:72): keys are
// compared after conversion to int, descending, and equal keys keep their insertion order (Q6).  Only its
// first two elements are read: they are picked here with the same comparisons.
static void best_variant(const VarMer &vm, uint32_t k, std::vector<int> &idxs, std::string &out) {   // varMer.C:396-467
  uint32_t numMissing = UINT32_MAX;
  idxs.clear();
  for (size_t ii = 0; ii < vm.npaths(); ii++) {
    const uint32_t m = vm.numM(ii);
    if (m == vm.seqLen(ii) - k + 1) continue;
    if (m < numMissing) { numMissing = m; idxs.clear(); idxs.push_back((int)ii); }
    else if (m == numMissing) idxs.push_back((int)ii);
  }
  if (numMissing == UINT32_MAX) return;
  if (idxs.size() == 1) { vm.homRecord(idxs[0], out); return; }
  if (idxs.size() > 1) {
    double k1 = vm.totdK(idxs[0]), k2 = vm.totdK(idxs[1]);
    int i1 = idxs[0], i2 = idxs[1];
    if ((int)k2 > (int)k1) { std::swap(k1, k2); std::swap(i1, i2); }
    for (size_t ii = 2; ii < idxs.size(); ii++) {
      const double kk = vm.totdK(idxs[ii]);
      if ((int)kk > (int)k1) { k2 = k1; i2 = i1; k1 = kk; i1 = idxs[ii]; }
      else if ((int)kk > (int)k2) { k2 = kk; i2 = idxs[ii]; }
    }
    if (k1 == k2) {
      if (vm.seqLen(i1) >= vm.seqLen(i2)) vm.hetRecord(i1, i2, out);
      else vm.hetRecord(i2, i1, out);
      return;
    }
    vm.homRecord(i1, out);
  }
}

static void best_filter(const VarMer &vm, uint32_t k, std::vector<int> &idxs, std::vector<int> &gtIdxs, std::string &out) {   // varMer.C:150-194
  uint32_t numMissing = UINT32_MAX;
  idxs.clear();
  for (size_t ii = 0; ii < vm.npaths(); ii++) {
    const uint32_t m = vm.numM(ii);
    if (m == vm.seqLen(ii) - k + 1) continue;
    if (m == 0) { idxs.push_back((int)ii); numMissing = 0; }
    if (m < numMissing) { numMissing = m; idxs.clear(); idxs.push_back((int)ii); }
    else if (m == numMissing) idxs.push_back((int)ii);
  }
  if (idxs.empty()) return;
  gtIdxs.clear();
  for (int idx : idxs) {
    const int *gp = vm.gtPath(idx);
    for (size_t i = 0; i < vm.c.nvar; i++)
      if (gp[i] > 0) gtIdxs.push_back((int)i);
  }
  std::sort(gtIdxs.begin(), gtIdxs.end());
  gtIdxs.erase(std::unique(gtIdxs.begin(), gtIdxs.end()), gtIdxs.end());
  for (int i : gtIdxs) out += vm.c.posGt->gts[i]->record->save();
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

  Batch B;
  Enumerator E(B);
  ScoreOut sc;
  Options sopt = opt;
  sopt.debug = debug != nullptr;                     // the per-window values are kept exactly when the debug records are written
  const KTable kt(&opt.km);
  std::vector<int> idxs, gtIdxs;                     // the selectors' lists, reused
  std::vector<uint32_t> alt_lo, alt_hi;
  const uint32_t K_PADD = opt.k - 1;
  for (size_t b0 = first; b0 < cluster_end;) {
    // the pre-roll clusters ride along with the first batch
    const size_t b1 = std::min(cluster_end, (b0 < cluster_begin ? cluster_begin : b0) + batch_clusters);
    B.clear();
    B.clusters.reserve(b1 - b0);
    for (size_t ci = b0; ci < b1; ci++) {
      PosGT *pg = clusters[ci];
      uint32_t rStart = pg->rStart, rEnd = pg->rEnd;
      if (rStart > K_PADD) rStart -= K_PADD; else rStart = 0;
      if (rEnd < seqLen - K_PADD) rEnd += K_PADD; else rEnd = (uint32_t)seqLen;        // (unsigned, as the reference)
      const bool keep = ci >= cluster_begin;
      if (rEnd < rStart || seqLen < rEnd) {                                            // dnaSeq::copy fails (sequence.C:521)
        if (keep) fprintf(stderr, "PANIC : Invalid region specified: %s : %u - %u\n", ident.c_str(), rStart, rEnd);
        continue;
      }
      if (pg->gts.size() > opt.comb) {
        if (keep) fprintf(stderr, "PANIC : Combination %s:%u-%u has too many variants ( found %lu > %u ) to evaluate. Consider filtering the vcf upfront. Skipping...\n",
                          ident.c_str(), rStart, rEnd, (unsigned long)pg->gts.size(), opt.comb);
        continue;
      }
      B.clusters.emplace_back();
      Cluster &c = B.clusters.back();
      c.posGt = pg; c.rStart = rStart; c.rEnd = rEnd; c.keep = keep;
      E.run(c, seq);
    }

    const size_t npaths = B.off.size() - 1;
    if (npaths > 0) {
      int st = scorer(B.paths, B.off, sc);
      if (st != 0) return st;
    }
    B.numM.assign(npaths, 0); B.totdK.assign(npaths, 0.0);
    if (sopt.debug) { B.kstrs.assign(npaths, std::vector<double>()); B.dkstrs.assign(npaths, std::vector<double>()); }

    for (const Cluster &c : B.clusters) {
      score(B, c, sopt, kt, sc.v_read, sc.v_asm, sc.valid, prob, alt_lo, alt_hi);
      if (!c.keep) continue;
      const VarMer vm{B, c};
      if (debug) {
        char buf[512];
        for (size_t idx = 0; idx < vm.npaths(); idx++) {
          snprintf(buf, sizeof buf, "%lu\t%s:%u-%u\t", (unsigned long)varMerId++, ident.c_str(), c.rStart, c.rEnd);
          *debug += buf; debug->append(vm.seq(idx), vm.seqLen(idx));
          snprintf(buf, sizeof buf, "\t%u\t%.5f\t%.5f\t%.5f\t%.5f\t%.5f\t", vm.numM(idx), vm.minAbsK(idx), vm.maxAbsK(idx),
                   vm.medAbsK(idx), vm.avgAbsK(idx), vm.totdK(idx));
          *debug += buf;
          const int *gp = vm.gtPath(idx);
          for (size_t i = 0; i < c.nvar; i++) {
            const int alt = gp[i];
            if (alt > 0) {
              const GtAllele *g = c.posGt->gts[i];
              *debug += ident + " " + std::to_string(g->pos + 1) + " . "; *debug += g->alleles[0]; *debug += " "; *debug += g->alleles[alt]; *debug += " . PASS . GT 1/1  ";
            }
          }
          *debug += "\n";
        }
      }
      switch (opt.mode) {
        case POLISH: best_variant(vm, opt.k, idxs, result); break;
        case BETTER: better_or_strict(vm, idxs, result); break;
        case STRICT: better_or_strict(vm, idxs, result); break;
        case LOOSE:  loose(vm, idxs, result); break;
        default:     best_filter(vm, opt.k, idxs, gtIdxs, result); break;
      }
    }
    b0 = b1;
  }
  return 0;
}

}  // namespace mfnv
