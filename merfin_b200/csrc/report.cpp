// Host replay of the K* arithmetic from integer (v_read, v_asm) pairs.  The device produces
// integers only; every double below is computed with the expressions the reference uses
// (cited per function), with the same compiler family and no FMA contraction, so K* bins are
// bit-exact and QV sums differ from the reference only by summation order.
#include "common.h"

#include <cmath>
#include <vector>

extern "C" void mfn_getK(const mfn_kmodel *km, uint32_t v_read, uint32_t v_asm,
                         double *readK, double *asmK, double *prob) {
  // merfin-globals.C:66-98
  double rk = 0.0, pr = 1.0;
  const double peak = km ? km->peak : 0.0;
  if (v_read == 0)                 rk = 0;
  else if ((double)v_read < peak)  rk = 1;
  else                             rk = round((double)v_read / peak);
  if (km && v_read > 0 && v_read <= km->nprob) {
    rk = (double)km->probK[v_read - 1];
    pr = km->probP[v_read - 1];
  }
  *readK = rk;
  *asmK = (double)v_asm;
  *prob = pr;
}

extern "C" double mfn_getKmetric(double readK, double asmK) {
  // merfin-globals.H:248-261
  if (readK == 0) return 0;
  if (asmK > readK) return (asmK / readK - 1) * -1;
  if (asmK < readK) return readK / asmK - 1;
  return 0;
}

// (uint32)(((num / den - 1) + 0.1) / 0.2) as written at merfin-histogram.C:72,85.  When the
// assembly multiplicity is 0 the expression is +inf; the reference's conversion is undefined
// in C++ and yields 0 on x86-64, which is what the golden outputs hold.
static uint32_t kstar_bin(double num, double den) {
  double x = ((num / den - 1) + 0.1) / 0.2;
  if (!(x < 4294967296.0)) return 0;
  return (uint32_t)x;
}

static double histoQV(double kval, double ktot, uint32_t k) {
  // merfin-histogram.C:22-31
  double base = kval / ktot;
  double kinv = 1.0 / k;
  return -10.0 * log10(1.0 - pow(1.0 - base, kinv));
}

extern "C" int mfn_hist_report(const mfn_pair_count *pairs, uint64_t npairs, const mfn_kmodel *km, uint32_t k,
                               uint64_t *undr, uint64_t *over, uint32_t cap, mfn_hist_summary *sum) {
  if ((npairs && !pairs) || !sum || !km) return MFN_ERR_INVALID;
  std::vector<uint64_t> U, O;
  uint64_t kasm = 0, kmissing = 0;
  double koverCpy = 0.0;
  for (uint64_t i = 0; i < npairs; i++) {
    double readK, asmK, prob;
    mfn_getK(km, pairs[i].v_read, pairs[i].v_asm, &readK, &asmK, &prob);
    const uint64_t c = pairs[i].count;
    kasm += c;
    if (readK == 0) { kmissing += c; continue; }                     // merfin-histogram.C:66-69
    if (asmK > readK) {                                              // :71-82
      uint32_t idx = kstar_bin(asmK, readK);
      if (idx >= U.size()) U.resize((size_t)idx + 1, 0);
      U[idx] += c;
      koverCpy += ((1.0 - readK / asmK) * prob) * (double)c;
    } else {                                                         // :84-90
      uint32_t idx = kstar_bin(readK, asmK);
      if (idx >= O.size()) O.resize((size_t)idx + 1, 0);
      O[idx] += c;
    }
  }
  sum->kasm = kasm; sum->kmissing = kmissing; sum->koverCpy = koverCpy;
  sum->qv_missing = histoQV((double)kmissing, (double)kasm, k);             // :171
  sum->qv_star = histoQV((double)kmissing + koverCpy, (double)kasm, k);     // :172
  sum->n_undr = (uint32_t)U.size(); sum->n_over = (uint32_t)O.size();
  if (U.size() > cap || O.size() > cap || (!undr && !U.empty()) || (!over && !O.empty())) return MFN_ERR_CAPACITY;
  for (size_t i = 0; i < U.size(); i++) undr[i] = U[i];
  for (size_t i = 0; i < O.size(); i++) over[i] = O[i];
  return MFN_OK;
}

extern "C" int mfn_hist_format(const uint64_t *undr, uint32_t n_undr, const uint64_t *over, uint32_t n_over,
                               char *buf, uint64_t buflen, uint64_t *needed) {
  // merfin-histogram.C:151-163
  uint64_t need = 0;
  char line[96];
  auto emit = [&](double x, uint64_t c) {
    int n = snprintf(line, sizeof line, "%.1f\t%lu\n", x, (unsigned long)c);
    if (buf && need + n < buflen) memcpy(buf + need, line, n);
    need += n;
  };
  for (uint64_t ii = n_undr; ii-- > 1;)
    if (undr[ii] > 0) emit((double)ii * -0.2, undr[ii]);
  emit(0.0, (n_undr ? undr[0] : 0) + (n_over ? over[0] : 0));
  for (uint64_t ii = 1; ii < n_over; ii++)
    if (over[ii] > 0) emit((double)ii * 0.2, over[ii]);
  if (buf && need < buflen) buf[need] = 0;
  if (needed) *needed = need;
  return (buf && need >= buflen) ? MFN_ERR_CAPACITY : MFN_OK;
}

extern "C" int mfn_completeness_report(const mfn_pair_count *pairs, uint64_t npairs, const mfn_kmodel *km,
                                       double *total, double *undr) {
  // merfin-completeness.C:79-117: readK and asmK are integer-valued, so the sums are exact
  // (< 2^53) and independent of the order the reference adds them in.
  if ((npairs && !pairs) || !km || !total || !undr) return MFN_ERR_INVALID;
  double t = 0, u = 0;
  for (uint64_t i = 0; i < npairs; i++) {
    double readK, asmK, prob;
    mfn_getK(km, pairs[i].v_read, pairs[i].v_asm, &readK, &asmK, &prob);
    t += readK * (double)pairs[i].count;
    if (readK > asmK) u += (readK - asmK) * (double)pairs[i].count;
  }
  *total = t; *undr = u;
  return MFN_OK;
}

// ---------------------------------------------------------------------------------------------
// -dump text (merfin-dump.C:87-93): "%s\t%lu\t%.2f\t%.2f\t%.2f\n" of (readK, asmK, k*) for every
// valid window whose three numbers are not all zero.  The reference formats every line with
// fprintf in its single writer thread; here the positions are cut into chunks that worker threads
// format side by side (the three %.2f fields depend only on the integer pair, so each thread keeps
// the formatted tail of every pair it has seen), and the chunks are written in order.
// ---------------------------------------------------------------------------------------------
#include <unistd.h>

#include <algorithm>
#include <string>
#include <thread>
#include <unordered_map>

namespace {
struct DumpChunk { std::string text; uint64_t kasm = 0, kmiss = 0; };

void format_chunk(const char *ident, size_t ident_len, uint64_t p0, uint64_t p1, const uint32_t *vr, const uint32_t *va,
                  const uint8_t *ok, const mfn_kmodel *km, bool want_text, DumpChunk &out) {
  std::unordered_map<uint64_t, std::string> tails;        // pair -> "\t%.2f\t%.2f\t%.2f\n", empty string = line suppressed
  std::unordered_map<uint64_t, bool> miss;
  out.text.clear();
  if (want_text) out.text.reserve((size_t)(p1 - p0) * (ident_len + 28));
  char num[24];
  for (uint64_t p = p0; p < p1; p++) {
    if (!ok[p]) continue;
    const uint64_t key = ((uint64_t)vr[p] << 32) | va[p];
    auto it = tails.find(key);
    if (it == tails.end()) {
      double readK, asmK, prob;
      mfn_getK(km, vr[p], va[p], &readK, &asmK, &prob);
      const double ks = mfn_getKmetric(readK, asmK);
      char buf[128];
      int n = 0;
      if (readK != 0.0 || asmK != 0.0 || ks != 0.0) n = snprintf(buf, sizeof buf, "\t%.2f\t%.2f\t%.2f\n", readK, asmK, ks);
      it = tails.emplace(key, std::string(buf, (size_t)n)).first;
      miss[key] = (readK == 0);
    }
    out.kasm++;
    if (miss[key]) out.kmiss++;
    if (!want_text || it->second.empty()) continue;
    out.text.append(ident, ident_len);
    out.text.push_back('\t');
    uint64_t v = p;                                     // position inside the record, as decimal
    int d = 0;
    do { num[d++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (d) out.text.push_back(num[--d]);
    out.text.append(it->second);
  }
}
}  // namespace

// Format one record's -dump lines into file descriptor fd (fd < 0: count only, as -skipMissing does)
// with up to `threads` worker threads.  v_read / v_asm / valid are the record's slices of mfn_dump's
// output (index = position in the record).  Returns the record's kasm / kmissing in *counts.
extern "C" int mfn_dump_write(int fd, const char *ident, uint64_t n, const uint32_t *v_read, const uint32_t *v_asm,
                              const uint8_t *valid, const mfn_kmodel *km, int threads, mfn_seq_counts *counts) {
  if (!ident || (n && (!v_read || !v_asm || !valid))) return MFN_ERR_INVALID;
  const size_t il = strlen(ident);
  const uint64_t CH = 1 << 20;
  const uint64_t nch = (n + CH - 1) / CH;
  unsigned T = threads > 0 ? (unsigned)threads : std::max(1u, std::thread::hardware_concurrency());
  T = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(T, nch));
  uint64_t kasm = 0, kmiss = 0;
  std::vector<DumpChunk> round(T);
  for (uint64_t c0 = 0; c0 < nch; c0 += T) {
    const unsigned m = (unsigned)std::min<uint64_t>(T, nch - c0);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < m; t++)
      th.emplace_back([&, t]() {
        const uint64_t a = (c0 + t) * CH, b = std::min(n, a + CH);
        format_chunk(ident, il, a, b, v_read, v_asm, valid, km, fd >= 0, round[t]);
      });
    for (auto &x : th) x.join();
    for (unsigned t = 0; t < m; t++) {
      kasm += round[t].kasm; kmiss += round[t].kmiss;
      const char *p = round[t].text.data();
      size_t left = fd >= 0 ? round[t].text.size() : 0;
      while (left) {
        const ssize_t w = write(fd, p, left);
        if (w <= 0) { mfn_set_error("-dump: write failed"); return MFN_ERR_IO; }
        p += w; left -= (size_t)w;
      }
    }
  }
  if (counts) { counts->kasm = kasm; counts->kmissing = kmiss; }
  return MFN_OK;
}
