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
