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
// fprintf in its single writer thread.  Here:
//   * the three %.2f fields depend only on the integer pair, so the formatted tail of a pair is
//     computed once and kept in a table shared by all threads and all calls with the same K* model;
//   * positions are cut into chunks that worker threads format side by side into reusable buffers
//     (decimal position kept as digits and incremented, fixed-width copies), while the calling
//     thread writes finished chunks in order.
// ---------------------------------------------------------------------------------------------
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>

namespace {

constexpr uint32_t TAIL_R = 2048, TAIL_A = 32;            // pairs kept in the shared table: v_read < TAIL_R, v_asm < TAIL_A
constexpr uint32_t TAIL_MAX = 60;                         // longest tail kept inline
constexpr uint64_t DUMP_CHUNK = 1 << 17;                  // positions per formatter task

struct Tail {                                             // 64 bytes
  std::atomic<uint8_t> state{0};                          // 0 empty, 1 being filled, 2 ready
  uint8_t len = 0;                                        // 0: the line is suppressed (all three numbers zero)
  uint8_t miss = 0;                                       // readK == 0
  char s[TAIL_MAX + 1];
};

struct LongTail { std::string s; bool miss = false; };

// "\t%.2f\t%.2f\t%.2f\n" of one pair; returns the length snprintf wanted
int make_tail(const mfn_kmodel *km, uint32_t vr, uint32_t va, char *buf, size_t cap, bool *miss) {
  double readK, asmK, prob;
  mfn_getK(km, vr, va, &readK, &asmK, &prob);
  const double ks = mfn_getKmetric(readK, asmK);
  *miss = (readK == 0);
  if (readK == 0.0 && asmK == 0.0 && ks == 0.0) { buf[0] = 0; return 0; }
  return snprintf(buf, cap, "\t%.2f\t%.2f\t%.2f\n", readK, asmK, ks);
}

struct TailTable {
  double peak = 0;
  std::vector<uint32_t> probK;
  std::vector<double> probP;
  std::unique_ptr<Tail[]> t{new Tail[(size_t)TAIL_R * TAIL_A]};
  bool same_model(const mfn_kmodel *km) const {
    if (memcmp(&peak, &km->peak, sizeof peak) != 0 || probK.size() != km->nprob) return false;
    return !km->nprob || (memcmp(probK.data(), km->probK, km->nprob * sizeof(uint32_t)) == 0 &&
                          memcmp(probP.data(), km->probP, km->nprob * sizeof(double)) == 0);
  }
};

// the table of the K* model in use (a call with another model installs a fresh one; calls still
// running keep theirs alive through the shared_ptr)
std::shared_ptr<TailTable> tail_table(const mfn_kmodel *km) {
  static std::mutex mu;
  static std::shared_ptr<TailTable> cur;
  std::lock_guard<std::mutex> g(mu);
  if (!cur || !cur->same_model(km)) {
    cur = std::make_shared<TailTable>();
    cur->peak = km->peak;
    if (km->nprob) { cur->probK.assign(km->probK, km->probK + km->nprob); cur->probP.assign(km->probP, km->probP + km->nprob); }
  }
  return cur;
}

// reusable text buffers (a fresh 6 MB allocation per chunk would spend its time in page faults)
struct BufferPool {
  std::mutex mu;
  std::vector<std::vector<char>> free_;
  std::vector<char> get() {
    std::lock_guard<std::mutex> g(mu);
    if (free_.empty()) return {};
    std::vector<char> b = std::move(free_.back());
    free_.pop_back();
    return b;
  }
  void put(std::vector<char> &&b) {
    std::lock_guard<std::mutex> g(mu);
    // keep what one call on every core can use again (two slots per worker), nothing oversized
    const size_t keep = 2 * (size_t)std::max(1u, std::thread::hardware_concurrency());
    if (free_.size() < keep && b.capacity() <= ((size_t)64 << 20)) free_.push_back(std::move(b));
  }
};
BufferPool &pool() { static BufferPool P; return P; }

struct DumpChunk { std::vector<char> text; size_t len = 0; uint64_t kasm = 0, kmiss = 0; };

struct ChunkFormatter {
  const mfn_kmodel *km;
  TailTable *table;
  std::unordered_map<uint64_t, LongTail> rare;            // pairs outside the shared table, or with a long tail
  Tail scratch;

  // tail text + flags of a pair
  // (inl: the text lives in a Tail, so a fixed 32-byte copy may read past its end)
  inline void lookup(uint32_t vr, uint32_t va, const char *&s, uint32_t &len, bool &miss, bool &inl) {
    inl = true;
    if (vr < TAIL_R && va < TAIL_A) {
      Tail &e = table->t[(size_t)va * TAIL_R + vr];
      uint8_t st = e.state.load(std::memory_order_acquire);
      if (st == 2) {
        if (e.len <= TAIL_MAX) { s = e.s; len = e.len; miss = e.miss; return; }
      } else {
        char buf[256];
        bool m;
        const int n = make_tail(km, vr, va, buf, sizeof buf, &m);
        uint8_t expect = 0;
        if (st == 0 && e.state.compare_exchange_strong(expect, 1, std::memory_order_acq_rel)) {
          e.miss = m;
          e.len = n <= (int)TAIL_MAX ? (uint8_t)n : (uint8_t)255;       // 255: too long, kept per thread
          if (n <= (int)TAIL_MAX) memcpy(e.s, buf, (size_t)n + 1);
          e.state.store(2, std::memory_order_release);
          if (n <= (int)TAIL_MAX) { s = e.s; len = (uint32_t)n; miss = m; return; }
        } else if (n <= (int)TAIL_MAX) {
          // another thread is filling the entry right now: use a private copy this once
          scratch.len = (uint8_t)n; scratch.miss = m;
          memcpy(scratch.s, buf, (size_t)n + 1);
          s = scratch.s; len = (uint32_t)n; miss = m;
          return;
        }
      }
    }
    inl = false;
    const uint64_t key = ((uint64_t)vr << 32) | va;
    auto it = rare.find(key);
    if (it == rare.end()) {
      std::string buf(64, '\0');
      bool m;
      int n = make_tail(km, vr, va, &buf[0], buf.size(), &m);
      if (n >= (int)buf.size()) { buf.assign((size_t)n + 1, '\0'); n = make_tail(km, vr, va, &buf[0], buf.size(), &m); }
      buf.resize((size_t)n);
      it = rare.emplace(key, LongTail{std::move(buf), m}).first;
    }
    s = it->second.s.data(); len = (uint32_t)it->second.s.size(); miss = it->second.miss;
  }

  void run(const char *ident, size_t il, uint64_t p0, uint64_t p1, const uint32_t *vr, const uint32_t *va, const uint8_t *ok,
           bool want_text, DumpChunk &out) {
    out.len = 0; out.kasm = 0; out.kmiss = 0;
    const size_t slack = il + 64 + 256;                   // one line with a short tail, plus the over-copies below
    if (want_text && out.text.size() < (size_t)(p1 - p0) * (il + 36) + slack) out.text.resize((size_t)(p1 - p0) * (il + 36) + slack);
    char *base = out.text.data(), *w = base, *lim = base + out.text.size();
    char idbuf[64] = {0};                                 // short names are copied with a fixed width
    const bool short_id = il <= 32;
    if (short_id) memcpy(idbuf, ident, il);
    // decimal digits of the position, right-aligned so that the last digit is dig[31]
    char dig[64] = {0};
    uint32_t nd = 0;
    uint64_t cur = 0;
    bool have_cur = false;
    uint64_t kasm = 0, kmiss = 0;
    for (uint64_t p = p0; p < p1; p++) {
      if (!ok[p]) continue;
      const char *ts; uint32_t tl; bool miss, inl;
      lookup(vr[p], va[p], ts, tl, miss, inl);
      kasm++;
      kmiss += miss;
      if (!want_text || tl == 0) continue;
      if (have_cur && p == cur + 1) {                     // next position: increment the digits
        int i = 31;
        while (dig[i] == '9') { dig[i] = '0'; i--; if (31 - i == (int)nd) { dig[i] = '0'; nd++; } }
        dig[i]++;
      } else {
        uint64_t v = p;
        nd = 0;
        do { dig[31 - nd++] = (char)('0' + v % 10); v /= 10; } while (v);
      }
      cur = p; have_cur = true;
      if ((size_t)(lim - w) < il + tl + slack) {          // rare: long tails
        const size_t used = (size_t)(w - base);
        out.text.resize(out.text.size() * 2 + il + tl + slack);
        base = out.text.data(); w = base + used; lim = base + out.text.size();
      }
      if (short_id) memcpy(w, idbuf, 32); else memcpy(w, ident, il);
      w += il;
      *w++ = '\t';
      memcpy(w, dig + 32 - nd, 20);                       // nd <= 20; what follows the digits is overwritten
      w += nd;
      if (inl && tl <= 32) memcpy(w, ts, 32); else memcpy(w, ts, tl);
      w += tl;
    }
    out.len = (size_t)(w - base);
    out.kasm = kasm; out.kmiss = kmiss;
  }
};

}  // namespace

// Format one record's -dump lines into file descriptor fd (fd < 0: count only, as -skipMissing does)
// with up to `threads` worker threads.  v_read / v_asm / valid are the record's slices of mfn_dump's
// output (index = position in the record).  Returns the record's kasm / kmissing in *counts.
extern "C" int mfn_dump_write(int fd, const char *ident, uint64_t n, const uint32_t *v_read, const uint32_t *v_asm,
                              const uint8_t *valid, const mfn_kmodel *km, int threads, mfn_seq_counts *counts) {
  if (!ident || !km || (n && (!v_read || !v_asm || !valid))) return MFN_ERR_INVALID;
  const size_t il = strlen(ident);
  const uint64_t CH = DUMP_CHUNK;
  const uint64_t nch = (n + CH - 1) / CH;
  unsigned T = threads > 0 ? (unsigned)threads : std::max(1u, std::thread::hardware_concurrency());
  T = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(T, nch));
  const bool want_text = fd >= 0;
  std::shared_ptr<TailTable> table = tail_table(km);
  uint64_t kasm = 0, kmiss = 0;
  auto write_all = [&](const char *p, size_t left) -> bool {
    while (left) {
      const ssize_t w = write(fd, p, left);
      if (w <= 0) { mfn_set_error("-dump: write failed"); return false; }
      p += w; left -= (size_t)w;
    }
    return true;
  };

  if (T == 1) {                                            // small records: no threads
    ChunkFormatter F{km, table.get()};
    DumpChunk c;
    c.text = pool().get();
    bool ok = true;
    for (uint64_t i = 0; i < nch && ok; i++) {
      F.run(ident, il, i * CH, std::min(n, (i + 1) * CH), v_read, v_asm, valid, want_text, c);
      kasm += c.kasm; kmiss += c.kmiss;
      if (want_text) ok = write_all(c.text.data(), c.len);
    }
    pool().put(std::move(c.text));
    if (!ok) return MFN_ERR_IO;
    if (counts) { counts->kasm = kasm; counts->kmissing = kmiss; }
    return MFN_OK;
  }

  // chunk c is formatted by worker c % T into slot c % (2 T); the caller writes the chunks in order
  // and hands the slots back, so every worker formats one chunk ahead of the writer
  const unsigned NS = 2 * T;
  std::vector<DumpChunk> slot(NS);
  for (auto &s : slot) s.text = pool().get();
  std::mutex mu;
  std::condition_variable cv;
  std::vector<uint8_t> ready(nch, 0);
  uint64_t written = 0;                                    // chunks below this are written: their slots are free
  bool failed = false;
  std::vector<std::thread> th;
  for (unsigned t = 0; t < T; t++)
    th.emplace_back([&, t]() {
      ChunkFormatter F{km, table.get()};
      for (uint64_t c = t; c < nch; c += T) {
        {
          std::unique_lock<std::mutex> lk(mu);
          cv.wait(lk, [&] { return failed || c < written + NS; });          // slot c % NS is free
          if (failed) return;
        }
        F.run(ident, il, c * CH, std::min(n, (c + 1) * CH), v_read, v_asm, valid, want_text, slot[c % NS]);
        {
          std::lock_guard<std::mutex> g(mu);
          ready[c] = 1;
        }
        cv.notify_all();
      }
    });
  for (uint64_t c = 0; c < nch; c++) {
    {
      std::unique_lock<std::mutex> lk(mu);
      cv.wait(lk, [&] { return ready[c] != 0; });
    }
    DumpChunk &d = slot[c % NS];
    kasm += d.kasm; kmiss += d.kmiss;
    const bool ok = !want_text || write_all(d.text.data(), d.len);
    {
      std::lock_guard<std::mutex> g(mu);
      written = c + 1;
      if (!ok) failed = true;
    }
    cv.notify_all();
    if (!ok) break;
  }
  for (auto &x : th) x.join();
  for (auto &s : slot) pool().put(std::move(s.text));
  if (failed) return MFN_ERR_IO;
  if (counts) { counts->kasm = kasm; counts->kmissing = kmiss; }
  return MFN_OK;
}
