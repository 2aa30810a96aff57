#include "dump_driver.h"

#include <cstdlib>
#include <string>
#include <thread>
#include <vector>

namespace {
struct Batch {
  size_t r0 = 0, r1 = 0;                       // records [r0, r1)
  std::vector<uint64_t> off;                   // local offsets, off[0] = 0
  std::vector<uint32_t> vr, va;
  std::vector<uint8_t> ok;
  int status = MFN_OK;
  std::string err;                             // mfn_last_error() is per thread: carried over from the worker
};
}  // namespace

int dump_records(const SeqSet &S, const mfn_kmodel *km, int fd, int threads, uint64_t batch_bases,
                 const DumpBatchFn &pairs, FILE *log, std::string *err) {
  if (batch_bases == 0) {
    const char *env = getenv("MFN_DUMP_BATCH");
    batch_bases = env ? strtoull(env, nullptr, 10) : (32ull << 20);
    if (batch_bases == 0) batch_bases = 1;
  }
  const size_t nrec = S.nseq();
  Batch slot[2];
  // fills slot b with the records starting at r0; returns the first record of the next batch
  auto compute = [&](Batch &B, size_t r0) {
    size_t r1 = r0;
    uint64_t n = 0;
    while (r1 < nrec && (r1 == r0 || n < batch_bases)) n += S.length(r1++);
    B.r0 = r0; B.r1 = r1;
    B.off.resize(r1 - r0 + 1);
    for (size_t i = r0; i <= r1; i++) B.off[i - r0] = S.off[i] - S.off[r0];
    if (B.vr.size() < n + 1) { B.vr.resize(n + 1); B.va.resize(n + 1); B.ok.resize(n + 1); }
    B.status = n ? pairs(S.bases.data() + S.off[r0], B.off.data(), (uint32_t)(r1 - r0), B.vr.data(), B.va.data(), B.ok.data()) : MFN_OK;
    if (B.status != MFN_OK) B.err = mfn_last_error();
  };
  uint64_t totMissing = 0, totAsm = 0;
  int cur = 0;
  if (nrec) compute(slot[0], 0);
  for (size_t r = 0; r < nrec;) {
    Batch &B = slot[cur];
    if (B.status != MFN_OK) { if (err) *err = B.err; return B.status; }
    // the next batch is computed while this one is formatted and written
    std::thread next;
    if (B.r1 < nrec) next = std::thread(compute, std::ref(slot[cur ^ 1]), B.r1);
    int st = MFN_OK;
    for (size_t i = B.r0; i < B.r1 && st == MFN_OK; i++) {
      mfn_seq_counts c{};
      const uint64_t a = B.off[i - B.r0], n = B.off[i - B.r0 + 1] - a;
      st = mfn_dump_write(fd, S.ident[i].c_str(), n, B.vr.data() + a, B.va.data() + a, B.ok.data() + a, km, threads, &c);
      if (st != MFN_OK) break;
      totMissing += c.kmissing; totAsm += c.kasm;
      if (log) fprintf(log, "%s\t%lu\t%lu\t%lu\n", S.ident[i].c_str(), (unsigned long)c.kmissing, (unsigned long)totMissing, (unsigned long)totAsm);
    }
    if (next.joinable()) next.join();
    if (st != MFN_OK) { if (err) *err = mfn_last_error(); return st; }
    r = B.r1;
    cur ^= 1;
  }
  return MFN_OK;
}
