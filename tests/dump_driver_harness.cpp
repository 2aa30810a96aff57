// Test harness: the CLI's -dump driver (merfin_b200/csrc/host/dump_driver.cpp: batching, the overlap of
// "compute the next batch" with "format this one", per-record log lines) with the CPU oracle standing in
// for mfn_dump, so the control logic is checked against the reference's golden -dump files without a GPU.
// Test infrastructure only.
//   dump_driver_harness <asm.fa> <reads.meryl> <asm.meryl> <peak> <out.dump> <out.perseq> <batch_bases> <threads>
#include <cstdio>
#include <cstdlib>
#include <fcntl.h>
#include <string>
#include <unistd.h>

#include "../merfin_b200/csrc/host/dump_driver.h"
#include "../oracle/mfn_oracle.h"

int main(int argc, char **argv) {
  if (argc < 9) return 2;
  SeqSet S; std::string err;
  if (!load_sequences(argv[1], S, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
  char eb[256];
  o_db *rd = o_db_load(argv[2], eb, sizeof eb), *ad = o_db_load(argv[3], eb, sizeof eb);
  if (!rd || !ad) { fprintf(stderr, "%s\n", eb); return 1; }
  o_lookup *R = o_lookup_build(rd, 0, 0xffffffffu), *A = o_lookup_build(ad, 0, 0xffffffffu);
  mfn_kmodel km = {atof(argv[4]), 0, nullptr, nullptr};
  int calls = 0;
  DumpBatchFn pairs = [&](const char *bases, const uint64_t *off, uint32_t nrec, uint32_t *vr, uint32_t *va, uint8_t *ok) {
    calls++;
    for (uint32_t i = 0; i < nrec; i++) o_pairs(R, A, bases + off[i], off[i + 1] - off[i], vr + off[i], va + off[i], ok + off[i]);
    return 0;
  };
  const int fd = open(argv[5], O_WRONLY | O_CREAT | O_TRUNC, 0644);
  FILE *log = fopen(argv[6], "w");
  const int st = dump_records(S, &km, fd, atoi(argv[8]), strtoull(argv[7], nullptr, 10), pairs, log, &err);
  close(fd); fclose(log);
  printf("%d\n", calls);
  return st;
}
