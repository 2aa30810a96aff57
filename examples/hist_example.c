/* Minimal C host of the -hist path through the C ABI (include/merfin_b200.h): load the two meryl
 * databases, fuse them, evaluate a FASTA-less batch of sequences given as host buffers, print
 * the summary block merfin prints (merfin-histogram.C:167-172).
 *
 *   gcc -std=c99 -Iinclude examples/hist_example.c -Lmerfin_b200 -lmerfin_b200 -Wl,-rpath,$PWD/merfin_b200 -o hist_example
 *   ./hist_example reads.meryl asm.meryl 15 ACGTTGCA... [more sequences]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "merfin_b200.h"

static void die(const char *what, int st) {
  fprintf(stderr, "%s: %s (%s)\n", what, mfn_strerror(st), mfn_last_error());
  exit(1);
}

int main(int argc, char **argv) {
  if (argc < 5) { fprintf(stderr, "usage: %s reads.meryl asm.meryl peak SEQUENCE...\n", argv[0]); return 2; }
  mfn_ctx *ctx = NULL;
  mfn_table *reads = NULL, *asmt = NULL;
  int st;
  if ((st = mfn_ctx_create(&ctx, 0)) != MFN_OK) die("mfn_ctx_create", st);
  if ((st = mfn_table_load(ctx, argv[1], 0, 0xffffffffu, &reads)) != MFN_OK) die(argv[1], st);
  if ((st = mfn_table_load(ctx, argv[2], 0, 0xffffffffu, &asmt)) != MFN_OK) die(argv[2], st);
  if ((st = mfn_tables_fuse(ctx, reads, asmt)) != MFN_OK) die("mfn_tables_fuse", st);

  /* concatenate the records: bases + nseq+1 offsets */
  uint32_t nseq = (uint32_t)(argc - 4), i;
  uint64_t *off = (uint64_t *)calloc(nseq + 1, sizeof *off);
  size_t total = 0;
  for (i = 0; i < nseq; i++) { off[i] = total; total += strlen(argv[4 + i]); }
  off[nseq] = total;
  char *bases = (char *)malloc(total + 1);
  for (i = 0; i < nseq; i++) memcpy(bases + off[i], argv[4 + i], strlen(argv[4 + i]));

  mfn_kmodel km;
  memset(&km, 0, sizeof km);
  km.peak = atof(argv[3]);

  mfn_seq_counts *per_seq = (mfn_seq_counts *)calloc(nseq, sizeof *per_seq);
  uint64_t cap = 1 << 16, npairs = 0;
  mfn_pair_count *pairs = (mfn_pair_count *)malloc(cap * sizeof *pairs);
  st = mfn_hist(ctx, reads, asmt, bases, off, nseq, &km, per_seq, pairs, cap, &npairs);
  if (st == MFN_ERR_CAPACITY && npairs > cap) {          /* the joint histogram has more distinct pairs: retry */
    cap = npairs;
    pairs = (mfn_pair_count *)realloc(pairs, cap * sizeof *pairs);
    st = mfn_hist(ctx, reads, asmt, bases, off, nseq, &km, per_seq, pairs, cap, &npairs);
  }
  if (st != MFN_OK) die("mfn_hist", st);

  uint32_t k = 0;
  mfn_table_info_t info;
  if (mfn_table_info(reads, &info) == MFN_OK) k = info.k;
  uint64_t undr[4096], over[4096];
  mfn_hist_summary sum;
  if ((st = mfn_hist_report(pairs, npairs, &km, k, undr, over, 4096, &sum)) != MFN_OK) die("mfn_hist_report", st);
  fprintf(stderr, "K-mers not found in reads (missing) : %llu\n", (unsigned long long)sum.kmissing);
  fprintf(stderr, "K-mers overly represented in assembly: %.2f\n", sum.koverCpy);
  fprintf(stderr, "K-mers found in the assembly: %llu\n", (unsigned long long)sum.kasm);
  fprintf(stderr, "Missing QV: %.2f\nMerfin QV*: %.2f\n", sum.qv_missing, sum.qv_star);

  free(pairs); free(per_seq); free(bases); free(off);
  mfn_table_free(reads); mfn_table_free(asmt);
  mfn_ctx_destroy(ctx);
  return 0;
}
