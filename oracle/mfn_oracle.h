/*
 * mfn_oracle.h -- CPU restatement of merfin's per-base k-mer multiplicity path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under merfin_b200/ (the product) may
 * include, link or call this.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, and only as the
 * checker.
 *
 * Parity status: PINNED.  Every function here is checked in tests/ against
 * outputs of the unmodified reference binaries (oracle/_ref/merfin,
 * oracle/_ref/meryl, built by oracle/Makefile.ref from /root/reference) that
 * are committed under tests/golden/ together with the script that made them
 * (tests/golden/make_golden.py).
 *
 * All file:line citations are relative to /root/reference/.
 */
#ifndef MFN_ORACLE_H
#define MFN_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned __int128 o_u128;

/* ---- meryl on-disk database ------------------------- */

typedef struct {
  uint32_t  k;              /* (prefixSize + suffixSize) / 2   kmers-reader.C:168 */
  uint32_t  prefixSize, suffixSize, numFilesBits, numBlocksBits, flags, version;
  uint64_t  numUnique, numDistinct, numTotal;
  uint64_t  histLen;        /* value histogram from the master index           */
  uint64_t *histValue, *histCount;
  uint64_t  n;              /* k-mers actually decoded from the data files      */
  uint64_t *lo, *hi;        /* k-mer bits, low/high 64; ascending (file order)  */
  uint32_t *val;
} o_db;

o_db    *o_db_load(const char *dir, char *errbuf, int errlen);
void     o_db_free(o_db *db);

/* ---- exact lookup with merfin's -min/-max semantics (T1, H4, Q1) ------ */

typedef struct {
  uint32_t  k;
  uint32_t  minValue, maxValue, valueOffset, valueBits;
  uint64_t  n;
  uint64_t *lo, *hi;
  uint32_t *val;            /* stored value = original - valueOffset            */
} o_lookup;

o_lookup *o_lookup_build(const o_db *db, uint32_t minV, uint32_t maxV);
void      o_lookup_free(o_lookup *t);
uint32_t  o_lookup_value(const o_lookup *t, uint64_t lo, uint64_t hi);
void      o_lookup_values(const o_lookup *t, const uint64_t *lo, const uint64_t *hi,
                          uint64_t n, uint32_t *out);

/* ---- window iteration + probes (H1-H3) --------------------- */

/* One event per base index i >= k-1 of a record; arrays have L-k+1 entries
 * (0 when L < k).  Returns the number of events written. */
uint64_t  o_windows(const char *bases, uint64_t L, uint32_t k,
                    uint64_t *f_lo, uint64_t *f_hi, uint64_t *r_lo, uint64_t *r_hi,
                    uint8_t *valid);

/* (v_read, v_asm) per window event; invalid windows get valid=0 and 0,0. */
uint64_t  o_pairs(const o_lookup *R, const o_lookup *A, const char *bases, uint64_t L,
                  uint32_t *v_read, uint32_t *v_asm, uint8_t *valid);

/* ---- K* model (H6) ------------------------------------------------------ */

typedef struct {
  double          peak;
  uint32_t        nprob;       /* lines of the -prob table                     */
  const uint32_t *probK;
  const double   *probP;
} o_kmodel;

void      o_getK(const o_kmodel *m, uint32_t v_read, uint32_t v_asm,
                 double *readK, double *asmK, double *prob);
double    o_getKmetric(double readK, double asmK);

/* ---- -hist (H7, H8) ----------------------------------------------------- */

typedef struct {
  uint64_t  kasm, kmissing;
  double    koverCpy;
  uint32_t  undrMax, overMax;
  uint64_t *undr, *over;
} o_hist;

o_hist   *o_hist_new(void);
void      o_hist_free(o_hist *h);
/* processHistogram for one record; then outputHistogram merge into 'total'.
 * per-record results returned through rec (may be NULL). */
void      o_hist_record(const o_lookup *R, const o_lookup *A, const o_kmodel *m,
                        const char *bases, uint64_t L, o_hist *total, o_hist *rec);
/* Same accumulation from already-known integer pairs (count windows with that pair). */
void      o_hist_add_pair(o_hist *h, const o_kmodel *m, uint32_t v_read, uint32_t v_asm,
                          uint64_t count);
/* reportHistogram: writes the histogram file text into buf (returns length needed). */
uint64_t  o_hist_report(const o_hist *h, char *buf, uint64_t buflen);
double    o_histoQV(double kval, double ktot, uint32_t k);

/* ---- -dump (H9) --------------------------------------------------------- */

/* Per position arrays of length L+1 as in processDump; returns kasm, kmissing. */
void      o_dump_record(const o_lookup *R, const o_lookup *A, const o_kmodel *m,
                        const char *bases, uint64_t L,
                        double *readK, double *asmK, double *kmetric,
                        uint64_t *kasm, uint64_t *kmissing);
/* outputDump text for one record appended to buf; returns bytes needed. */
uint64_t  o_dump_format(const char *ident, uint64_t L, const double *readK, const double *asmK,
                        const double *kmetric, char *buf, uint64_t buflen);

/* ---- -completeness (H10) ------------------------------------------------ */

void      o_completeness(const o_db *reads, const o_db *asmdb, const o_kmodel *m,
                         double total64[64], double undr64[64],
                         double *total, double *undr);

/* ---- variant path scoring (H11) ------------------------------------------ */

/* One path string.  alt_idx/alt_len/alt_gt describe the alleles on the path
 * (varMer::idxPaths / lenPaths / gtPaths).  kstr and dk need L-k+1 doubles
 * (unused when filter_mode).  *prob_io carries the 'prob' local across calls
 * (Q5: it is only assigned on valid windows).  Returns numM. */
uint32_t  o_score_path(const o_lookup *R, const o_lookup *A, const o_kmodel *m,
                       const char *path, uint64_t L, int filter_mode,
                       const uint32_t *alt_idx, const uint32_t *alt_len, const int *alt_gt,
                       uint32_t nalt, double *kstr, double *dk, uint64_t *nwin, double *prob_io);

#ifdef __cplusplus
}
#endif
#endif
