/*
 * mfn_oracle.c -- CPU restatement of merfin's per-base k-mer multiplicity path.
 *
 * TEST INFRASTRUCTURE ONLY (see mfn_oracle.h).  Plain C, single thread, written
 * for clarity: a k-mer lookup is a binary search over the whole sorted k-mer
 * list, which returns exactly what merylExactLookup::value() returns for a
 * (non-multiset) database without reproducing its bucket layout.
 *
 * Parity status: PINNED against the unmodified reference binaries, see
 * tests/golden/ and tests/test_oracle_golden.py.
 *
 * All file:line citations are relative to /root/reference/.
 */
#define _GNU_SOURCE
#include "mfn_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ======================================================================= */
/*  stuffedBits container (R3): bits.C:157-235 (load), bits.H:665-707       */
/*  (moveToNextBlock), bits.C:371-410 (getUnary), bits.C:523-572 (getBinary)*/
/* ======================================================================= */

typedef struct {
  uint32_t   nsub;
  uint64_t  *len;      /* bits in each sub-block           */
  uint64_t **dat;      /* words of each sub-block          */
  uint32_t   blk;      /* read head: sub-block              */
  uint64_t   pos;      /* read head: bit inside sub-block   */
} sbits;

static void sb_free(sbits *s) {
  if (!s) return;
  for (uint32_t i = 0; i < s->nsub; i++) free(s->dat[i]);
  free(s->dat); free(s->len); free(s);
}

/* Reads one dumped container from F; returns NULL at EOF. */
static sbits *sb_read(FILE *F) {
  uint64_t maxBits; uint32_t nsub, blocksMax;
  if (fread(&maxBits, 8, 1, F) != 1) return NULL;
  if (fread(&nsub, 4, 1, F) != 1) return NULL;
  if (fread(&blocksMax, 4, 1, F) != 1) return NULL;
  sbits *s = calloc(1, sizeof(sbits));
  s->nsub = nsub;
  s->len = calloc(nsub ? nsub : 1, 8);
  s->dat = calloc(nsub ? nsub : 1, sizeof(uint64_t *));
  uint64_t *bgn = calloc(nsub ? nsub : 1, 8);
  int ok = (fread(bgn, 8, nsub, F) == nsub) && (fread(s->len, 8, nsub, F) == nsub);
  free(bgn);
  for (uint32_t i = 0; ok && i < nsub; i++) {
    uint64_t nw = (s->len[i] + 63) / 64;
    s->dat[i] = calloc(nw + 2, 8);
    ok = fread(s->dat[i], 8, nw, F) == nw;
  }
  if (!ok) { sb_free(s); return NULL; }
  return s;
}

static uint64_t sb_total_bits(const sbits *s) {
  uint64_t t = 0;
  for (uint32_t i = 0; i < s->nsub && s->len[i] > 0; i++) t += s->len[i];
  return t;
}

/* A coded value never straddles sub-blocks: when it does not fit in what is
 * left, the head jumps to the start of the next one (bits.H:665-707). */
static void sb_fit(sbits *s, uint64_t width) {
  if (s->pos + width <= s->len[s->blk]) return;
  s->blk++; s->pos = 0;
}

static uint64_t sb_bit(sbits *s) {           /* MSB-first inside each u64 word */
  uint64_t w = s->dat[s->blk][s->pos >> 6];
  uint64_t b = (w >> (63 - (s->pos & 63))) & 1;
  s->pos++;
  return b;
}

static uint64_t sb_binary(sbits *s, uint32_t width) {
  if (width == 0) return 0;
  sb_fit(s, width);
  uint64_t v = 0;
  for (uint32_t i = 0; i < width; i++) v = (v << 1) | sb_bit(s);
  return v;
}

static uint64_t sb_unary(sbits *s) {         /* x zero bits then a one bit */
  sb_fit(s, 1);
  uint64_t v = 0;
  while (sb_bit(s) == 0) v++;
  return v;
}

/* ======================================================================= */
/*  meryl database: master index kmers-reader.C:127-177, stats              */
/*  kmers-histogram.C:180-207, file names kmers.C:29-55, blocks             */
/*  kmers-files.C:24-128                                                    */
/* ======================================================================= */

static void seterr(char *eb, int el, const char *msg, const char *arg) {
  if (eb && el > 0) snprintf(eb, el, "%s%s", msg, arg ? arg : "");
}

void o_db_free(o_db *db) {
  if (!db) return;
  free(db->histValue); free(db->histCount);
  free(db->lo); free(db->hi); free(db->val); free(db);
}

o_db *o_db_load(const char *dir, char *eb, int el) {
  char path[4096];
  snprintf(path, sizeof path, "%s/merylIndex", dir);
  FILE *F = fopen(path, "rb");
  if (!F) { seterr(eb, el, "cannot open ", path); return NULL; }
  sbits *mi = sb_read(F);
  fclose(F);
  if (!mi) { seterr(eb, el, "bad master index ", path); return NULL; }

  o_db *db = calloc(1, sizeof(o_db));
  uint64_t m1 = sb_binary(mi, 64), m2 = sb_binary(mi, 64);
  if (m1 != 0x646e496c7972656dull) { seterr(eb, el, "magic mismatch ", path); sb_free(mi); free(db); return NULL; }
  if      (m2 == 0x31302e765f5f7865ull) db->version = 1;
  else if (m2 == 0x32302e765f5f7865ull) db->version = 2;
  else if (m2 == 0x33302e765f5f7865ull) db->version = 3;
  else { seterr(eb, el, "version mismatch ", path); sb_free(mi); free(db); return NULL; }

  db->prefixSize    = sb_binary(mi, 32);
  db->suffixSize    = sb_binary(mi, 32);
  db->numFilesBits  = sb_binary(mi, 32);
  db->numBlocksBits = sb_binary(mi, 32);
  db->flags         = (db->version >= 2) ? sb_binary(mi, 32) : 0;
  db->k             = (db->prefixSize + db->suffixSize) / 2;

  if (db->version == 3) {                       /* kmers-histogram.C:180-207 */
    db->numUnique   = sb_binary(mi, 64);
    db->numDistinct = sb_binary(mi, 64);
    db->numTotal    = sb_binary(mi, 64);
    db->histLen     = sb_binary(mi, 64);
    db->histValue   = calloc(db->histLen + 1, 8);
    db->histCount   = calloc(db->histLen + 1, 8);
    for (uint64_t i = 0; i < db->histLen; i++) {
      db->histValue[i] = sb_binary(mi, 64);
      db->histCount[i] = sb_binary(mi, 64);
    }
  }
  sb_free(mi);

  /* data files, in file order == ascending k-mer order */
  uint64_t cap = 1 << 16;
  db->lo = malloc(cap * 8); db->hi = malloc(cap * 8); db->val = malloc(cap * 4);
  uint32_t nfiles = 1u << db->numFilesBits;
  for (uint32_t ff = 0; ff < nfiles; ff++) {
    char bits[80]; uint32_t bp = 0;
    bits[bp++] = '0'; bits[bp++] = 'x';
    for (int b = (int)db->numFilesBits - 1; b >= 0; b--) bits[bp++] = ((ff >> b) & 1) ? '1' : '0';
    bits[bp] = 0;
    snprintf(path, sizeof path, "%s/%s.merylData", dir, bits);
    FILE *D = fopen(path, "rb");
    if (!D) { seterr(eb, el, "cannot open ", path); o_db_free(db); return NULL; }
    sbits *blk;
    while ((blk = sb_read(D)) != NULL) {
      if (sb_total_bits(blk) == 0) { sb_free(blk); break; }
      uint64_t b1 = sb_binary(blk, 64), b2 = sb_binary(blk, 64);
      if (b1 != 0x7461446c7972656dull || b2 != 0x0a3030656c694661ull) {
        seterr(eb, el, "block magic mismatch in ", path); sb_free(blk); fclose(D); o_db_free(db); return NULL;
      }
      uint64_t blockPrefix = (uint32_t)sb_binary(blk, 64);
      uint64_t nKmers      = sb_binary(blk, 64);
      uint32_t kCode       = sb_binary(blk, 8);
      (void)                 sb_binary(blk, 32);          /* unaryBits: informational */
      uint32_t binaryBits  = sb_binary(blk, 32);
      (void)                 sb_binary(blk, 64);
      uint32_t cCode       = sb_binary(blk, 8);
      (void)                 sb_binary(blk, 64);
      (void)                 sb_binary(blk, 64);
      if (kCode != 1 || (cCode != 1 && cCode != 2)) {
        seterr(eb, el, "unknown block coding in ", path); sb_free(blk); fclose(D); o_db_free(db); return NULL;
      }
      while (db->n + nKmers > cap) {
        cap *= 2;
        db->lo = realloc(db->lo, cap * 8); db->hi = realloc(db->hi, cap * 8); db->val = realloc(db->val, cap * 4);
      }
      o_u128 high = 0;                                     /* kmers-files.C:88-107 */
      for (uint64_t i = 0; i < nKmers; i++) {
        high += sb_unary(blk);
        uint32_t ls = (binaryBits <= 64) ? 0 : binaryBits - 64;
        uint32_t rs = (binaryBits <= 64) ? binaryBits : 64;
        o_u128 suf = high;
        suf = (ls ? (suf << ls) : suf) | sb_binary(blk, ls);
        suf = ((rs == 64) ? (suf << 32) << 32 : (suf << rs)) | sb_binary(blk, rs);
        o_u128 km = ((o_u128)blockPrefix << db->suffixSize) | suf;
        db->lo[db->n + i] = (uint64_t)km;
        db->hi[db->n + i] = (uint64_t)(km >> 64);
      }
      for (uint64_t i = 0; i < nKmers; i++)                /* kmers-files.C:116-124 */
        db->val[db->n + i] = (uint32_t)sb_binary(blk, cCode == 1 ? 32 : 64);
      db->n += nKmers;
      sb_free(blk);
    }
    fclose(D);
  }

  if (db->version != 3) {
    /* v01/v02 master indexes carry an older statistics layout; rebuild the
     * value histogram from the data instead (same content for a consistent DB). */
    uint32_t vmax = 0;
    for (uint64_t i = 0; i < db->n; i++) if (db->val[i] > vmax) vmax = db->val[i];
    uint64_t *cnt = calloc((uint64_t)vmax + 2, 8);
    for (uint64_t i = 0; i < db->n; i++) cnt[db->val[i]]++;
    uint64_t nd = 0;
    for (uint64_t v = 0; v <= vmax; v++) if (cnt[v]) nd++;
    db->histLen = nd;
    db->histValue = calloc(nd + 1, 8); db->histCount = calloc(nd + 1, 8);
    nd = 0;
    for (uint64_t v = 0; v <= vmax; v++) if (cnt[v]) { db->histValue[nd] = v; db->histCount[nd++] = cnt[v]; }
    free(cnt);
    db->numDistinct = db->n;
  }
  return db;
}

/* ======================================================================= */
/*  exact lookup: filter + offset kmers-exact.C:34-73,582-624; query        */
/*  kmers-lookup.H:334-380 (value() returns the STORED value, Q1)           */
/* ======================================================================= */

static uint32_t nbits64(uint64_t x) { uint32_t n = 0; while (x) { n++; x >>= 1; } return n; }

o_lookup *o_lookup_build(const o_db *db, uint32_t minV, uint32_t maxV) {
  o_lookup *t = calloc(1, sizeof(o_lookup));
  t->k = db->k;
  if (minV == 0) minV = 1;                                         /* kmers-exact.C:43 */
  if (maxV == 0xffffffffu)                                         /* kmers-exact.C:44 */
    maxV = db->histLen ? (uint32_t)db->histValue[db->histLen - 1] : 0;
  t->minValue = minV; t->maxValue = maxV;
  if (maxV >= minV) {                                              /* kmers-exact.C:57-61 */
    t->valueOffset = minV - 1;
    t->valueBits   = nbits64((uint64_t)maxV + 1 - minV);
  }
  t->lo = malloc((db->n + 1) * 8); t->hi = malloc((db->n + 1) * 8); t->val = malloc((db->n + 1) * 4);
  for (uint64_t i = 0; i < db->n; i++) {
    uint32_t v = db->val[i];
    if (v < minV || v > maxV) continue;                            /* kmers-exact.C:578-580 */
    t->lo[t->n] = db->lo[i]; t->hi[t->n] = db->hi[i];
    t->val[t->n] = v - t->valueOffset;                             /* kmers-exact.C:613 */
    t->n++;
  }
  return t;
}

void o_lookup_free(o_lookup *t) { if (t) { free(t->lo); free(t->hi); free(t->val); free(t); } }

uint32_t o_lookup_value(const o_lookup *t, uint64_t lo, uint64_t hi) {
  uint64_t a = 0, b = t->n;
  while (a < b) {
    uint64_t m = a + (b - a) / 2;
    if (t->hi[m] == hi && t->lo[m] == lo)
      return t->valueBits == 0 ? 1 : t->val[m];                    /* kmers-lookup.H:352-356 */
    if (t->hi[m] < hi || (t->hi[m] == hi && t->lo[m] < lo)) a = m + 1; else b = m;
  }
  return 0;
}

void o_lookup_values(const o_lookup *t, const uint64_t *lo, const uint64_t *hi, uint64_t n, uint32_t *out) {
  for (uint64_t i = 0; i < n; i++) out[i] = o_lookup_value(t, lo[i], hi ? hi[i] : 0);
}

/* ======================================================================= */
/*  windows: kmers-tiny.H:69-83 (encoding, addR/addL),                      */
/*  kmers-iterator.H:93-160 (nextBase / isValid / position)                 */
/* ======================================================================= */

static int is_acgt(char c) {
  return c == 'A' || c == 'a' || c == 'C' || c == 'c' || c == 'G' || c == 'g' || c == 'T' || c == 't';
}

typedef struct { o_u128 f, r; uint32_t load, k; o_u128 fullMask, leftMask; uint32_t leftShift; } o_iter;

static void it_init(o_iter *it, uint32_t k) {
  memset(it, 0, sizeof *it);
  it->k = k;
  it->fullMask  = (k == 64) ? ~(o_u128)0 : ((((o_u128)1) << (2 * k)) - 1);
  it->leftMask  = (((o_u128)1) << (2 * k - 2)) - 1;
  it->leftShift = 2 * k - 2;
}

/* consume base c; returns 1 when the k-mer ending here is valid */
static int it_push(o_iter *it, char c) {
  if (!is_acgt(c)) { it->load = 0; return 0; }          /* f, r keep their stale bases */
  o_u128 e = ((unsigned char)c >> 1) & 3;
  it->f = ((it->f << 2) & it->fullMask) | e;
  it->r = ((it->r >> 2) & it->leftMask) | ((e ^ 2) << it->leftShift);
  if (it->load < it->k) it->load++;
  return it->load == it->k;
}

uint64_t o_windows(const char *bases, uint64_t L, uint32_t k,
                   uint64_t *f_lo, uint64_t *f_hi, uint64_t *r_lo, uint64_t *r_hi, uint8_t *valid) {
  o_iter it; it_init(&it, k);
  uint64_t n = 0;
  for (uint64_t i = 0; i < L; i++) {
    int ok = it_push(&it, bases[i]);
    if (i + 1 < k) continue;                              /* silent preload of k-1 bases */
    if (f_lo) f_lo[n] = (uint64_t)it.f;  if (f_hi) f_hi[n] = (uint64_t)(it.f >> 64);
    if (r_lo) r_lo[n] = (uint64_t)it.r;  if (r_hi) r_hi[n] = (uint64_t)(it.r >> 64);
    if (valid) valid[n] = (uint8_t)ok;
    n++;
  }
  return n;
}

/* merfin-globals.C:101-110 -- f and r are looked up separately and summed (Q2) */
static void probe_pair(const o_lookup *R, const o_lookup *A, const o_iter *it, uint32_t *vr, uint32_t *va) {
  uint64_t fl = (uint64_t)it->f, fh = (uint64_t)(it->f >> 64);
  uint64_t rl = (uint64_t)it->r, rh = (uint64_t)(it->r >> 64);
  *vr = o_lookup_value(R, fl, fh) + o_lookup_value(R, rl, rh);
  *va = o_lookup_value(A, fl, fh) + o_lookup_value(A, rl, rh);
}

uint64_t o_pairs(const o_lookup *R, const o_lookup *A, const char *bases, uint64_t L,
                 uint32_t *v_read, uint32_t *v_asm, uint8_t *valid) {
  o_iter it; it_init(&it, R->k);
  uint64_t n = 0;
  for (uint64_t i = 0; i < L; i++) {
    int ok = it_push(&it, bases[i]);
    if (i + 1 < R->k) continue;
    uint32_t vr = 0, va = 0;
    if (ok) probe_pair(R, A, &it, &vr, &va);
    v_read[n] = vr; v_asm[n] = va; valid[n] = (uint8_t)ok;
    n++;
  }
  return n;
}

/* ======================================================================= */
/*  K* model: merfin-globals.C:66-98, merfin-globals.H:248-261              */
/* ======================================================================= */

void o_getK(const o_kmodel *m, uint32_t v_read, uint32_t v_asm, double *readK, double *asmK, double *prob) {
  *readK = 0.0;
  *asmK  = (double)v_asm;
  *prob  = 1.0;
  if      (v_read == 0)               *readK = 0;
  else if ((double)v_read < m->peak)  *readK = 1;
  else                                *readK = round((double)v_read / m->peak);
  if (v_read > 0 && v_read <= m->nprob) {
    *readK = (double)m->probK[v_read - 1];
    *prob  = m->probP[v_read - 1];
  }
}

double o_getKmetric(double readK, double asmK) {
  if (readK == 0)   return 0;
  if (asmK > readK) return (asmK / readK - 1) * -1;
  if (asmK < readK) return readK / asmK - 1;
  return 0;
}

/* ======================================================================= */
/*  -hist: merfin-histogram.C:22-31 (histoQV), 35-92 (process),             */
/*  96-136 (merge), 140-176 (report)                                        */
/* ======================================================================= */

o_hist *o_hist_new(void) { return calloc(1, sizeof(o_hist)); }
void    o_hist_free(o_hist *h) { if (h) { free(h->undr); free(h->over); free(h); } }

static void grow(uint64_t **a, uint32_t *max, uint32_t need) {
  if (need <= *max) return;
  uint32_t nm = *max ? *max : 1024;
  while (nm < need) nm += 1024;
  *a = realloc(*a, (uint64_t)nm * 8);
  memset(*a + *max, 0, (uint64_t)(nm - *max) * 8);
  *max = nm;
}

/* The bin index is a double -> uint32 conversion written exactly as the
 * reference writes it.  asmK == 0 makes the expression +inf; the conversion is
 * undefined in C and observed to give bin 0 with the reference binary (Q3). */
static uint32_t bin_of(double num, double den) {
  double x = ((num / den - 1) + 0.1) / 0.2;
  if (!(x < 4294967296.0)) return 0;
  return (uint32_t)x;
}

void o_hist_add_pair(o_hist *h, const o_kmodel *m, uint32_t v_read, uint32_t v_asm, uint64_t count) {
  double readK, asmK, prob;
  o_getK(m, v_read, v_asm, &readK, &asmK, &prob);
  h->kasm += count;
  if (readK == 0) { h->kmissing += count; return; }
  if (asmK > readK) {
    uint32_t idx = bin_of(asmK, readK);
    grow(&h->undr, &h->undrMax, idx + 1);
    h->undr[idx] += count;
    for (uint64_t c = 0; c < count; c++) h->koverCpy += (1.0 - readK / asmK) * prob;
  } else {
    uint32_t idx = bin_of(readK, asmK);
    grow(&h->over, &h->overMax, idx + 1);
    h->over[idx] += count;
  }
}

void o_hist_record(const o_lookup *R, const o_lookup *A, const o_kmodel *m,
                   const char *bases, uint64_t L, o_hist *total, o_hist *rec) {
  o_hist *s = o_hist_new();
  grow(&s->undr, &s->undrMax, 1024);
  grow(&s->over, &s->overMax, 1024);
  o_iter it; it_init(&it, R->k);
  for (uint64_t i = 0; i < L; i++) {
    int ok = it_push(&it, bases[i]);
    if (i + 1 < R->k || !ok) continue;
    uint32_t vr, va;
    probe_pair(R, A, &it, &vr, &va);
    o_hist_add_pair(s, m, vr, va, 1);
  }
  if (total) {                                            /* outputHistogram */
    total->kmissing += s->kmissing;
    total->kasm     += s->kasm;
    total->koverCpy += s->koverCpy;
    grow(&total->undr, &total->undrMax, s->undrMax > 2048 ? s->undrMax : 2048);
    grow(&total->over, &total->overMax, s->overMax > 2048 ? s->overMax : 2048);
    for (uint32_t i = 0; i < s->undrMax; i++) total->undr[i] += s->undr[i];
    for (uint32_t i = 0; i < s->overMax; i++) total->over[i] += s->over[i];
  }
  if (rec) { rec->kasm = s->kasm; rec->kmissing = s->kmissing; rec->koverCpy = s->koverCpy; }
  o_hist_free(s);
}

uint64_t o_hist_report(const o_hist *h, char *buf, uint64_t buflen) {
  uint64_t need = 0;
  char line[128];
#define EMIT(...) do { int n_ = snprintf(line, sizeof line, __VA_ARGS__); \
    if (buf && need + n_ < buflen) memcpy(buf + need, line, n_); need += n_; } while (0)
  for (uint64_t ii = h->undrMax; ii-- > 1; )
    if (h->undr[ii] > 0) EMIT("%.1f\t%lu\n", (double)ii * -0.2, (unsigned long)h->undr[ii]);
  EMIT("%.1f\t%lu\n", 0.0, (unsigned long)((h->undrMax ? h->undr[0] : 0) + (h->overMax ? h->over[0] : 0)));
  for (uint64_t ii = 1; ii < h->overMax; ii++)
    if (h->over[ii] > 0) EMIT("%.1f\t%lu\n", (double)ii * 0.2, (unsigned long)h->over[ii]);
#undef EMIT
  if (buf && need < buflen) buf[need] = 0;
  return need;
}

double o_histoQV(double kval, double ktot, uint32_t k) {
  double base = kval / ktot;
  double kinv = 1.0 / k;
  return -10.0 * log10(1.0 - pow(1.0 - base, kinv));
}

/* ======================================================================= */
/*  -dump: merfin-dump.C:20-104                                             */
/* ======================================================================= */

void o_dump_record(const o_lookup *R, const o_lookup *A, const o_kmodel *m,
                   const char *bases, uint64_t L, double *readK, double *asmK, double *kmetric,
                   uint64_t *kasm, uint64_t *kmissing) {
  for (uint64_t i = 0; i <= L; i++) readK[i] = asmK[i] = kmetric[i] = 0;
  *kasm = *kmissing = 0;
  o_iter it; it_init(&it, R->k);
  for (uint64_t i = 0; i < L; i++) {
    int ok = it_push(&it, bases[i]);
    if (i + 1 < R->k || !ok) continue;
    uint32_t vr, va; double rk, ak, pr;
    probe_pair(R, A, &it, &vr, &va);
    o_getK(m, vr, va, &rk, &ak, &pr);
    (*kasm)++;
    if (rk == 0) (*kmissing)++;
    uint64_t pp = i + 1 - R->k;                           /* kmers-iterator.H:160 */
    readK[pp] = rk; asmK[pp] = ak; kmetric[pp] = o_getKmetric(rk, ak);
  }
}

uint64_t o_dump_format(const char *ident, uint64_t L, const double *readK, const double *asmK,
                       const double *kmetric, char *buf, uint64_t buflen) {
  uint64_t need = 0;
  char line[1200];
  for (uint64_t pp = 0; pp < L; pp++) {
    if (readK[pp] == 0.0 && asmK[pp] == 0.0 && kmetric[pp] == 0.0) continue;
    int n = snprintf(line, sizeof line, "%s\t%lu\t%.2f\t%.2f\t%.2f\n", ident, (unsigned long)pp,
                     readK[pp], asmK[pp], kmetric[pp]);
    if (buf && need + n < buflen) memcpy(buf + need, line, n);
    need += n;
  }
  if (buf && need < buflen) buf[need] = 0;
  return need;
}

/* ======================================================================= */
/*  -completeness: merfin-completeness.C:48-144 (raw values, one strand, Q4)*/
/* ======================================================================= */

static int kcmp(uint64_t ah, uint64_t al, uint64_t bh, uint64_t bl) {
  if (ah != bh) return ah < bh ? -1 : 1;
  if (al != bl) return al < bl ? -1 : 1;
  return 0;
}

void o_completeness(const o_db *S, const o_db *A, const o_kmodel *m,
                    double total64[64], double undr64[64], double *total, double *undr) {
  /* The reference walks the 64 files independently; a k-mer's file is its top
   * numFilesBits bits, so one global merge assigns each k-mer to the same file. */
  for (int i = 0; i < 64; i++) total64[i] = undr64[i] = 0.0;
  uint32_t shift = 2 * S->k - 6;
  uint64_t si = 0, ai = 0;
  while (si < S->n) {
    int c = (ai < A->n) ? kcmp(S->hi[si], S->lo[si], A->hi[ai], A->lo[ai]) : -1;
    if (c > 0) { ai++; continue; }
    double readK, asmK, prob;
    o_getK(m, S->val[si], c == 0 ? A->val[ai] : 0, &readK, &asmK, &prob);
    o_u128 km = ((o_u128)S->hi[si] << 64) | S->lo[si];
    uint32_t ff = (uint32_t)(km >> shift) & 63;
    total64[ff] += readK;
    if (readK > asmK) undr64[ff] += readK - asmK;
    si++; if (c == 0) ai++;
  }
  *total = *undr = 0;
  for (int i = 0; i < 64; i++) { *total += total64[i]; *undr += undr64[i]; }
}

/* ======================================================================= */
/*  variant path scoring: varMer.C:48-145                                    */
/* ======================================================================= */

uint32_t o_score_path(const o_lookup *R, const o_lookup *A, const o_kmodel *m,
                      const char *path, uint64_t L, int filter_mode,
                      const uint32_t *alt_idx, const uint32_t *alt_len, const int *alt_gt, uint32_t nalt,
                      double *kstr, double *dk, uint64_t *nwin, double *prob_io) {
  uint32_t k = R->k, numM = 0, idx = 0;
  double prob = *prob_io;
  o_iter it; it_init(&it, k);
  for (uint64_t i = 0; i < L; i++) {
    int ok = it_push(&it, path[i]);
    if (i + 1 < k) continue;
    double readK = 0, asmK = 0;
    if (ok) {
      uint32_t vr, va;
      probe_pair(R, A, &it, &vr, &va);
      o_getK(m, vr, va, &readK, &asmK, &prob);
    }
    if (readK == 0) numM++;
    if (filter_mode) { idx++; continue; }
    double oD = fabs(readK - asmK) * prob;
    for (uint32_t j = 0; j < nalt; j++) {
      /* unsigned arithmetic on purpose: idxPath + 1 - k wraps when idxPath < k-1 (Q5) */
      if (alt_gt[j] > 0 && (uint32_t)(alt_idx[j] + 1 - k) <= idx && idx < (uint32_t)(alt_idx[j] + alt_len[j] + k)) {
        asmK++;
        break;
      }
    }
    double ks;
    if      (readK == 0)   ks = -1;
    else if (readK > asmK) ks = readK / asmK - 1;
    else                   ks = asmK / readK - 1;
    double nD = fabs(readK - asmK) * prob;
    kstr[idx] = ks; dk[idx] = oD - nD;
    idx++;
  }
  *nwin = idx;
  *prob_io = prob;
  return numM;
}
