// Device-side pieces of the k = 33..64 path (wide.cu): the sorted table's lookup and the window arithmetic on
// 128-bit k-mers held as two 64-bit halves.  Everything here is plain integer code marked __host__ __device__,
// so that tests/wide_emulation_harness.cu can run the very same functions on the CPU against the oracle.
#pragma once
#include "table.cuh"

struct DevWide {
  const unsigned long long *hi, *lo;   // n k-mers, ascending by (hi, lo)
  const uint32_t *val;                 // stored values (-min offset applied)
  const unsigned long long *index;     // [2^ibits + 1]: first entry whose top ibits bits are >= p
  uint64_t n;
  uint32_t ibits, k;
};

// top ibits bits of a 2k-bit k-mer (2k > 64 >= ibits)
__host__ __device__ __forceinline__ uint64_t wide_prefix(uint32_t k, uint32_t ibits, uint64_t hi, uint64_t lo) {
  const uint32_t s = 2 * k - ibits;                       // 40 <= s < 128
  return s >= 64 ? (hi >> (s - 64)) : ((hi << (64 - s)) | (lo >> s));
}

// merylExactLookup::value (kmers-lookup.H:334-380): the stored value of the k-mer exactly as given, 0 if absent
__host__ __device__ __forceinline__ uint32_t wide_value(const DevWide &T, uint64_t hi, uint64_t lo) {
  if (T.n == 0) return 0;
  const uint64_t p = wide_prefix(T.k, T.ibits, hi, lo);
  uint64_t a = T.index[p], b = T.index[p + 1];
  while (a < b) {
    const uint64_t m = a + ((b - a) >> 1);
    const uint64_t mh = T.hi[m], ml = T.lo[m];
    if (mh == hi && ml == lo) return T.val[m];
    if (mh < hi || (mh == hi && ml < lo)) a = m + 1; else b = m;
  }
  return 0;
}

// entry i's share of the index (i == n closes it): index[q] = i for every prefix q after entry i - 1's up to entry i's
__host__ __device__ __forceinline__ void wide_index_fill(const DevWide &T, unsigned long long *index, uint64_t i) {
  const uint64_t np = 1ull << T.ibits;
  const uint64_t cur = i < T.n ? wide_prefix(T.k, T.ibits, T.hi[i], T.lo[i]) : np;
  const uint64_t first = i == 0 ? 0 : wide_prefix(T.k, T.ibits, T.hi[i - 1], T.lo[i - 1]) + 1;
  for (uint64_t q = first; q <= cur; q++) index[q] = i;
}

__host__ __device__ __forceinline__ uint32_t wide_index_bits(uint64_t n) {
  uint32_t ib = 1;
  while (ib < 26 && (1ull << ib) < n + 1) ib++;
  return ib;
}

// masks for k = 33..64: the high half of a k-mer, and k consecutive valid bases
__host__ __device__ __forceinline__ uint64_t wide_hmask(uint32_t k) { return k == 64 ? ~0ull : ((1ull << (2 * k - 64)) - 1); }
__host__ __device__ __forceinline__ uint64_t wide_kbits(uint32_t k) { return k == 64 ? ~0ull : ((1ull << k) - 1); }

// A thread's 8 windows start at positions j0 .. j0 + 7 (j0 a multiple of 8) of an array of packed words (2 bits per
// base, base g at bits 2 (g % 32) of word g / 32) and validity words (1 bit per base): they need bases
// j0 .. j0 + 70 = a 192-bit slice (s0, s1, s2) starting at base j0, and validity bits j0 .. j0 + 70 = (ok0, ok1).
struct WideSlice { uint64_t s0, s1, s2, ok0, ok1; };

__host__ __device__ __forceinline__ WideSlice wide_slice(const unsigned long long *seq_words, const unsigned long long *ok_words, uint32_t j0) {
  WideSlice S;
  const uint32_t w = j0 >> 5, sh = (j0 & 31) * 2;
  const uint64_t a = seq_words[w], b = seq_words[w + 1], c = seq_words[w + 2];
  S.s0 = sh ? ((a >> sh) | (b << (64 - sh))) : a;
  S.s1 = sh ? ((b >> sh) | (c << (64 - sh))) : b;
  S.s2 = c >> sh;
  const uint32_t wo = j0 >> 6, so = j0 & 63;
  const uint64_t d = ok_words[wo], e = ok_words[wo + 1];
  S.ok0 = so ? ((d >> so) | (e << (64 - so))) : d;
  S.ok1 = e >> so;
  return S;
}

// window uu (0..7) of a slice: its 2k bits (first base in the low bits) and whether all its k bases are valid
__host__ __device__ __forceinline__ bool wide_window(const WideSlice &S, int uu, uint64_t hmask, uint64_t kb, uint64_t &xh, uint64_t &xl) {
  const uint32_t t2 = 2 * uu;
  xl = uu ? ((S.s0 >> t2) | (S.s1 << (64 - t2))) : S.s0;
  xh = (uu ? ((S.s1 >> t2) | (S.s2 << (64 - t2))) : S.s1) & hmask;
  const uint64_t vb = uu ? ((S.ok0 >> uu) | (S.ok1 << (64 - uu))) : S.ok0;
  return (vb & kb) == kb;
}

// The forward k-mer (first base in the top bits, kmers-tiny.H:82) and the reverse complement of the window whose
// bases sit in (xh:xl), first base in the low bits, already cut to 2k bits.
__host__ __device__ __forceinline__ void wide_kmers(uint32_t k, uint64_t hmask, uint64_t xh, uint64_t xl, uint64_t &fh, uint64_t &fl,
                                                    uint64_t &rh, uint64_t &rl) {
  rl = xl ^ 0xaaaaaaaaaaaaaaaaull;
  rh = (xh ^ 0xaaaaaaaaaaaaaaaaull) & hmask;
  const uint64_t qh = mfn_rev2(xl), ql = mfn_rev2(xh);        // the 64 two-bit groups of (xh:xl) in reverse order
  const uint32_t t = 128 - 2 * k;                             // 0 <= t <= 62
  fl = t ? ((ql >> t) | (qh << (64 - t))) : ql;
  fh = qh >> t;
}

// value(f) + value(r) of one window in both tables: merfinGlobal::getK's lookups (merfin-globals.C:100-111), uint32 adds
__host__ __device__ __forceinline__ void wide_pair_of(const DevWide &R, const DevWide &A, uint32_t k, uint64_t hmask, uint64_t xh, uint64_t xl,
                                                      uint32_t &vr, uint32_t &va) {
  uint64_t fh, fl, rh, rl;
  wide_kmers(k, hmask, xh, xl, fh, fl, rh, rl);
  vr = wide_value(R, fh, fl) + wide_value(R, rh, rl);
  va = wide_value(A, fh, fl) + wide_value(A, rh, rl);
}
