// Host-side packing of ASCII bases into the device layout of mfn_seq (seq.cu): 2 bits per base
// ((ascii >> 1) & 3, kmers-tiny.H:69-83) + 1 validity bit per base (ACGTacgt, kmers-iterator.H:97-102),
// in GLOBAL coordinates (one separator slot after every record).
//
// Why on the host at all: the host-buffer entry points are bound by the PCIe copy of 1 byte per
// base.  A packed piece is 0.375 bytes per base, so pieces packed by idle host cores cross the
// link 2.7x cheaper; mfn_hist_host runs both lanes side by side (api.cu).
//
// Plain C++ (compiled by g++, not nvcc); the inner kernel is chosen at run time: AVX2 (two 32-base halves per 64-base
// block) where the CPU has it, a scalar loop otherwise, and a 64-base AVX-512BW kernel on request.  Bit-identical to k_pack by construction; the GPU
// parity test packs the same records both ways and compares the arrays.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <immintrin.h>

// 32 bases at p -> (2-bit codes, base i at bits 2i; validity mask, base i at bit i); invalid bases have code 0
static inline void pack32_scalar(const unsigned char *p, uint64_t *codes, uint32_t *okmask) {
  uint64_t c = 0;
  uint32_t m = 0;
  for (int i = 0; i < 32; i++) {
    const unsigned u = p[i] | 0x20u;
    if (u == 'a' || u == 'c' || u == 'g' || u == 't') {
      c |= (uint64_t)((p[i] >> 1) & 3u) << (2 * i);
      m |= 1u << i;
    }
  }
  *codes = c;
  *okmask = m;
}

__attribute__((target("avx2"))) static inline void pack32_avx2(const unsigned char *p, uint64_t *codes, uint32_t *okmask) {
  const __m256i v = _mm256_loadu_si256((const __m256i *)p);
  const __m256i u = _mm256_or_si256(v, _mm256_set1_epi8(0x20));
  // expected byte for each low nibble: 1 -> 'a', 3 -> 'c', 7 -> 'g', 4 -> 't'; 0 for every other nibble, which no byte
  // with bit 0x20 set can equal (0xff as the "never" value made the bytes 0xdf and 0xff valid bases)
  const __m256i lut = _mm256_setr_epi8(0, 'a', 0, 'c', 't', 0, 0, 'g', 0, 0, 0, 0, 0, 0, 0, 0,
                                       0, 'a', 0, 'c', 't', 0, 0, 'g', 0, 0, 0, 0, 0, 0, 0, 0);
  const __m256i want = _mm256_shuffle_epi8(lut, _mm256_and_si256(u, _mm256_set1_epi8(0x0f)));
  const __m256i ok = _mm256_cmpeq_epi8(want, u);
  *okmask = (uint32_t)_mm256_movemask_epi8(ok);
  // codes 0..3 per byte, zero where invalid
  __m256i c = _mm256_and_si256(_mm256_srli_epi16(v, 1), _mm256_set1_epi8(3));
  c = _mm256_and_si256(c, ok);
  // pairs of bytes -> 4 bits (c0 + 4 c1), pairs of those -> 8 bits (x0 + 16 x1): one byte per 32-bit lane
  const __m256i x = _mm256_maddubs_epi16(c, _mm256_set1_epi16(0x0401));
  const __m256i y = _mm256_madd_epi16(x, _mm256_set1_epi32(0x00100001));
  // gather the 8 low bytes of the 8 lanes
  const __m256i z = _mm256_shuffle_epi8(y, _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                                            0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1));
  const uint32_t lo = (uint32_t)_mm256_extract_epi32(z, 0), hi = (uint32_t)_mm256_extract_epi32(z, 4);
  *codes = ((uint64_t)hi << 32) | lo;
}

// 64 bases at p with AVX-512BW: the validity comparison yields the 64-bit mask directly, and the down-conversion of the
// sixteen 32-bit lanes (each holding four bases' codes in its low byte) yields the two code words
__attribute__((target("avx512f,avx512bw"))) static inline void pack64_avx512(const unsigned char *p, uint64_t *c0, uint64_t *c1, uint64_t *okmask) {
  const __m512i v = _mm512_loadu_si512((const void *)p);
  const __m512i u = _mm512_or_si512(v, _mm512_set1_epi8(0x20));
  const __m512i lut = _mm512_broadcast_i32x4(_mm_setr_epi8(0, 'a', 0, 'c', 't', 0, 0, 'g', 0, 0, 0, 0, 0, 0, 0, 0));
  const __m512i want = _mm512_shuffle_epi8(lut, _mm512_and_si512(u, _mm512_set1_epi8(0x0f)));
  const __mmask64 ok = _mm512_cmpeq_epi8_mask(want, u);
  const __m512i c = _mm512_maskz_mov_epi8(ok, _mm512_and_si512(_mm512_srli_epi16(v, 1), _mm512_set1_epi8(3)));
  const __m512i x = _mm512_maddubs_epi16(c, _mm512_set1_epi16(0x0401));
  const __m512i y = _mm512_madd_epi16(x, _mm512_set1_epi32(0x00100001));
  const __m128i z = _mm512_cvtepi32_epi8(y);                    // byte j = bases 4j .. 4j + 3
  *c0 = (uint64_t)_mm_cvtsi128_si64(z);
  *c1 = (uint64_t)_mm_extract_epi64(z, 1);
  *okmask = (uint64_t)ok;
}

// OR `n` (<= 32) bases starting at global position g into the word arrays (which hold zeros there)
static inline void deposit(uint64_t *seq2, uint64_t *ok, uint64_t g, uint64_t codes, uint32_t okmask) {
  const uint64_t w = g >> 5, sh = (g & 31) * 2;
  seq2[w] |= codes << sh;
  if (sh) { const uint64_t hi = codes >> (64 - sh); if (hi) seq2[w + 1] |= hi; }     // never touch a word we put nothing in:
  const uint64_t wo = g >> 6, so = g & 63;                                          // it may belong to another thread
  ok[wo] |= (uint64_t)okmask << so;
  if (so > 32) { const uint64_t hi = (uint64_t)okmask >> (64 - so); if (hi) ok[wo + 1] |= hi; }
}

// Zero the words of global positions [z0, z1) (multiples of 64).
static inline void zero_words(uint64_t *s2, uint64_t *so, uint64_t z0, uint64_t z1) {
  if (z1 <= z0) return;
  memset(s2 + (z0 >> 5), 0, (z1 - z0) / 32 * 8);
  memset(so + (z0 >> 6), 0, (z1 - z0) / 64 * 8);
}

#define MFN_PACK_PREFETCH 2048        // bytes ahead of the read pointer (one core streams ~1.5x faster with it)

// One record segment [a, b) of global positions whose first source byte is p.  The output words
// are NOT zeroed beforehand: whole 64-position blocks inside the segment are stored outright, and
// `z` (a multiple of 64) says up to where everything else has been zeroed so far; blocks that only
// get part of their bits from this segment are zeroed here before the bits are OR-ed in.
template <class PACK32, class PACK64>
static inline __attribute__((always_inline)) void pack_segment(PACK32 pack32, PACK64 pack64, const unsigned char *p, uint64_t a, uint64_t b,
                                                               uint64_t *s2, uint64_t *so, uint64_t &z) {
  uint64_t g = a;
  auto dep = [&](uint64_t n) {                       // n <= 32 bases at g, OR-ed into zeroed words
    unsigned char tmp[32];
    memset(tmp, 0, sizeof tmp);
    memcpy(tmp, p, n);
    uint64_t codes;
    uint32_t okm;
    pack32(tmp, &codes, &okm);
    deposit(s2, so, g, codes, okm);
    g += n; p += n;
  };
  const uint64_t fa = (a + 63) & ~63ull, fb = b & ~63ull;      // whole blocks: [fa, fb)
  if (fa < fb) {
    zero_words(s2, so, z, fa);                       // the block the segment starts in (and any gap before it)
    while (g < fa) dep(std::min<uint64_t>(32, fa - g));
    {
      // (locals: g and p are captured by reference above and would live in memory otherwise)
      const unsigned char *q = p;
      uint64_t *o2 = s2 + (fa >> 5), *oo = so + (fa >> 6);
      const uint64_t nblk = (fb - fa) >> 6;
      for (uint64_t i = 0; i < nblk; i++, q += 64) {
        _mm_prefetch((const char *)q + MFN_PACK_PREFETCH, _MM_HINT_T0);
        uint64_t c0, c1, m;
        pack64(q, &c0, &c1, &m);
        o2[2 * i] = c0;
        o2[2 * i + 1] = c1;
        oo[i] = m;
      }
      p = q; g = fb;
    }
    z = fb;
    if (g < b) {                                     // the block the segment ends in
      zero_words(s2, so, fb, fb + 64);
      z = fb + 64;
      while (g < b) dep(std::min<uint64_t>(32, b - g));
    }
  } else {
    // a short segment: no whole block; make sure the blocks it touches are zeroed, then OR it in
    const uint64_t zend = (b + 63) & ~63ull;
    if (zend > z) { zero_words(s2, so, z, zend); z = zend; }
    while (g < b) dep(std::min<uint64_t>(32, b - g));
  }
}
static inline void pack64_scalar(const unsigned char *q, uint64_t *c0, uint64_t *c1, uint64_t *m) {
  uint32_t m0, m1;
  pack32_scalar(q, c0, &m0);
  pack32_scalar(q + 32, c1, &m1);
  *m = ((uint64_t)m1 << 32) | m0;
}
__attribute__((target("avx2"))) static inline void pack64_avx2(const unsigned char *q, uint64_t *c0, uint64_t *c1, uint64_t *m) {
  uint32_t m0, m1;
  pack32_avx2(q, c0, &m0);
  pack32_avx2(q + 32, c1, &m1);
  *m = ((uint64_t)m1 << 32) | m0;
}
static void pack_segment_scalar(const unsigned char *p, uint64_t a, uint64_t b, uint64_t *s2, uint64_t *so, uint64_t &z) {
  pack_segment(pack32_scalar, pack64_scalar, p, a, b, s2, so, z);
}
__attribute__((target("avx2"))) static void pack_segment_avx2(const unsigned char *p, uint64_t a, uint64_t b, uint64_t *s2, uint64_t *so, uint64_t &z) {
  pack_segment(pack32_avx2, pack64_avx2, p, a, b, s2, so, z);
}
__attribute__((target("avx512f,avx512bw,avx2"))) static void pack_segment_avx512(const unsigned char *p, uint64_t a, uint64_t b, uint64_t *s2, uint64_t *so, uint64_t &z) {
  pack_segment(pack32_avx2, pack64_avx512, p, a, b, s2, so, z);
}
typedef void (*segment_fn)(const unsigned char *, uint64_t, uint64_t, uint64_t *, uint64_t *, uint64_t &);
// The kernel: AVX2 where the CPU has it, else scalar.  MFN_PACK_ISA = scalar | avx2 | avx512 picks another one the CPU
// supports (tests compare all three; MFN_PACK_SCALAR=1 is the older spelling of "scalar").  The AVX-512 kernel is not the
// default: one core packs 5-10 % faster with it, but from a handful of threads on both are bound by memory bandwidth
// (8 threads: 41-49 GB/s of ASCII either way on the development host).
static int pack_level() {
  __builtin_cpu_init();
  const bool avx2 = __builtin_cpu_supports("avx2");
  const bool avx512 = avx2 && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw");
  int level = avx2 ? 1 : 0;
  const char *env = getenv("MFN_PACK_SCALAR");
  if (env && atoi(env)) level = 0;
  if (const char *isa = getenv("MFN_PACK_ISA")) {
    if (!strcmp(isa, "scalar")) level = 0;
    else if (!strcmp(isa, "avx512") && avx512) level = 2;
  }
  return level;
}
static segment_fn pick_segment() {
  const int level = pack_level();
  return level == 2 ? pack_segment_avx512 : level == 1 ? pack_segment_avx2 : pack_segment_scalar;
}

// Pack global positions [g_begin, g_end) (both multiples of 64) of a sequence set into
// seq2_out / ok_out, whose word 0 is global word g_begin/32 resp. g_begin/64.  rec_start[r] is the
// global coordinate of record r's first base (nseq + 1 entries, rec_start[r] = source offset + r);
// `bases` is the caller's concatenated ASCII.  Every word of the range is written exactly once where a
// record covers it whole and zeroed-then-OR-ed elsewhere (zeros where nothing is); no up-front memset.
extern "C" void mfn_pack_host_range(const char *bases, const uint64_t *rec_start, uint32_t nseq, uint64_t g_begin,
                                    uint64_t g_end, uint64_t *seq2_out, uint64_t *ok_out) {
  static const segment_fn segment = pick_segment();
  if (g_end <= g_begin) return;
  uint64_t *const s2 = seq2_out - g_begin / 32, *const so = ok_out - g_begin / 64;      // index by global word
  uint64_t z = g_begin;                              // everything below z is final (stored or zeroed)
  if (nseq == 0) { zero_words(s2, so, z, g_end); return; }
  // first record that ends after g_begin
  uint32_t r = (uint32_t)(std::upper_bound(rec_start, rec_start + nseq + 1, g_begin) - rec_start);
  r = r ? r - 1 : 0;
  const unsigned char *src = (const unsigned char *)bases;
  for (; r < nseq && rec_start[r] < g_end; r++) {
    // bases of record r occupy [rec_start[r], rec_start[r+1] - 1); the last slot is the separator
    const uint64_t a = std::max(rec_start[r], g_begin), b = std::min(rec_start[r + 1] - 1, g_end);
    if (a < b) segment(src + (a - r), a, b, s2, so, z);        // a - r: source byte of global position a
  }
  zero_words(s2, so, z, g_end);
}

// Which inner kernel this CPU gets (for logs and tests).
extern "C" const char *mfn_pack_host_isa(void) { const int level = pack_level(); return level == 2 ? "avx512" : level == 1 ? "avx2" : "scalar"; }
