// Device-side table primitives: bijective k-mer mix, bucket probe, find-or-insert.
// See common.h for the layout.  Everything here is integer arithmetic.
#pragma once

#include "common.h"

#define MFN_EMPTY64 0xffffffffffffffffull

// --- bijection on [0, 2^2k): xor-seed, xorshift and odd multiplies are all invertible -----
__host__ __device__ __forceinline__ uint64_t mfn_mix(uint64_t c, uint64_t seed, uint64_t kmask, uint32_t k) {
  uint64_t x = (c ^ seed) & kmask;
  x ^= x >> k;
  x = (x * 0xff51afd7ed558ccdull) & kmask;
  x ^= x >> k;
  x = (x * 0xc4ceb9fe1a85ec53ull) & kmask;
  x ^= x >> k;
  return x;
}

// h -> (bucket, tag) with bucket = h / D, tag = h % D.  magic = floor(2^64 / D) (or ~0 for D == 1).
__device__ __forceinline__ void mfn_split(uint64_t h, uint64_t D, uint64_t magic, uint64_t &bucket, uint64_t &tag) {
  uint64_t q = __umul64hi(h, magic);
  uint64_t r = h - q * D;
  if (r >= D) { q++; r -= D; }
  bucket = q; tag = r;
}

// reverse the order of the 32 two-bit groups of x
__host__ __device__ __forceinline__ uint64_t mfn_rev2(uint64_t x) {
#ifdef __CUDA_ARCH__
  // two BREVs reverse all 64 bits; then swap the two bits inside every group back
  uint64_t y = ((uint64_t)__brev((uint32_t)x) << 32) | __brev((uint32_t)(x >> 32));
  return ((y >> 1) & 0x5555555555555555ull) | ((y & 0x5555555555555555ull) << 1);
#else
  x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
  x = ((x >> 4) & 0x0f0f0f0f0f0f0f0full) | ((x & 0x0f0f0f0f0f0f0f0full) << 4);
  x = ((x >> 8) & 0x00ff00ff00ff00ffull) | ((x & 0x00ff00ff00ff00ffull) << 8);
  x = ((x >> 16) & 0x0000ffff0000ffffull) | ((x & 0x0000ffff0000ffffull) << 16);
  return (x >> 32) | (x << 32);
#endif
}

__host__ __device__ __forceinline__ uint64_t mfn_revcomp(uint64_t f, uint32_t k) {
  // complement every base (A=00 <-> T=10, C=01 <-> G=11), reverse the 2-bit groups, realign
  return mfn_rev2(f ^ 0xaaaaaaaaaaaaaaaaull) >> (64 - 2 * k);
}

// 32-byte bucket load: one LDG.256, read-only path, no L1 allocation (no reuse).
struct Bucket { uint32_t w[8]; };

__device__ __forceinline__ void mfn_load_bucket(const void *slots, uint64_t bucket, Bucket &b) {
  const char *p = (const char *)slots + bucket * 32;
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(b.w[0]), "=r"(b.w[1]), "=r"(b.w[2]), "=r"(b.w[3]),
                 "=r"(b.w[4]), "=r"(b.w[5]), "=r"(b.w[6]), "=r"(b.w[7])
               : "l"(p));
}

// Search a loaded bucket.  Returns the stored value (0 = not here); full = all slots occupied.
__device__ __forceinline__ uint32_t mfn_match32(const Bucket &b, uint64_t tag, uint32_t vbits, uint32_t vmax, bool &full) {
  uint32_t want = (uint32_t)tag << vbits;
  uint32_t val = 0;
  bool f = true;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint32_t d = b.w[i] ^ want;            // equal tag and non-zero value  <=>  1 <= d <= vmax
    if (d - 1u < vmax) val = d;
    f = f && (b.w[i] != 0);
  }
  full = f;
  return val;
}

__device__ __forceinline__ uint32_t mfn_match64(const Bucket &b, uint64_t tag, uint32_t vbits, uint32_t vmax, bool &full) {
  uint64_t want = tag << vbits;
  uint32_t val = 0;
  bool f = true;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint64_t s = ((uint64_t)b.w[2 * i + 1] << 32) | b.w[2 * i];
    uint64_t d = s ^ want;
    if (d - 1ull < (uint64_t)vmax) val = (uint32_t)d;
    f = f && (s != 0);
  }
  full = f;
  return val;
}

__device__ __forceinline__ uint32_t mfn_oa_hash(uint64_t key) {
  key ^= key >> 33; key *= 0xff51afd7ed558ccdull; key ^= key >> 29;
  return (uint32_t)key;
}

__device__ __forceinline__ uint32_t mfn_oa_get(const OaTable &t, uint64_t key) {
  if (key == MFN_EMPTY64) return *t.special;
  uint32_t i = mfn_oa_hash(key) & t.mask;
  for (uint32_t n = 0; n <= t.mask; n++) {          // bounded even if the table is full
    uint64_t cur = t.keys[i];
    if (cur == key) return t.vals[i];
    if (cur == MFN_EMPTY64) return 0;
    i = (i + 1) & t.mask;
  }
  return 0;
}

// Levels 2 and 3 and the saturated-value excess: rare, kept out of line.
static __device__ __noinline__ uint32_t mfn_probe_slow(const DevTable &t, uint64_t c, uint32_t val, bool full) {
  if (val == 0 && full) {
    uint64_t bucket, tag;
    mfn_split(mfn_mix(c, t.lv[1].seed, t.kmask, t.k), t.lv[1].D, t.lv[1].magic, bucket, tag);
    Bucket b;
    mfn_load_bucket(t.lv[1].slots, bucket, b);
    bool full2;
    val = t.w64 ? mfn_match64(b, tag, t.vbits, t.vmax, full2) : mfn_match32(b, tag, t.vbits, t.vmax, full2);
    if (val == 0 && full2) return mfn_oa_get(t.stash, c);     // stash keeps full values
  }
  if (val == t.vmax) val += mfn_oa_get(t.big, c);
  return val;
}

// Address of the level-1 bucket of k-mer c and its tag.
__device__ __forceinline__ void mfn_locate(const DevTable &t, uint64_t c, uint64_t &bucket, uint64_t &tag) {
  mfn_split(mfn_mix(c, t.lv[0].seed, t.kmask, t.k), t.lv[0].D, t.lv[0].magic, bucket, tag);
}

// Finish a probe whose level-1 bucket is already in registers.
__device__ __forceinline__ uint32_t mfn_resolve(const DevTable &t, uint64_t c, const Bucket &b, uint64_t tag) {
  bool full;
  uint32_t val = t.w64 ? mfn_match64(b, tag, t.vbits, t.vmax, full) : mfn_match32(b, tag, t.vbits, t.vmax, full);
  if ((val == 0 && full) || val == t.vmax) val = mfn_probe_slow(t, c, val, full);
  return val;
}

// Complete lookup of one k-mer exactly as given: merylExactLookup::value() (kmers-lookup.H:334-380).
__device__ __forceinline__ uint32_t mfn_value(const DevTable &t, uint64_t c) {
  uint64_t bucket, tag;
  mfn_locate(t, c, bucket, tag);
  Bucket b;
  mfn_load_bucket(t.lv[0].slots, bucket, b);
  return mfn_resolve(t, c, b, tag);
}

// ---------------------------------------------------------------------------
// inverse of mfn_mix (for export / iteration over stored entries)
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint64_t mfn_unmix(uint64_t h, uint64_t seed, uint64_t kmask, uint32_t k) {
  // xorshift by half the width is an involution; the multipliers are undone by their inverses mod 2^64
  uint64_t x = h;
  x ^= x >> k;
  x = (x * 0x9cb4b2f8129337dbull) & kmask;     // inverse of 0xc4ceb9fe1a85ec53
  x ^= x >> k;
  x = (x * 0x4f74430c22a54005ull) & kmask;     // inverse of 0xff51afd7ed558ccd
  x ^= x >> k;
  return (x ^ seed) & kmask;
}

static_assert(0xff51afd7ed558ccdull * 0x4f74430c22a54005ull == 1ull, "mix multiplier 1 inverse");
static_assert(0xc4ceb9fe1a85ec53ull * 0x9cb4b2f8129337dbull == 1ull, "mix multiplier 2 inverse");


// Walk every stored entry: level slots (k-mer rebuilt from bucket and tag), then the stash.
template <class F>
__device__ __forceinline__ void mfn_foreach_entry(const DevTable &t, F f) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const int S = t.w64 ? 4 : 8;
  for (int lvl = 0; lvl < 2; lvl++) {
    const uint64_t nslots = t.lv[lvl].buckets * S;
    for (uint64_t s = tid; s < nslots; s += stride) {
      uint64_t slot = t.w64 ? ((const unsigned long long *)t.lv[lvl].slots)[s]
                            : (uint64_t)((const unsigned int *)t.lv[lvl].slots)[s];
      if (slot == 0) continue;
      uint64_t tag = slot >> t.vbits;
      uint32_t val = (uint32_t)(slot & t.vmax);
      uint64_t h = (s / S) * t.lv[lvl].D + tag;
      uint64_t c = mfn_unmix(h, t.lv[lvl].seed, t.kmask, t.k);
      if (val == t.vmax) val += mfn_oa_get(t.big, c);
      f(c, val);
    }
  }
  for (uint64_t s = tid; s <= t.stash.mask; s += stride) {
    uint64_t key = t.stash.keys[s];
    if (key != MFN_EMPTY64) f(key, t.stash.vals[s]);
  }
  if (tid == 0 && *t.stash.special) f(MFN_EMPTY64, *t.stash.special);
}

