// Joint (read, assembly) table: ONE probe per window instead of one per table.
//
// Measured on B200 (profiles/r01_gather_microbench.txt, r01_ncu_scan_v1_summary.txt): a random
// 32-byte sector miss moves a whole 128-byte line from HBM, and the chip sustains ~36.5e9 such
// misses per second whatever the load instruction.  So the unit of cost is the LINE, and the
// layout below spends exactly one line per window:
//
//   line   = 128 B = 4 sectors of 32 B; line index = h / D, tag = h % D (same bijective mix as
//            the per-database tables, so a range of h is a contiguous range of lines)
//   slot   = tag << (rbits + abits) | v_read << abits | v_asm   (32- or 64-bit; 0 = empty)
//   home   = sector (tag & 3); an insert that finds it full spills to the next sector of the
//            SAME line (already in L2 after the first miss), wrapping; a full line spills to a
//            small open-addressing side table with full keys and values
//   values that do not fit their field saturate and keep the excess in a second side table
//
// A lookup reads the home sector (one LDG.256) and stops at the first sector that is not full.
#pragma once

#include "table.cuh"

struct OaTable2 {          // open addressing, 64-bit keys (EMPTY = ~0), two 32-bit values
  unsigned long long *keys;
  uint32_t           *v0, *v1;
  uint32_t            mask;
  uint32_t           *special;   // [2]: values for key == ~0 (k == 32 only)
};

struct JointStats { unsigned long long n_home, n_spill, n_stash, n_big, errors; };

struct JointDev {
  void    *lines;
  uint64_t nlines, D, seed, kmask;
  // D = fs_d << fs_t with fs_d in {1, 3, 5, 9, 15}: h / D is a shift and an exact 32-bit
  // multiply-high division by the small odd factor (fs_m, fs_s), no 64-bit divide in the probe path
  uint32_t fs_t, fs_d, fs_m, fs_s;
  uint32_t k, w64;
  uint32_t rbits, abits;       // value field widths
  uint32_t rmax, amax;         // saturated field values (field masks)
  OaTable2 stash, big;
  JointStats *stats;
};

// h -> (line, tag) with line = h / D, tag = h % D for D = d << t (h >> t always fits 32 bits)
__device__ __forceinline__ void joint_split(uint64_t h, uint32_t t, uint32_t d, uint32_t m, uint32_t s,
                                            uint64_t &line, uint64_t &tag) {
  const uint32_t x = (uint32_t)(h >> t);
  const uint32_t q = d == 1u ? x : (__umulhi(x, m) >> s);   // x / d, exact for every 32-bit x
  const uint32_t rem = x - q * d;
  line = q;
  tag = ((uint64_t)rem << t) | (h & ((1ull << t) - 1ull));
}

__device__ __forceinline__ void joint_locate(const JointDev &J, uint64_t c, uint64_t &line, uint64_t &tag) {
  joint_split(mfn_mix(c, J.seed, J.kmask, J.k), J.fs_t, J.fs_d, J.fs_m, J.fs_s, line, tag);
}

__device__ __forceinline__ void joint_load_sector(const JointDev &J, uint64_t line, uint32_t sector, Bucket &b) {
  const char *p = (const char *)J.lines + line * 128 + sector * 32;
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(b.w[0]), "=r"(b.w[1]), "=r"(b.w[2]), "=r"(b.w[3]),
                 "=r"(b.w[4]), "=r"(b.w[5]), "=r"(b.w[6]), "=r"(b.w[7])
               : "l"(p));
}

// Search one sector.  found: slot payload (both value fields) in 'pay'.
__device__ __forceinline__ bool joint_match(const JointDev &J, const Bucket &b, uint64_t tag, uint64_t &pay, bool &full) {
  const uint32_t vb = J.rbits + J.abits;
  bool hit = false, f = true;
  if (J.w64) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      uint64_t s = ((uint64_t)b.w[2 * i + 1] << 32) | b.w[2 * i];
      if (s != 0 && (s >> vb) == tag) { hit = true; pay = s; }
      f = f && (s != 0);
    }
  } else {
    const uint32_t t32 = (uint32_t)tag;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      uint32_t s = b.w[i];
      if (s != 0 && (s >> vb) == t32) { hit = true; pay = s; }
      f = f && (s != 0);
    }
  }
  full = f;
  return hit;
}

// Fast-path search of the home sector: hit and payload only (everything else goes to joint_full).
// A stored slot always has a non-zero value field, so  slot == want | values  <=>  1 <= slot ^ want < 2^vb.
__device__ __forceinline__ bool joint_match_fast(const JointDev &J, const Bucket &b, uint64_t tag, uint64_t &pay) {
  const uint32_t vb = J.rbits + J.abits;
  bool hit = false;
  if (J.w64) {
    const uint64_t want = tag << vb, lim = (1ull << vb) - 1;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const uint64_t d = (((uint64_t)b.w[2 * i + 1] << 32) | b.w[2 * i]) ^ want;
      if (d - 1ull < lim) { hit = true; pay = d; }
    }
  } else {
    const uint32_t want = (uint32_t)tag << vb, lim = (1u << vb) - 1u;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const uint32_t d = b.w[i] ^ want;
      if (d - 1u < lim) { hit = true; pay = d; }
    }
  }
  return hit;
}

__device__ __forceinline__ bool oa2_get(const OaTable2 &t, uint64_t key, uint32_t &a, uint32_t &b) {
  if (key == MFN_EMPTY64) { a = t.special[0]; b = t.special[1]; return (a | b) != 0; }
  uint32_t i = mfn_oa_hash(key) & t.mask;
  for (uint32_t n = 0; n <= t.mask; n++) {          // bounded even if the table is full
    uint64_t cur = t.keys[i];
    if (cur == key) { a = t.v0[i]; b = t.v1[i]; return true; }
    if (cur == MFN_EMPTY64) break;
    i = (i + 1) & t.mask;
  }
  a = 0; b = 0;
  return false;
}

// Everything past the home sector: spill sectors, side table, saturated values.
static __device__ __noinline__ void joint_slow(const JointDev &J, uint64_t c, uint64_t line, uint64_t tag,
                                               bool hit, bool full, uint64_t pay, uint32_t &vr, uint32_t &va) {
  const uint32_t home = (uint32_t)tag & 3u;
  for (uint32_t j = 1; !hit && full && j < 4; j++) {
    Bucket b;
    joint_load_sector(J, line, (home + j) & 3u, b);
    hit = joint_match(J, b, tag, pay, full);
  }
  if (hit) {
    va = (uint32_t)pay & J.amax;
    vr = (uint32_t)(pay >> J.abits) & J.rmax;
    if (vr == J.rmax || va == J.amax) {
      uint32_t er, ea;
      oa2_get(J.big, c, er, ea);
      if (vr == J.rmax) vr += er;
      if (va == J.amax) va += ea;
    }
  } else if (full) {
    oa2_get(J.stash, c, vr, va);
  } else {
    vr = 0; va = 0;
  }
}

// Finish a lookup whose home sector is already in registers.
__device__ __forceinline__ void joint_resolve(const JointDev &J, uint64_t c, uint64_t line, uint64_t tag,
                                              const Bucket &b, uint32_t &vr, uint32_t &va) {
  uint64_t pay = 0;
  bool full;
  bool hit = joint_match(J, b, tag, pay, full);
  va = (uint32_t)pay & J.amax;
  vr = (uint32_t)(pay >> J.abits) & J.rmax;
  if ((!hit && full) || (hit && (vr == J.rmax || va == J.amax))) joint_slow(J, c, line, tag, hit, full, pay, vr, va);
  else if (!hit) { vr = 0; va = 0; }
}

// Complete lookup from scratch (used for the few lookups the home sector alone cannot answer).
static __device__ __noinline__ void joint_full(const JointDev &J, uint64_t c, uint64_t line, uint64_t tag,
                                               uint32_t &vr, uint32_t &va) {
  Bucket b;
  joint_load_sector(J, line, (uint32_t)tag & 3u, b);
  uint64_t pay = 0;
  bool full;
  bool hit = joint_match(J, b, tag, pay, full);
  joint_slow(J, c, line, tag, hit, full, pay, vr, va);
}

__device__ __forceinline__ void joint_value(const JointDev &J, uint64_t c, uint32_t &vr, uint32_t &va) {
  uint64_t line, tag;
  joint_locate(J, c, line, tag);
  Bucket b;
  joint_load_sector(J, line, (uint32_t)tag & 3u, b);
  joint_resolve(J, c, line, tag, b, vr, va);
}
