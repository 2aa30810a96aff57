// Device-side insertion into a joint table (shared by joint.cu and joint_keys.cu).
#pragma once
#include "joint.cuh"
#include "table_host.h"

#include <functional>
#include <type_traits>

struct JointCounts { uint32_t home, spill, stash, big; };

__device__ __forceinline__ void oa2_add(const OaTable2 &t, uint64_t key, int which, uint32_t v, JointStats *st) {
  if (key == MFN_EMPTY64) { atomicAdd(&t.special[which], v); return; }
  uint32_t i = mfn_oa_hash(key) & t.mask;
  for (uint32_t n = 0; n <= t.mask; n++) {
    unsigned long long cur = t.keys[i];
    if (cur == MFN_EMPTY64) {
      cur = atomicCAS(&t.keys[i], MFN_EMPTY64, (unsigned long long)key);
      if (cur == MFN_EMPTY64) cur = key;
    }
    if (cur == key) { atomicAdd(which ? &t.v1[i] : &t.v0[i], v); return; }
    i = (i + 1) & t.mask;
  }
  atomicAdd(&st->errors, 1ull);
}

// Set field 'which' (0 = read value, 1 = assembly value) of k-mer c to v, creating the slot if needed.
template <bool W64>
__device__ __forceinline__ void joint_insert(const JointDev &J, uint64_t c, int which, uint32_t v, JointCounts &cnt) {
  typedef typename std::conditional<W64, unsigned long long, unsigned int>::type slot_t;
  constexpr int S = W64 ? 4 : 8;
  const uint32_t vb = J.rbits + J.abits;
  const uint32_t fmax = which ? J.amax : J.rmax;
  const uint32_t sat = v < fmax ? v : fmax;
  const slot_t field = which ? (slot_t)sat : ((slot_t)sat << J.abits);
  uint64_t line, tag;
  joint_locate(J, c, line, tag);
  const slot_t want = (slot_t)tag << vb;
  const uint32_t home = (uint32_t)tag & 3u;
#pragma unroll 1
  for (uint32_t j = 0; j < 4; j++) {
    slot_t *slots = (slot_t *)((char *)J.lines + line * 128 + ((home + j) & 3u) * 32);
#pragma unroll 1
    for (int i = 0; i < S; i++) {
      slot_t cur = *(volatile slot_t *)(slots + i);
      if (cur == 0) {
        slot_t old = atomicCAS(slots + i, (slot_t)0, (slot_t)(want | field));
        if (old == 0) {
          if (j == 0) cnt.home++; else cnt.spill++;
          if (v > fmax) { oa2_add(J.big, c, which, v - fmax, J.stats); cnt.big++; }
          return;
        }
        cur = old;
      }
      if ((cur >> vb) == (slot_t)tag) {            // the k-mer has a slot already: fill in our field
        atomicOr(slots + i, field);
        if (v > fmax) { oa2_add(J.big, c, which, v - fmax, J.stats); cnt.big++; }
        return;
      }
    }
  }
  uint32_t a, b;
  if (!oa2_get(J.stash, c, a, b)) cnt.stash++;
  oa2_add(J.stash, c, which, v, J.stats);
}


// host pieces of the build shared between the translation units
int mfn_joint_alloc(mfn_ctx *ctx, const mfn_table *R, const mfn_table *A, uint64_t n_kmers, bool want_pow2,
                    const std::function<uint64_t(uint32_t, uint32_t)> &count_big, mfn_joint **out);
uint64_t mfn_joint_count_saturated(mfn_ctx *ctx, const mfn_table *R, const mfn_table *A, uint32_t rmax, uint32_t amax);
int mfn_joint_finish(mfn_ctx *ctx, mfn_joint *j, bool exact_count);
