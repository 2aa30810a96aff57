// The hot path: scan every window of a packed sequence set, canonicalise the k-mer, probe the
// read and assembly multiplicities, and either accumulate the joint (v_read, v_asm) histogram
// (-hist, merfin-histogram.C:54-91) or write the pair per position (-dump / path scoring,
// merfin-dump.C:44-67, varMer.C:66-90).  All integer work; the K* doubles are replayed on the host.
//
// Persistent CTAs pull 2048-window tiles from a global counter.  The tile's 2-bit words and
// validity bits (528 + 272 bytes incl. a 64-base halo) are staged in shared memory by two
// 1-D TMA bulk copies (cp.async.bulk + mbarrier), double-buffered against the probes of the
// previous tile.  A thread owns 8 consecutive windows: it slices them out of a 128-bit
// register window with shifts (no serial rolling dependency between threads), issues all its
// bucket loads back to back (8 independent 32-byte sectors in flight per thread), then
// resolves them.
//
// MODE bit 0: DUMP (write pairs per position) instead of histogram.
// MODE bit 1: JOINT: one probe of the fused (read, assembly) table per window (joint.cuh)
//             instead of one probe per database.
#pragma once
#include "joint.cuh"
#include "table_host.h"

#include <algorithm>

#define SCAN_THREADS 256
#define WPT 8                              // windows per thread per tile
#define SEQ_WORDS (MFN_TILE / 32 + 2)      // 66 words = 528 B
#define OK_WORDS (MFN_TILE / 64 + 2)       // 34 words = 272 B
#define SH_R 512u                          // shared-memory tier of the joint histogram
#define SH_A 16u
#define SCAN_DUMP 1
#define SCAN_JOINT 2

static_assert(SCAN_THREADS * WPT == MFN_TILE, "tile geometry");
static_assert((SEQ_WORDS * 8) % 16 == 0 && (OK_WORDS * 8) % 16 == 0, "TMA sizes are multiples of 16 B");

struct ScanParams {
  DevTable R, A;
  JointDev J;
  const unsigned long long *seq2, *ok, *rec_start;
  uint32_t nseq;
  uint64_t tile_begin, tile_end;
  unsigned long long *tile_counter;        // zeroed before launch
  // -hist
  unsigned long long *acc;
  PairTable sparse;
  const unsigned long long *zmap;          // bit v set: readK(v) == 0 (v < zlen); NULL -> only v == 0
  uint32_t zlen;
  // -dump
  uint32_t *out_vr, *out_va;
  uint8_t *out_valid;
};

// ---- mbarrier / TMA bulk copy -------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- rare pairs outside the dense tier -----------------------------------------------------
static __device__ __noinline__ void sparse_add(const PairTable &t, unsigned long long *acc, uint32_t vr, uint32_t va,
                                               unsigned long long n) {
  unsigned long long key = ((unsigned long long)vr << 32) | va;
  if (key == MFN_EMPTY64) { atomicAdd(&acc[3], n); return; }
  uint32_t i = mfn_oa_hash(key) & t.mask;
  for (uint32_t p = 0; p <= t.mask; p++) {
    unsigned long long cur = t.keys[i];
    if (cur == MFN_EMPTY64) {
      cur = atomicCAS(&t.keys[i], MFN_EMPTY64, key);
      if (cur == MFN_EMPTY64) { cur = key; atomicAdd(&acc[1], 1ull); atomicAdd(t.local_n, 1ull); }
    }
    if (cur == key) { atomicAdd(&t.counts[i], n); return; }
    i = (i + 1) & t.mask;
  }
  atomicAdd(&acc[2], n);                     // table full: reported as MFN_ERR_CAPACITY
}

template <int MODE>
__global__ void __launch_bounds__(SCAN_THREADS, (MODE & SCAN_JOINT) ? 2 : 1) k_scan(const __grid_constant__ ScanParams P) {
  constexpr bool DUMP = (MODE & SCAN_DUMP) != 0;
  constexpr bool JOINT = (MODE & SCAN_JOINT) != 0;
  __shared__ __align__(16) unsigned long long s_seq[2][SEQ_WORDS];
  __shared__ __align__(16) unsigned long long s_ok[2][OK_WORDS];
  __shared__ __align__(8) unsigned long long s_bar[2];
  __shared__ unsigned long long s_tile[2];
  __shared__ uint32_t s_hist[DUMP ? 1 : SH_R * SH_A];

  const uint32_t tid = threadIdx.x;
  const uint32_t k = P.R.k;
  const uint64_t kmask = P.R.kmask;
  const uint64_t kbits = (1ull << k) - 1;
  const uint64_t cmask = 0xaaaaaaaaaaaaaaaaull & kmask;

  if (!DUMP) for (uint32_t i = tid; i < SH_R * SH_A; i += SCAN_THREADS) s_hist[i] = 0;
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto fetch = [&](int buf) {                // thread 0: claim the next tile and start its TMA
    unsigned long long t = P.tile_begin + atomicAdd(P.tile_counter, 1ull);
    s_tile[buf] = t;
    if (t < P.tile_end) {
      mbar_expect_tx(&s_bar[buf], (SEQ_WORDS + OK_WORDS) * 8);
      tma_load_1d(s_seq[buf], P.seq2 + t * (MFN_TILE / 32), SEQ_WORDS * 8, &s_bar[buf]);
      tma_load_1d(s_ok[buf], P.ok + t * (MFN_TILE / 64), OK_WORDS * 8, &s_bar[buf]);
    }
  };
  if (tid == 0) fetch(0);

  unsigned long long my_valid = 0;           // windows this thread counted (for acc[0])
  // valid / missing windows of record cur_rec seen by this thread, flushed when the record changes
  uint32_t cur_rec = 0xffffffffu, n_ok0 = 0, n_miss0 = 0;
  auto flush_rec = [&]() {
    if (DUMP) return;
    uint32_t a = n_ok0, b = n_miss0;
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o); }
    if ((tid & 31) == 0 && cur_rec < P.nseq) {
      if (a) atomicAdd(&P.acc[MFN_ACC_HEADER + MFN_DENSE_R * MFN_DENSE_A + 2 * cur_rec], (unsigned long long)a);
      if (b) atomicAdd(&P.acc[MFN_ACC_HEADER + MFN_DENSE_R * MFN_DENSE_A + 2 * cur_rec + 1], (unsigned long long)b);
    }
    n_ok0 = 0; n_miss0 = 0;
  };

  for (uint32_t it = 0;; it++) {
    const int buf = it & 1;
    __syncthreads();                         // s_tile[buf] visible; everyone is done with buf^1
    const uint64_t tile = s_tile[buf];
    if (tile >= P.tile_end) break;
    if (tid == 0) fetch(buf ^ 1);
    mbar_wait(&s_bar[buf], (it >> 1) & 1);

    // ---- this thread's 8 windows: a 128-bit slice of bases and a 64-bit slice of validity
    const uint32_t j0 = tid * WPT;
    const uint64_t g0 = tile * MFN_TILE + j0;
    uint64_t lo, hi, okb;
    {
      const uint32_t w = j0 >> 5, sh = (j0 & 31) * 2;
      uint64_t a = s_seq[buf][w], b = s_seq[buf][w + 1];
      lo = sh ? ((a >> sh) | (b << (64 - sh))) : a;
      hi = b >> sh;
      const uint32_t wo = j0 >> 6, so = j0 & 63;
      uint64_t c = s_ok[buf][wo], d = s_ok[buf][wo + 1];
      okb = so ? ((c >> so) | (d << (64 - so))) : c;
    }

    // record of the tile's first window (binary search, uniform) and of this thread's windows
    uint32_t r0;
    {
      const uint64_t gt = tile * MFN_TILE;
      uint32_t a = 0, b = P.nseq;
      while (a < b) { uint32_t m = (a + b + 1) >> 1; if (P.rec_start[m] <= gt) a = m; else b = m - 1; }
      r0 = a;
    }
    if (r0 != cur_rec) { flush_rec(); cur_rec = r0; }        // uniform across the CTA
    uint32_t rec = r0;
    uint64_t next_start = (rec < P.nseq) ? P.rec_start[rec + 1] : ~0ull;

    // forward / reverse-complement k-mer of window uu (kmers-tiny.H:82-83)
    auto kmers_of = [&](int uu, uint64_t &f, uint64_t &r) {
      uint64_t x = uu ? ((lo >> (2 * uu)) | (hi << (64 - 2 * uu))) : lo;
      x &= kmask;                                          // base g in the low bits
      r = x ^ cmask;
      f = mfn_revcomp(r, k);
    };
    // what happens to the integer pair of window uu
    auto emit = [&](int uu, bool valid, uint32_t vr, uint32_t va) {
      const uint64_t g = g0 + uu;
      while (g >= next_start) { rec++; next_start = (rec < P.nseq) ? P.rec_start[rec + 1] : ~0ull; }
      if (DUMP) {
        // positions are reported in the caller's coordinates: one separator per earlier record
        if (rec < P.nseq && g + 1 < next_start) {
          const uint64_t s = g - rec;
          P.out_vr[s] = vr; P.out_va[s] = va; P.out_valid[s] = valid ? 1 : 0;
        }
      } else if (valid) {
        bool miss = (vr == 0);
        if (vr != 0 && vr < P.zlen) miss = (__ldg(&P.zmap[vr >> 6]) >> (vr & 63)) & 1;
        if (rec == r0) { n_ok0++; n_miss0 += miss; }
        else {
          atomicAdd(&P.acc[MFN_ACC_HEADER + MFN_DENSE_R * MFN_DENSE_A + 2 * rec], 1ull);
          if (miss) atomicAdd(&P.acc[MFN_ACC_HEADER + MFN_DENSE_R * MFN_DENSE_A + 2 * rec + 1], 1ull);
        }
        my_valid++;
        if (vr < SH_R && va < SH_A) atomicAdd(&s_hist[va * SH_R + vr], 1u);
        else if (vr < MFN_DENSE_R && va < MFN_DENSE_A) atomicAdd(&P.acc[MFN_ACC_HEADER + va * MFN_DENSE_R + vr], 1ull);
        else sparse_add(P.sparse, P.acc, vr, va, 1ull);
      }
    };

    if (JOINT) {
      // one line of the fused table per window; 4 home sectors in flight per thread per batch.
      // Lookups that need more than the home sector (spilled entry, side table, saturated
      // value: a few %) are finished in a second, lane-parallel loop so that one lane's slow
      // path does not serialise the whole warp once per window.
#pragma unroll
      for (int half = 0; half < 2; half++) {
        uint64_t ck[4], ln[4], tg[4];
        uint32_t vr[4], va[4];
        bool val[4], pal[4];
        Bucket b[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int uu = half * 4 + u;
          uint64_t f, r;
          kmers_of(uu, f, r);
          val[u] = ((okb >> uu) & kbits) == kbits;
          pal[u] = (f == r);
          ck[u] = f < r ? f : r;
          if (val[u]) {
            joint_locate(P.J, ck[u], ln[u], tg[u]);
            joint_load_sector(P.J, ln[u], (uint32_t)tg[u] & 3u, b[u]);
          }
        }
        uint32_t pending = 0;
#pragma unroll
        for (int u = 0; u < 4; u++) {
          vr[u] = 0; va[u] = 0;
          if (val[u]) {
            uint64_t pay = 0;
            const bool hit = joint_match_fast(P.J, b[u], tg[u], pay);
            va[u] = (uint32_t)pay & P.J.amax;
            vr[u] = (uint32_t)(pay >> P.J.abits) & P.J.rmax;
            if (!hit || vr[u] == P.J.rmax || va[u] == P.J.amax) pending |= 1u << u;
          }
        }
        while (__any_sync(0xffffffffu, pending != 0)) {
          if (pending) {
            const int u = __ffs(pending) - 1;
            pending &= pending - 1;
            const uint64_t c = u == 0 ? ck[0] : u == 1 ? ck[1] : u == 2 ? ck[2] : ck[3];
            const uint64_t l = u == 0 ? ln[0] : u == 1 ? ln[1] : u == 2 ? ln[2] : ln[3];
            const uint64_t t = u == 0 ? tg[0] : u == 1 ? tg[1] : u == 2 ? tg[2] : tg[3];
            uint32_t r2, a2;
            joint_full(P.J, c, l, t, r2, a2);
#pragma unroll
            for (int q = 0; q < 4; q++) if (q == u) { vr[q] = r2; va[q] = a2; }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          if (val[u] && pal[u]) { vr[u] += vr[u]; va[u] += va[u]; }   // value(f) + value(r), f == r (Q2)
          emit(half * 4 + u, val[u], vr[u], va[u]);
        }
      }
    } else {
#pragma unroll
      for (int half = 0; half < 2; half++) {
        uint64_t fk[4], rk[4];
        bool val[4];
        Bucket br[4], ba[4];
        uint64_t tr[4], ta[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int uu = half * 4 + u;
          kmers_of(uu, fk[u], rk[u]);
          val[u] = ((okb >> uu) & kbits) == kbits;
          if (val[u]) {
            const uint64_t c = fk[u] < rk[u] ? fk[u] : rk[u];
            uint64_t bkt;
            mfn_locate(P.R, P.R.probe_both ? fk[u] : c, bkt, tr[u]);
            mfn_load_bucket(P.R.lv[0].slots, bkt, br[u]);
            mfn_locate(P.A, P.A.probe_both ? fk[u] : c, bkt, ta[u]);
            mfn_load_bucket(P.A.lv[0].slots, bkt, ba[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          uint32_t vr = 0, va = 0;
          if (val[u]) {
            const uint64_t f = fk[u], r = rk[u], c = f < r ? f : r;
            vr = mfn_resolve(P.R, P.R.probe_both ? f : c, br[u], tr[u]);
            va = mfn_resolve(P.A, P.A.probe_both ? f : c, ba[u], ta[u]);
            // value(f) + value(r), uint32 adds (merfin-globals.C:107-108, Q2): a canonical table
            // holds at most one of the two unless f == r, which then counts twice
            if (P.R.probe_both) vr += mfn_value(P.R, r); else if (f == r) vr += vr;
            if (P.A.probe_both) va += mfn_value(P.A, r); else if (f == r) va += va;
          }
          emit(half * 4 + u, val[u], vr, va);
        }
      }
    }
  }

  flush_rec();
  if (!DUMP) {
    __syncthreads();
    for (uint32_t i = tid; i < SH_R * SH_A; i += SCAN_THREADS) {
      uint32_t c = s_hist[i];
      if (c) atomicAdd(&P.acc[MFN_ACC_HEADER + (i / SH_R) * MFN_DENSE_R + (i % SH_R)], (unsigned long long)c);
    }
    for (int o = 16; o > 0; o >>= 1) my_valid += __shfl_down_sync(0xffffffffu, my_valid, o);
    if ((tid & 31) == 0 && my_valid) atomicAdd(&P.acc[0], my_valid);
  }
}

// One translation unit per instantiation (scan_mode*.cu): ptxas 12.9 crashes when several of
// these entries share a module.
int mfn_scan_grid(mfn_ctx *ctx, int mode);
void mfn_scan_launch(const ScanParams &P, int mode, int grid, cudaStream_t stream);

#define MFN_SCAN_INSTANCE(MODE)                                                                     \
  int mfn_scan_grid_##MODE(mfn_ctx *ctx) {                                                          \
    int per_sm = 0;                                                                                 \
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_scan<MODE>, SCAN_THREADS, 0);          \
    if (per_sm < 1) per_sm = 1;                                                                     \
    const char *env = getenv("MFN_SCAN_CTAS_PER_SM");                                               \
    if (env && atoi(env) > 0) per_sm = std::min(per_sm, atoi(env));                                 \
    return ctx->sm_count * per_sm;                                                                  \
  }                                                                                                 \
  void mfn_scan_launch_##MODE(const ScanParams &P, int grid, cudaStream_t stream) {                 \
    k_scan<MODE><<<grid, SCAN_THREADS, 0, stream>>>(P);                                             \
  }
