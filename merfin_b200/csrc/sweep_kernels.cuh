// Range-partitioned -hist scan ("sweep").
//
// A direct scan spends one random 128-byte HBM line per window, and the B200 sustains only
// ~36.5e9 of those per second.  But the windows of a whole assembly hit
// every line of the joint table many times over, so the random accesses can be traded for
// streaming ones:
//
//   pass 1  k_sweep_scatter   every valid window -> h = mix(canonical k-mer) -> (line, tag) of the
//                             joint table.  Lines are grouped into bins of 2^lbits consecutive
//                             lines; (line in bin, tag, record id) is appended to the window's
//                             bin.  A CTA sorts 8192 windows by bin in shared memory and writes
//                             each bin's run with coalesced stores.
//   pass 2  k_sweep_probe     bins are processed in order by the whole grid.  A bin is a
//                             contiguous slice of the joint table of <= 32 MB, which stays
//                             L2-resident while the bin's windows probe it: each line comes from
//                             HBM once, every further probe is an L2 hit.  No hashing, no
//                             division here: the entry already holds line and tag.
//
// DRAM traffic per window: 8 B written + 8 B read + (table bytes / windows) instead of 128 B.
// Integer results are identical to the direct scan (same lookups, sums are order-free).
#pragma once
#include "scan_kernel.cuh"

#define SW_THREADS 256
#define SW_CHUNK_TILES 4
#define SW_CHUNK (SW_CHUNK_TILES * MFN_TILE)      // 8192 windows per CTA round
#define SW_MAX_BINS 2048
#define SW_SLICE 1024                              // entries per probe work item (one WARP: 32 per lane)
#define SW_NOSPACE 0xffffffffu

struct SweepParams {
  JointDev J;
  const unsigned long long *seq2, *ok, *rec_start;
  uint32_t nseq;
  uint64_t tile_begin, tile_end;
  uint32_t lbits, tbits, nbins, nbins_pad;   // bin = line >> lbits; entry low word = (line & L-1) << tbits | tag
  unsigned long long *entries;               // bin b: entries[b * cap .. b * cap + n_b); entry = rec << 32 | low word
  uint64_t cap;
  unsigned long long *bin_count;             // [nbins_pad] reserved entries (may exceed cap)
  unsigned long long *bin_valid;             // [nbins_pad] first reservation that did not fit (~0: none)
  unsigned long long *over_lt;               // windows that found no room in their bin: line << 32 | tag ...
  uint32_t *over_rec;                        //   ... and record id
  unsigned long long *over_n;
  uint64_t over_cap;
  unsigned long long *counter;               // work distribution (one per pass)
  unsigned long long *acc;
  PairTable sparse;
  const unsigned long long *zmap;
  uint32_t zlen;
};

// ------------------------------------------------------------------------------------------
// pass 1
// ------------------------------------------------------------------------------------------
template <int UNUSED>
__global__ void __launch_bounds__(SW_THREADS, 2) k_sweep_scatter(const __grid_constant__ SweepParams P) {
  extern __shared__ __align__(16) unsigned char sw_smem[];            // 88 KB: dynamic
  uint32_t *s_low = (uint32_t *)sw_smem;                               // [SW_CHUNK]
  uint32_t *s_meta = s_low + SW_CHUNK;                                 // [SW_CHUNK] bin | record delta << 12
  uint32_t *s_cnt = s_meta + SW_CHUNK, *s_off = s_cnt + SW_MAX_BINS, *s_base = s_off + SW_MAX_BINS;
  __shared__ uint32_t s_warp[SW_THREADS / 32];
  __shared__ unsigned long long s_chunk;

  const uint32_t tid = threadIdx.x;
  const uint32_t k = P.J.k;
  const uint64_t kmask = P.J.kmask;
  const uint64_t kbits = (1ull << k) - 1;
  const uint64_t cmask = 0xaaaaaaaaaaaaaaaaull & kmask;
  const uint32_t lmask = (1u << P.lbits) - 1u;
  const uint64_t nchunks = ceil_div_u64(P.tile_end - P.tile_begin, SW_CHUNK_TILES);
  const uint32_t per_thread_bins = P.nbins_pad / SW_THREADS;

  unsigned long long my_valid = 0;
  uint32_t cur_rec = 0xffffffffu, n_ok0 = 0;
  auto flush_rec = [&]() {
    uint32_t a = n_ok0;
    for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
    if ((tid & 31) == 0 && cur_rec < P.nseq && a)
      atomicAdd(&P.acc[MFN_ACC_HEADER + MFN_DENSE_R * MFN_DENSE_A + 2 * cur_rec], (unsigned long long)a);
    n_ok0 = 0;
  };

  while (true) {
    __syncthreads();
    if (tid == 0) s_chunk = atomicAdd(P.counter, 1ull);
    for (uint32_t i = tid; i < P.nbins_pad; i += SW_THREADS) s_cnt[i] = 0;
    __syncthreads();
    const uint64_t chunk = s_chunk;
    if (chunk >= nchunks) break;
    const uint64_t tile0 = P.tile_begin + chunk * SW_CHUNK_TILES;

    // record of the chunk's first window (uniform); does the chunk stay inside it?
    uint32_t r0;
    {
      const uint64_t gt = tile0 * MFN_TILE;
      uint32_t a = 0, b = P.nseq;
      while (a < b) { uint32_t m = (a + b + 1) >> 1; if (P.rec_start[m] <= gt) a = m; else b = m - 1; }
      r0 = a;
    }
    if (r0 != cur_rec) { flush_rec(); cur_rec = r0; }
    const uint64_t r0_end = (r0 < P.nseq) ? P.rec_start[r0 + 1] : ~0ull;
    const bool one_record = (tile0 + SW_CHUNK_TILES) * MFN_TILE <= r0_end;

    // ---- hash this thread's 32 windows; rank them inside their bins ------------------------
    uint32_t low[SW_CHUNK_TILES * WPT], meta[SW_CHUNK_TILES * WPT];     // meta: bin | rank << 12 | rdelta << 25
#pragma unroll
    for (int q = 0; q < SW_CHUNK_TILES; q++) {
      const uint64_t tile = tile0 + q;
      const uint64_t g0 = tile * MFN_TILE + tid * WPT;
      uint64_t lo = 0, hi = 0, okb = 0;
      if (tile < P.tile_end) {
        const uint64_t w = g0 >> 5, sh = (g0 & 31) * 2;
        const uint64_t a = __ldg(&P.seq2[w]), b = __ldg(&P.seq2[w + 1]);
        lo = sh ? ((a >> sh) | (b << (64 - sh))) : a;
        hi = b >> sh;
        const uint64_t wo = g0 >> 6, so = g0 & 63;
        const uint64_t c = __ldg(&P.ok[wo]), d = __ldg(&P.ok[wo + 1]);
        okb = so ? ((c >> so) | (d << (64 - so))) : c;
      }
      uint32_t rec = r0;
      uint64_t next_start = r0_end;
#pragma unroll
      for (int u = 0; u < WPT; u++) {
        const int idx = q * WPT + u;
        meta[idx] = 0xffffffffu;
        low[idx] = 0;
        if (((okb >> u) & kbits) != kbits) continue;
        uint64_t x = u ? ((lo >> (2 * u)) | (hi << (64 - 2 * u))) : lo;
        x &= kmask;                                             // base g in the low 2 bits: reversed f
        const uint64_t r = x ^ cmask, f = mfn_rev2(x) >> (64 - 2 * k);
        uint64_t line, tag;
        joint_locate(P.J, f < r ? f : r, line, tag);
        uint32_t rd = 0;
        if (!one_record) {
          const uint64_t g = g0 + u;
          while (g >= next_start) { rec++; next_start = (rec < P.nseq) ? P.rec_start[rec + 1] : ~0ull; }
          rd = rec - r0;
        }
        my_valid++;
        if (rd == 0) n_ok0++;
        else atomicAdd(&P.acc[MFN_ACC_HEADER + MFN_DENSE_R * MFN_DENSE_A + 2 * rec], 1ull);
        if (rd > 126) {                                   // a chunk spanning > 126 records: side list
          unsigned long long o = atomicAdd(P.over_n, 1ull);
          if (o < P.over_cap) { P.over_lt[o] = (line << 32) | tag; P.over_rec[o] = rec; } else atomicAdd(&P.acc[2], 1ull);
          continue;
        }
        const uint32_t bin = (uint32_t)(line >> P.lbits);
        const uint32_t rank = atomicAdd(&s_cnt[bin], 1u);
        low[idx] = (((uint32_t)line & lmask) << P.tbits) | (uint32_t)tag;
        meta[idx] = bin | (rank << 12) | (rd << 25);
      }
    }
    __syncthreads();

    // ---- exclusive scan of the bin counts; reserve space in the global bins ----------------
    {
      uint32_t sum = 0;
      const uint32_t b0 = tid * per_thread_bins;
      for (uint32_t i = 0; i < per_thread_bins; i++) sum += s_cnt[b0 + i];
      uint32_t incl = sum;
      for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= (uint32_t)o) incl += v; }
      if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
      __syncthreads();
      uint32_t wbase = 0;
      for (uint32_t w = 0; w < (tid >> 5); w++) wbase += s_warp[w];
      uint32_t run = wbase + incl - sum;
      // reserve space in the global bins: all atomics of this thread in flight together
      unsigned long long got[SW_MAX_BINS / SW_THREADS];
#pragma unroll
      for (uint32_t i = 0; i < SW_MAX_BINS / SW_THREADS; i++) {
        got[i] = 0;
        if (i < per_thread_bins) {
          const uint32_t c = s_cnt[b0 + i];
          if (c) got[i] = atomicAdd(&P.bin_count[b0 + i], (unsigned long long)c);
        }
      }
#pragma unroll
      for (uint32_t i = 0; i < SW_MAX_BINS / SW_THREADS; i++) {
        if (i < per_thread_bins) {
          const uint32_t c = s_cnt[b0 + i];
          s_off[b0 + i] = run;
          run += c;
          uint32_t base = (uint32_t)got[i];
          if (c && got[i] + c > P.cap) { atomicMin(&P.bin_valid[b0 + i], got[i]); base = SW_NOSPACE; }
          s_base[b0 + i] = base;
        }
      }
    }
    __syncthreads();

    // ---- counting sort into shared memory --------------------------------------------------
#pragma unroll
    for (int idx = 0; idx < SW_CHUNK_TILES * WPT; idx++) {
      const uint32_t m = meta[idx];
      if (m == 0xffffffffu) continue;
      const uint32_t bin = m & 0xfffu, rank = (m >> 12) & 0x1fffu, rd = m >> 25;
      const uint32_t p = s_off[bin] + rank;
      s_low[p] = low[idx];
      s_meta[p] = bin | (rd << 12);
    }
    __syncthreads();

    // ---- coalesced copy-out: consecutive threads write consecutive entries of a bin's run --
    {
      const uint32_t last = P.nbins_pad - 1;
      const uint32_t total = s_off[last] + s_cnt[last];
      for (uint32_t i = tid; i < total; i += SW_THREADS) {
        const uint32_t m = s_meta[i];
        const uint32_t bin = m & 0xfffu, rec = r0 + (m >> 12);
        const uint32_t base = s_base[bin];
        if (base != SW_NOSPACE) {
          P.entries[(uint64_t)bin * P.cap + base + (i - s_off[bin])] = ((unsigned long long)rec << 32) | s_low[i];
        } else {
          unsigned long long o = atomicAdd(P.over_n, 1ull);
          const uint32_t lw = s_low[i];
          if (o < P.over_cap) {
            P.over_lt[o] = ((((uint64_t)bin << P.lbits) | (lw >> P.tbits)) << 32) | (lw & ((1u << P.tbits) - 1u));
            P.over_rec[o] = rec;
          } else atomicAdd(&P.acc[2], 1ull);               // reported as MFN_ERR_CAPACITY at fetch
        }
      }
    }
  }
  flush_rec();
  for (int o = 16; o > 0; o >>= 1) my_valid += __shfl_down_sync(0xffffffffu, my_valid, o);
  if ((tid & 31) == 0 && my_valid) atomicAdd(&P.acc[0], my_valid);
}

// ------------------------------------------------------------------------------------------
// pass 2 (and the overflow list: OVER = true probes (line, tag) pairs from the side list)
// ------------------------------------------------------------------------------------------
template <bool OVER>
__global__ void __launch_bounds__(SW_THREADS, 3) k_sweep_probe(const __grid_constant__ SweepParams P) {
  __shared__ uint32_t s_hist[SH_R * SH_A];
  const uint32_t tid = threadIdx.x;
  for (uint32_t i = tid; i < SH_R * SH_A; i += SW_THREADS) s_hist[i] = 0;
  __syncthreads();

  const uint64_t slices_per_bin = ceil_div_u64(P.cap, SW_SLICE);
  const uint64_t nitems = OVER ? ceil_div_u64(min(*P.over_n, (unsigned long long)P.over_cap), SW_SLICE)
                               : (uint64_t)P.nbins * slices_per_bin;
  const bool even_k = (P.J.k & 1u) == 0;
  const uint32_t tmask = (1u << P.tbits) - 1u;

  auto account = [&](uint32_t line, uint32_t tag, uint32_t rec, uint32_t vr, uint32_t va) {
    if (even_k) {                                          // palindromes count twice (Q2)
      const uint64_t c = mfn_unmix((uint64_t)line * P.J.D + tag, P.J.seed, P.J.kmask, P.J.k);
      if (c == mfn_revcomp(c, P.J.k)) { vr += vr; va += va; }
    }
    bool miss = (vr == 0);
    if (vr != 0 && vr < P.zlen) miss = (__ldg(&P.zmap[vr >> 6]) >> (vr & 63)) & 1;
    if (miss) atomicAdd(&P.acc[MFN_ACC_HEADER + MFN_DENSE_R * MFN_DENSE_A + 2 * rec + 1], 1ull);
    if (vr < SH_R && va < SH_A) atomicAdd(&s_hist[va * SH_R + vr], 1u);
    else if (vr < MFN_DENSE_R && va < MFN_DENSE_A) atomicAdd(&P.acc[MFN_ACC_HEADER + va * MFN_DENSE_R + vr], 1ull);
    else sparse_add(P.sparse, P.acc, vr, va, 1ull);
  };

  // per-warp queue of lookups that need more than the home sector: (line, tag, record)
  __shared__ uint32_t s_q[SW_THREADS / 32][3][64];
  const uint32_t lane = tid & 31, wid = tid >> 5;
  uint32_t qn = 0;                                           // warp-uniform
  auto drain = [&](bool all) {
    while (qn >= 32 || (all && qn > 0)) {
      const uint32_t take = qn < 32 ? qn : 32;
      if (lane < take) {
        const uint32_t i = qn - take + lane;
        const uint32_t l = s_q[wid][0][i], t = s_q[wid][1][i], rec = s_q[wid][2][i];
        uint32_t r2, a2;                                     // side tables are keyed by k-mer
        joint_full(P.J, mfn_unmix((uint64_t)l * P.J.D + t, P.J.seed, P.J.kmask, P.J.k), l, t, r2, a2);
        account(l, t, rec, r2, a2);
      }
      qn -= take;
      __syncwarp();
    }
  };

  // work items are handed out per WARP (no CTA barrier in the loop)
  while (true) {
    uint64_t item = 0;
    if (lane == 0) item = atomicAdd(P.counter, 1ull);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= nitems) break;
    uint64_t n, first;                 // entries [first, first + n) of this item
    const unsigned long long *src;
    uint32_t line0 = 0;
    if (OVER) {
      const uint64_t tot = min(*P.over_n, (unsigned long long)P.over_cap);
      first = item * SW_SLICE; n = min((uint64_t)SW_SLICE, tot - first); src = P.over_lt;
    } else {
      const uint64_t bin = item / slices_per_bin, s = item % slices_per_bin;
      const uint64_t cnt = min(min(P.bin_count[bin], P.bin_valid[bin]), (unsigned long long)P.cap);
      if (s * SW_SLICE >= cnt) continue;
      first = bin * P.cap + s * SW_SLICE; n = min((uint64_t)SW_SLICE, cnt - s * SW_SLICE); src = P.entries;
      line0 = (uint32_t)bin << P.lbits;
    }
    // each thread: batches of 4 consecutive entries (32 B per lane, 1 KB per warp, coalesced)
    const uint32_t nbatch = (uint32_t)ceil_div_u64(n, 32 * 4);             // uniform: every lane of the warp
    auto load_entries = [&](uint32_t bt, unsigned long long (&e)[4]) {
      const uint64_t i0 = (uint64_t)bt * 32 * 4 + lane * 4;
      if (i0 + 4 <= n) {
        // streamed once: keep it from evicting the table slice
        asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.u64 {%0,%1,%2,%3}, [%4];"
                     : "=l"(e[0]), "=l"(e[1]), "=l"(e[2]), "=l"(e[3]) : "l"(src + first + i0));
      } else {
#pragma unroll
        for (int u = 0; u < 4; u++) e[u] = (i0 + u < n) ? src[first + i0 + u] : 0;
      }
    };
    unsigned long long enext[4];
    load_entries(0, enext);
#pragma unroll 1                                                           // (every lane takes part in the warp votes below)
    for (uint32_t bt = 0; bt < nbatch; bt++) {
      const uint64_t i0 = (uint64_t)bt * 32 * 4 + lane * 4;
      unsigned long long e[4];
#pragma unroll
      for (int u = 0; u < 4; u++) e[u] = enext[u];
      if (bt + 1 < nbatch) load_entries(bt + 1, enext);        // in flight while this batch probes
      uint32_t ln[4], tg[4], rc[4];
      bool val[4];
      Bucket b[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        val[u] = i0 + u < n;
        if (OVER) { ln[u] = (uint32_t)(e[u] >> 32); tg[u] = (uint32_t)e[u]; rc[u] = val[u] ? P.over_rec[first + i0 + u] : 0; }
        else {
          const uint32_t lw = (uint32_t)e[u];
          ln[u] = line0 + (lw >> P.tbits); tg[u] = lw & tmask; rc[u] = (uint32_t)(e[u] >> 32);
        }
        if (val[u]) joint_load_sector(P.J, ln[u], tg[u] & 3u, b[u]);
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        uint64_t pay = 0;
        const bool hit = val[u] && joint_match_fast(P.J, b[u], tg[u], pay);
        const uint32_t a1 = (uint32_t)pay & P.J.amax, r1 = (uint32_t)(pay >> P.J.abits) & P.J.rmax;
        // the few lookups the home sector cannot answer are queued per warp and finished 32 at a
        // time with every lane busy, instead of dragging the whole warp through the slow path
        const bool slow = val[u] && (!hit || r1 == P.J.rmax || a1 == P.J.amax);
        if (val[u] && !slow) account(ln[u], tg[u], rc[u], r1, a1);
        const uint32_t m = __ballot_sync(0xffffffffu, slow);
        if (slow) {
          const uint32_t pos = qn + __popc(m & ((1u << lane) - 1u));
          s_q[wid][0][pos] = ln[u]; s_q[wid][1][pos] = tg[u]; s_q[wid][2][pos] = rc[u];
        }
        qn += __popc(m);
        __syncwarp();
        if (qn >= 32) drain(false);
      }
    }
  }
  drain(true);
  __syncthreads();
  for (uint32_t i = tid; i < SH_R * SH_A; i += SW_THREADS) {
    uint32_t c = s_hist[i];
    if (c) atomicAdd(&P.acc[MFN_ACC_HEADER + (i / SH_R) * MFN_DENSE_R + (i % SH_R)], (unsigned long long)c);
  }
}

// k_sweep_probe<true> lives in its own translation unit (sweep_over.cu): ptxas 12.9 crashes when
// two entries that share an out-of-line device function are compiled in one module.
void mfn_sweep_launch_over(const SweepParams &P, int grid, cudaStream_t stream);
