// Pass 1 of the range-partitioned -hist scan, third version ("scatter3"): same job and same entry
// format as k_sweep_scatter (sweep_kernels.cuh), for odd 16 < k < 32 with a narrow low part and at
// most S3_MAX_BINS bins -- the benchmark's case.  What changed, and what the ncu profile of the
// previous version (profiles/r02_ncu_scatter_v6_summary.txt) said about it:
//
//   * 4 CTAs of 256 threads per SM, chunks of 4096 windows (16 per thread).  The pass has four
//     phases that each load a different unit (hash: ALU, sort: shared-memory stores, copy-out: LSU);
//     more and smaller CTAs in different phases keep more of the SM busy than two big ones.
//   * a thread's 16 low words wait for the sort phase in shared memory (4 x STS.128, conflict free)
//     instead of registers: with 32 live words per thread the 64-register kernel spilled, and the
//     local-memory loads of the sort phase were 10 % of all stall samples.
//   * a straight-line fast path for the usual group of 8 windows (all valid, chunk inside one
//     record, odd k so no k-mer is its own reverse complement): no per-window predicate, no record
//     delta, no branch around the shared-memory atomic (the predicated atom became BSSY/BRA/BSYNC).
//   * the counters are zeroed by the scan that reads them and the destination of a bin's run is a
//     32-bit entry index, so the copy-out is LDS.64 + LDS.32 + IMAD.WIDE + STG.64 per entry and a
//     chunk needs four barriers instead of five.
//   * a run that crosses the end of its bin is split: what fits is stored, the rest goes to the side
//     list (the probe pass reads min(count, cap) entries, so bin_valid is not needed here).
#pragma once
#include "sweep_kernels.cuh"

#define S3_MAX_BINS 512
// staging word: bits 0..1 low part bits 32..33 | bits 2..12 bin | bits 13..18 record delta | rank above (12 bits for
// 4096-window chunks, 13 for 8192)
#define S3_BIN_FIELD 0x1ffcu
#define S3_RD_SHIFT 13
#define S3_RD_MASK 0x3fu

// shape: TILES = 2 -> 256 threads, 4096-window chunks, 4 CTAs per SM (up to 256 bins);
//        TILES = 4 -> 512 threads, 8192-window chunks, 2 CTAs per SM: with 512 bins the runs of a 4096-window chunk
//        are 8 entries (64 B) long and the pass takes 14.7 ms instead of 10.6 at 3 Gbp; twice the chunk, same runs
static inline size_t s3_smem_bytes(uint32_t nbins_pad, int tiles) { return (size_t)tiles * MFN_TILE * 12 + (size_t)nbins_pad * 12; }

// DUMP: the -dump / path-scoring flavour (sweep_dump.cu): no records and no counters; every valid window's
// entry index is written to P.where[window] (0xffffffff: no window), so that the pair the probe pass leaves in
// the entry can be gathered back in position order afterwards.
template <int KK, int GEO, int TILES, bool DUMP = false>
__global__ void __launch_bounds__(TILES * 128, TILES == 2 ? 4 : 2) k_sweep_scatter3(const __grid_constant__ SweepParams P) {
  static_assert(KK > 16 && KK < 32 && (KK & 1), "scatter3: odd 16 < k < 32");
  static_assert(TILES == 2 || TILES == 4, "scatter3 shapes");
  constexpr uint32_t S3_THREADS = TILES * 128, S3_TILES = TILES, S3_CHUNK = TILES * MFN_TILE, S3_NW = 16;
  constexpr uint32_t S3_SEQ_WORDS = S3_CHUNK / 32 + 2, S3_OK_WORDS = S3_CHUNK / 64 + 2, S3_NWARPS = S3_THREADS / 32;
  constexpr uint32_t S3_RANK_SHIFT = TILES == 2 ? 20 : 19, S3_KEEP_MASK = (1u << S3_RANK_SHIFT) - 1u;
  static_assert(S3_CHUNK <= (1u << (32 - S3_RANK_SHIFT)), "rank field");
  static_assert((S3_SEQ_WORDS * 8) % 16 == 0 && (S3_OK_WORDS * 8) % 16 == 0, "TMA sizes are multiples of 16 B");
  typedef Kmer2<KK> H;
  extern __shared__ __align__(16) unsigned char sw_smem[];
  unsigned long long *s_ent = (unsigned long long *)sw_smem;            // [S3_CHUNK] entries sorted by bin
  uint4 *s_low4 = (uint4 *)(s_ent + S3_CHUNK);                            // [4][S3_THREADS] low words, unsorted
  uint32_t *s_cnt = (uint32_t *)(s_low4 + 4 * S3_THREADS);                // [nbins_pad] counts (zero between chunks)
  uint32_t *s_off = s_cnt + P.nbins_pad;                                  // [nbins_pad] exclusive offsets
  uint32_t *s_dst = s_off + P.nbins_pad;                                  // [nbins_pad] entry index of sorted position 0 in the bin's run
  __shared__ __align__(16) uint32_t s_warp[S3_NWARPS];
  __shared__ unsigned long long s_chunk[2];
  __shared__ uint32_t s_total, s_over;
  // the chunk's 2-bit words and validity bits (incl. a 64-base halo), staged by TMA one chunk ahead
  // (with a single buffer the copy could only start after barrier (1) and its latency showed: +4 %)
  __shared__ __align__(16) unsigned long long s_seq[2][S3_SEQ_WORDS];
  __shared__ __align__(16) unsigned long long s_ok[2][S3_OK_WORDS];
  __shared__ __align__(8) unsigned long long s_bar[2];

  const uint32_t tid = threadIdx.x;
  const uint32_t s_cnt_sa = smem_u32(s_cnt);
  constexpr uint32_t k = KK;
  constexpr uint64_t kmask = (1ull << (2 * KK)) - 1ull;
  constexpr uint64_t kbits = (1ull << KK) - 1;
  constexpr uint64_t cmask = 0xaaaaaaaaaaaaaaaaull & kmask;
  constexpr uint64_t need = (1ull << (S3_NW + KK - 1)) - 1;               // validity bits behind a thread's 16 windows
  const uint32_t lmask = (1u << P.lbits) - 1u;
  const uint32_t xbits = P.lowbits - 32;
  const uint64_t nchunks = ceil_div_u64(P.tile_end - P.tile_begin, S3_TILES);
  const uint32_t bpt = P.nbins_pad / S3_THREADS;                          // bins per thread in the scan: 1 or 2
  unsigned long long *const rec_acc = P.acc + MFN_ACC_HEADER + MFN_DENSE_R * MFN_DENSE_A;
#ifdef SW_PROFILE
  long long prof_t = clock64();
#endif

  unsigned long long my_valid = 0;
  uint32_t cur_rec = 0, n_ok0 = 0;
  uint64_t cur_end = P.nseq ? P.rec_start[1] : ~0ull;
  auto flush_rec = [&]() {
    uint32_t a = n_ok0;
    for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
    if ((tid & 31) == 0 && cur_rec < P.nseq && a) atomicAdd(&rec_acc[2 * cur_rec], (unsigned long long)a);
    my_valid += n_ok0;
    n_ok0 = 0;
  };
  auto stage = [&](unsigned long long c, int buf) {
    s_chunk[buf] = c;
    if (c >= nchunks) return;
    const uint64_t t0 = P.tile_begin + c * S3_TILES;
    const uint64_t w2 = t0 * (MFN_TILE / 32), wo = t0 * (MFN_TILE / 64);
    const uint32_t n2 = (uint32_t)min((uint64_t)S3_SEQ_WORDS, P.seq_words - w2) & ~1u;
    const uint32_t no = (uint32_t)min((uint64_t)S3_OK_WORDS, P.ok_words - wo) & ~1u;
    mbar_expect_tx(&s_bar[buf], (n2 + no) * 8);
    tma_load_1d(s_seq[buf], P.seq2 + w2, n2 * 8, &s_bar[buf]);
    tma_load_1d(s_ok[buf], P.ok + wo, no * 8, &s_bar[buf]);
  };
  // side list: a window that cannot go to its bin (h = line * D + tag)
  auto to_side_list = [&](uint64_t h, uint32_t rec_flag) {
    unsigned long long o = atomicAdd(P.over_n, 1ull);
    if (o < P.over_cap) { P.over_lt[o] = h; P.over_rec[o] = rec_flag; }
    else atomicAdd(&P.acc[2], 1ull);                                      // reported as MFN_ERR_CAPACITY at fetch
  };

  unsigned long long pre = 0;
  for (uint32_t i = tid; i < P.nbins_pad; i += S3_THREADS) s_cnt[i] = 0;
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    stage(atomicAdd(P.counter, 1ull), 0);
    pre = atomicAdd(P.counter, 1ull);
  }
  __syncthreads();

  for (uint32_t it = 0;; it++) {
    const int buf = it & 1;
    const uint64_t chunk = s_chunk[buf];              // written before the previous round's barriers (or the one above)
    if (chunk >= nchunks) break;
    const uint64_t tile0 = P.tile_begin + chunk * S3_TILES;
    if (!DUMP) {
      const uint64_t gt = tile0 * MFN_TILE;
      if (gt >= cur_end) {
        flush_rec();
        while (gt >= cur_end) { cur_rec++; cur_end = (cur_rec < P.nseq) ? P.rec_start[cur_rec + 1] : ~0ull; }
      }
    }
    const uint32_t r0 = cur_rec;
    const uint64_t r0_end = cur_end;
    const bool one_record = DUMP || (tile0 + S3_TILES) * MFN_TILE <= r0_end;
    uint32_t *const where = DUMP ? P.where + (tile0 - P.tile_begin) * MFN_TILE : nullptr;    // this chunk's slice
    // the other staging buffer was last read before barrier (1) of the previous round
    if (tid == 0) { stage(pre, buf ^ 1); pre = atomicAdd(P.counter, 1ull); }
    mbar_wait(&s_bar[buf], (it >> 1) & 1);
    SW_PROF_MARK(0);

    // ---- hash this thread's 16 consecutive windows; rank them inside their bins ------------
    uint32_t meta[S3_NW];
    uint32_t live = 0;
    {
      const uint32_t j0 = tid * S3_NW;                                    // position in the staged chunk
      const uint64_t tile = tile0 + (j0 >> 11);
      uint64_t lo = 0, hi = 0, okb = 0;
      if (tile < P.tile_end) {
        const uint32_t w = j0 >> 5;
        const uint64_t a = s_seq[buf][w], b = s_seq[buf][w + 1];
        lo = (j0 & 16) ? ((a >> 32) | (b << 32)) : a;                     // j0 is a multiple of 16: the shift is 0 or 32 bits
        hi = (j0 & 16) ? (b >> 32) : b;
        const uint32_t wo = j0 >> 6, so = j0 & 63;
        const uint64_t c = s_ok[buf][wo], d = s_ok[buf][wo + 1];
        okb = so ? ((c >> so) | (d << (64 - so))) : c;
      }
      const uint32_t w0 = (uint32_t)lo, w1 = (uint32_t)(lo >> 32), w2 = (uint32_t)hi;
      const uint64_t F0 = mfn_rev2(lo & kmask) >> (64 - 2 * k);
      uint32_t Fl = (uint32_t)F0, Fh = (uint32_t)(F0 >> 32);
      uint32_t lows[S3_NW];
      // canonical k-mer of window u, hashed: (ch, cl); rolls the forward k-mer F by one base
      auto hash_window = [&](int u, uint32_t &ch, uint32_t &cl) {
        const uint32_t xl = u ? __funnelshift_r(w0, w1, 2 * u) : w0;
        const uint32_t xf = u ? __funnelshift_r(w1, w2, 2 * u) : w1;
        if (u) {
          const uint32_t nb = (xf >> (H::HB - 2)) & 3u;
          Fh = __funnelshift_l(Fl, Fh, 2) & H::HMASK;
          Fl = Fl * 4u + nb;
        }
        const uint32_t rl = xl ^ (uint32_t)cmask, rh = (xf & H::HMASK) ^ (uint32_t)(cmask >> 32);
        const bool lt = (((uint64_t)Fh << 32) | Fl) < (((uint64_t)rh << 32) | rl);
        cl = lt ? Fl : rl; ch = lt ? Fh : rh;
        H::mix(ch, cl);
      };
      // (ch, cl) -> low word, staging word without the rank (bin << 2 | low part bits 32..33)
      auto place = [&](uint32_t ch, uint32_t cl, uint32_t &low, uint32_t &m0) {
        if (GEO == 1) {
          low = cl; m0 = ch;                                             // 34-bit low part: the high half IS bin << 2 | bits 32..33
        } else if (GEO == 2) {
          low = cl; m0 = 2u * ch - (ch & 1u);                            // 33-bit low part: (ch >> 1) << 2 | (ch & 1)
        } else {
          uint64_t line, tag;
          joint_split(((uint64_t)ch << 32) | cl, P.J.fs_t, P.J.fs_d, P.J.fs_m, P.J.fs_s, line, tag);
          const uint64_t lp = ((uint64_t)((uint32_t)line & lmask) << P.tbits) | tag;
          low = (uint32_t)lp;
          m0 = (((uint32_t)line >> P.lbits) << 2) | (uint32_t)(lp >> 32);
        }
      };
      if (one_record && (okb & need) == need) {
        // the usual case: 16 valid windows of the chunk's record
        n_ok0 += S3_NW;
        live = 0xffffu;
#pragma unroll
        for (int u = 0; u < S3_NW; u++) {
          uint32_t ch, cl, m0;
          hash_window(u, ch, cl);
          place(ch, cl, lows[u], m0);
          const uint32_t rank = atomicAdd((uint32_t *)((char *)s_cnt + (m0 & S3_BIN_FIELD)), 1u);
          meta[u] = m0 | (rank << S3_RANK_SHIFT);
          SC_SCHED_FENCE(u);
        }
      } else {
        uint32_t vmask = 0;
#pragma unroll
        for (int u = 0; u < S3_NW; u++) if (((okb >> u) & kbits) == kbits) vmask |= 1u << u;
        // record of each window: 4 bits of delta per window here, bigger deltas take the side list
        uint64_t rdpack = 0;
        if (DUMP) {
        } else if (one_record || vmask == 0) n_ok0 += __popc(vmask);
        else {
          const uint64_t g0 = tile0 * MFN_TILE + j0;
          uint32_t rec = r0;
          uint64_t next_start = r0_end;
          for (int u = 0; u < S3_NW; u++) {
            if (!((vmask >> u) & 1u)) continue;
            const uint64_t g = g0 + u;
            while (g >= next_start) { rec++; next_start = (rec < P.nseq) ? P.rec_start[rec + 1] : ~0ull; }
            const uint32_t rd = rec - r0;
            if (rd == 0) { n_ok0++; continue; }
            atomicAdd(&rec_acc[2 * rec], 1ull);
            my_valid++;
            if (rd <= 15u) { rdpack |= (uint64_t)rd << (4 * u); continue; }
            // more records in this chunk than the staging word tells apart: side list
            vmask &= ~(1u << u);
            const uint64_t x = (u ? ((lo >> (2 * u)) | (hi << (64 - 2 * u))) : lo) & kmask;
            const uint64_t r = x ^ cmask, f = mfn_rev2(x) >> (64 - 2 * k);
            to_side_list(mfn_mix(f < r ? f : r, 0, kmask, k), rec);
          }
        }
        live = vmask;
#pragma unroll
        for (int u = 0; u < S3_NW; u++) {
          uint32_t ch, cl, m0;
          hash_window(u, ch, cl);
          place(ch, cl, lows[u], m0);
          const uint32_t rank = sc_rank_atomic(s_cnt_sa + (m0 & S3_BIN_FIELD), (vmask >> u) & 1u);
          meta[u] = m0 | (rank << S3_RANK_SHIFT) | (((uint32_t)(rdpack >> (4 * u)) & 15u) << S3_RD_SHIFT);
        }
      }
#pragma unroll
      for (int j = 0; j < S3_NW / 4; j++) s_low4[j * S3_THREADS + tid] = make_uint4(lows[4 * j], lows[4 * j + 1], lows[4 * j + 2], lows[4 * j + 3]);
    }
    SW_PROF_MARK(1);
    __syncthreads();                                                      // (1) the counts are complete
    SW_PROF_MARK(2);

    // ---- exclusive scan of the bin counts; reserve space in the global bins ----------------
    const uint32_t b0 = tid * bpt;
    uint32_t cnt0, cnt1 = 0;
    unsigned long long got0 = 0, got1 = 0;
    {
      cnt0 = s_cnt[b0];
      s_cnt[b0] = 0;
      if (bpt > 1) { cnt1 = s_cnt[b0 + 1]; s_cnt[b0 + 1] = 0; }
      // the reservations are in flight until the end of the sort
      if (cnt0) got0 = atomicAdd(&P.bin_count[(size_t)b0 * P.cstride], (unsigned long long)cnt0);
      if (cnt1) got1 = atomicAdd(&P.bin_count[(size_t)(b0 + 1) * P.cstride], (unsigned long long)cnt1);
      const uint32_t sum = cnt0 + cnt1;
      uint32_t incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= (uint32_t)o) incl += v; }
      if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
      if (tid == 0) s_over = 0;
      __syncthreads();                                                    // (2)
      const uint32_t w = tid >> 5;
      uint32_t wbase = 0;
#pragma unroll
      for (uint32_t x = 0; x < S3_NWARPS; x += 4) {
        const uint4 t4 = *(const uint4 *)&s_warp[x];
        wbase += w > x ? t4.x : 0u; wbase += w > x + 1 ? t4.y : 0u; wbase += w > x + 2 ? t4.z : 0u; wbase += w > x + 3 ? t4.w : 0u;
      }
      const uint32_t run = wbase + incl - sum;
      s_off[b0] = run;
      if (bpt > 1) s_off[b0 + 1] = run + cnt0;
      if (tid == S3_THREADS - 1) s_total = run + sum;
    }
    __syncthreads();                                                      // (3) the offsets are ready
    SW_PROF_MARK(3);

    // ---- counting sort into shared memory (the reservations are still in flight) -----------
    {
      uint32_t lows[S3_NW];
#pragma unroll
      for (int j = 0; j < S3_NW / 4; j++) {
        const uint4 v = s_low4[j * S3_THREADS + tid];
        lows[4 * j] = v.x; lows[4 * j + 1] = v.y; lows[4 * j + 2] = v.z; lows[4 * j + 3] = v.w;
      }
      // all offset loads of a half before its stores: the compiler cannot tell that s_ent and s_off do
      // not overlap, and would otherwise wait for every LDS before the next STS
#pragma unroll
      for (int h = 0; h < S3_NW; h += 8) {
        uint32_t pos[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const uint32_t m = meta[h + i];
          pos[i] = *(const uint32_t *)((const char *)s_off + (m & S3_BIN_FIELD)) + (m >> S3_RANK_SHIFT);
        }
        // staged high word: bin << 2 | low-part bits 32..33, and above them the record delta -- or, DUMP, the
        // window's index in the chunk
        auto staged = [&](int idx) -> unsigned long long {
          const uint32_t hw = DUMP ? ((meta[idx] & 0x1fffu) | ((tid * S3_NW + (uint32_t)idx) << 13)) : (meta[idx] & S3_KEEP_MASK);
          return ((unsigned long long)hw << 32) | lows[idx];
        };
        if (live == 0xffffu) {
#pragma unroll
          for (int i = 0; i < 8; i++) s_ent[pos[i]] = staged(h + i);
        } else {
#pragma unroll
          for (int i = 0; i < 8; i++)
            if ((live >> (h + i)) & 1u) s_ent[pos[i]] = staged(h + i);
        }
      }
    }
    // entry index (in P.entries) of the chunk's sorted position 0 in each bin's run; the sums stay below 2^32
    // (checked on the host), intermediate values may wrap
    {
      const uint32_t cap = (uint32_t)P.cap;
      s_dst[b0] = b0 * cap + (uint32_t)got0 - s_off[b0];
      if (bpt > 1) s_dst[b0 + 1] = (b0 + 1) * cap + (uint32_t)got1 - s_off[b0 + 1];
      if ((cnt0 && got0 + cnt0 > P.cap) || (cnt1 && got1 + cnt1 > P.cap)) s_over = 1;       // a run crosses the end of its bin
      if (DUMP) {
        // where every bin's run of this chunk starts and how long it is: the gather pass reads whole runs
        uint2 *const runs = P.runs + chunk * P.nbins_pad;
        runs[b0] = make_uint2(b0 * cap + (uint32_t)got0, cnt0);
        if (bpt > 1) runs[b0 + 1] = make_uint2((b0 + 1) * cap + (uint32_t)got1, cnt1);
      }
    }
    SW_PROF_MARK(4);
    __syncthreads();                                                      // (4) sorted entries and destinations are ready
    SW_PROF_MARK(5);

    // ---- coalesced copy-out: consecutive threads write consecutive entries of a bin's run --
    {
      const uint32_t total = s_total;
      const uint32_t r0s = r0 << xbits;
      unsigned long long *const ent = P.entries;
      if (DUMP) {
        // entry = low part only; where[window] = bin << 13 | rank of the window in the bin's run of this chunk, staged
        // in the low-word area (dead since the sort) so that it reaches global memory as 16-byte stores.  A run
        // that does not fit its bin makes the whole call fall back to the direct scan (over_n != 0 is checked by
        // the host before the probe pass).
        uint32_t *const s_where = (uint32_t *)s_low4;
        if (live != 0xffffu) {
#pragma unroll
          for (int u = 0; u < (int)S3_NW; u++) if (!((live >> u) & 1u)) s_where[tid * S3_NW + u] = 0xffffffffu;
        }
        for (uint32_t i = tid; i < total; i += S3_THREADS) {
          const unsigned long long e = s_ent[i];
          const uint32_t hs = (uint32_t)(e >> 32);
          const uint32_t bin = (hs & S3_BIN_FIELD) >> 2;
          const uint32_t d = *(const uint32_t *)((const char *)s_dst + (hs & S3_BIN_FIELD)) + i;
          const uint32_t rank = i - *(const uint32_t *)((const char *)s_off + (hs & S3_BIN_FIELD));
          if ((unsigned long long)d < (unsigned long long)(bin + 1) * P.cap) *(uint2 *)(ent + d) = make_uint2((uint32_t)e, hs & 3u);
          else atomicAdd(P.over_n, 1ull);
          s_where[hs >> 13] = (bin << 13) | rank;
        }
        __syncthreads();                                                  // (5, DUMP only)
#pragma unroll
        for (uint32_t i = tid; i < S3_CHUNK / 4; i += S3_THREADS) ((uint4 *)where)[i] = ((const uint4 *)s_where)[i];
        __syncthreads();                                                  // (6, DUMP only) the next round's low words may be written
      } else if (one_record && !s_over) {
        // the entry's high word: the chunk's record next to the staged bits 32..33 of the low part
        if (total == S3_CHUNK) {
#pragma unroll
          for (int j = 0; j < S3_NW; j++) {
            const uint32_t i = tid + j * S3_THREADS;
            const unsigned long long e = s_ent[i];
            const uint32_t hs = (uint32_t)(e >> 32);
            const uint32_t d = *(const uint32_t *)((const char *)s_dst + (hs & S3_BIN_FIELD)) + i;
            *(uint2 *)(ent + d) = make_uint2((uint32_t)e, (hs & 3u) | r0s);
          }
        } else {
#pragma unroll 4
          for (uint32_t i = tid; i < total; i += S3_THREADS) {
            const unsigned long long e = s_ent[i];
            const uint32_t hs = (uint32_t)(e >> 32);
            const uint32_t d = *(const uint32_t *)((const char *)s_dst + (hs & S3_BIN_FIELD)) + i;
            *(uint2 *)(ent + d) = make_uint2((uint32_t)e, (hs & 3u) | r0s);
          }
        }
      } else {
        for (uint32_t i = tid; i < total; i += S3_THREADS) {
          const unsigned long long e = s_ent[i];
          const uint32_t hs = (uint32_t)(e >> 32);
          const uint32_t bin = (hs & S3_BIN_FIELD) >> 2;
          const uint32_t d = s_dst[bin] + i;
          const uint32_t rec = r0 + ((hs >> S3_RD_SHIFT) & S3_RD_MASK);
          const unsigned long long lowpart = (unsigned long long)(uint32_t)e | ((unsigned long long)(hs & 3u) << 32);
          if ((unsigned long long)d < (unsigned long long)(bin + 1) * P.cap) {
            ent[d] = lowpart | ((unsigned long long)rec << P.lowbits);
          } else {
            const uint64_t line = ((uint64_t)bin << P.lbits) | (lowpart >> P.tbits);
            to_side_list(line * P.J.D + (lowpart & ((1ull << P.tbits) - 1ull)), rec);
          }
        }
      }
    }
    SW_PROF_MARK(6);
  }
  if (DUMP) return;
  flush_rec();
  for (int o = 16; o > 0; o >>= 1) my_valid += __shfl_down_sync(0xffffffffu, my_valid, o);
  if ((tid & 31) == 0 && my_valid) atomicAdd(&P.acc[0], my_valid);
}
