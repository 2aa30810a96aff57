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
//
// Entry (8 bytes):  bits 0..lowbits-1   (line in bin) << tbits | tag
//                   bits lowbits..62   record id
//                   bit  63            the k-mer is its own reverse complement: counts twice (Q2)
// lowbits = 32 ("narrow") whenever line-in-bin and tag fit 32 bits (k = 21 at any table size);
// wide tags (k = 31: 34-46 bits) use lowbits = lbits + tbits <= 52 and leave fewer bits for the
// record id (the scan falls back to k_scan when the records do not fit).
#pragma once
#include "scan_kernel.cuh"

#define SW_THREADS 256
// The scatter CTA's shape can be overridden at compile time for experiments (see `make variants`
// in the Makefile; MFN_LIB selects the library at run time).
#ifndef SW_CHUNK_TILES
#define SW_CHUNK_TILES 4
#endif
#define SW_CHUNK (SW_CHUNK_TILES * MFN_TILE)      // 8192 windows per CTA round
#ifndef SC_THREADS
#define SC_THREADS 512                             // scatter CTA: two 2048-window tiles side by side, two rounds
#endif
#ifndef SC_CTAS_PER_SM
#define SC_CTAS_PER_SM (SC_THREADS > 512 ? 1 : (1024 / SC_THREADS))
#endif
// staging word of a window ("meta", 32 bits), laid out so that the sort and the copy-out need the
// fewest ALU operations (the pass is bound by the ALU pipe):
//   bits 0..1    bits 32..33 of the entry's low part (narrow entries wider than 32 bits)
//   bits 2..12   bin: meta & 0x1ffc is the byte offset of the bin's 32-bit counter
//   bits 13..    record delta against the chunk's first record
//   next bit     the k-mer counts twice
//   top bits     rank of the window inside its bin (< SW_CHUNK): meta >> SC_RANK_SHIFT
#define SC_BIN_SHIFT 2
#define SC_BIN_BITS 11
#define SC_BIN_FIELD (((1u << SC_BIN_BITS) - 1u) << SC_BIN_SHIFT)
#define SC_RANK_BITS (SW_CHUNK_TILES <= 1 ? 11 : SW_CHUNK_TILES <= 2 ? 12 : SW_CHUNK_TILES <= 4 ? 13 : 14)
#define SC_RANK_SHIFT (32 - SC_RANK_BITS)
#define SC_TWICE_BIT (SC_RANK_SHIFT - 1)
#define SC_RD_SHIFT (SC_BIN_SHIFT + SC_BIN_BITS)
#define SC_RD_BITS (SC_TWICE_BIT - SC_RD_SHIFT)
#define SC_RD_MASK ((1u << SC_RD_BITS) - 1u)
#define SC_KEEP_MASK ((1u << SC_RANK_SHIFT) - 1u)                    // everything but the rank
static_assert(SW_CHUNK <= (1 << SC_RANK_BITS) && SW_CHUNK % (SC_THREADS * WPT) == 0 && SC_THREADS % (MFN_TILE / WPT) == 0, "scatter CTA shape");
static_assert(SC_RD_BITS >= 3 && SC_RD_BITS <= 7, "record delta field");
#define SC_TPT (SW_CHUNK / SC_THREADS / WPT)       // 8-window groups per thread per chunk
#define SC_NW (SC_TPT * WPT)                       // windows per thread per chunk
#define SC_SEQ_WORDS (SW_CHUNK / 32 + 2)           // staged 2-bit words of a chunk (2064 B)
#define SC_OK_WORDS (SW_CHUNK / 64 + 2)            // staged validity words (1040 B)
#define SW_MAX_BINS 2048
#define SW_MAX_LOWBITS_NARROW 34                   // low part of a narrow entry: 32 bits in a register + 2 in the staging word
#ifndef SW_SLICE
#define SW_SLICE 1024                              // entries per probe work item (one WARP)
#endif
#define SW_QCAP 160                                // per-warp queue of slow lookups: < 32 carried + 128 per batch
#define SW_NOSPACE 0xffffffffu
#define SW_MAX_RD (SC_RD_MASK - 1u)                // record delta inside one chunk that fits the staging word

struct SweepParams {
  JointDev J;
  DevTable R, A;                             // restricted joint table only: where a k-mer it does not hold is looked up
  uint32_t restricted;
  const unsigned long long *seq2, *ok, *rec_start;
  uint64_t seq_words, ok_words;              // allocated length of seq2 / ok (TMA copies are clamped to it)
  uint32_t nseq;
  uint64_t tile_begin, tile_end;
  uint32_t lbits, tbits, nbins, nbins_pad;   // bin = line >> lbits; entry low part = (line & L-1) << tbits | tag
  uint32_t lowbits;                          // width of the low part: 32..34 (narrow) or lbits + tbits (wide)
  unsigned long long *entries;               // bin b: entries[b * cap .. b * cap + n_b)
  uint64_t cap;
  unsigned long long *bin_count;             // [nbins_pad * cstride] reserved entries of bin b at [b * cstride] (may exceed cap)
  uint32_t cstride;                          // words between two bins' counters (spreads the reservation atomics over L2 lines)
  unsigned long long *bin_valid;             // [nbins_pad] first reservation that did not fit (~0: none)
  unsigned long long *over_lt;               // windows that found no room in their bin: h = line * D + tag ...
  uint32_t *over_rec;                        //   ... and record id (bit 31: counts twice)
  unsigned long long *over_n;
  uint64_t over_cap;
  unsigned long long *counter;               // work distribution (one per pass)
  unsigned long long *acc;
  PairTable sparse;
  const unsigned long long *zmap;
  uint32_t zlen;
  uint32_t zhi;                              // largest multiplicity v >= 1 with readK(v) == 0 under -prob (0: none)
  uint32_t pr_lim, pr_wmask;                 // 32-bit slots: 2^vb - 1 and the tag field mask (vb = value bits)
  uint32_t *where;                           // -dump flavour: bin << 13 | rank in the bin's run of the chunk, per window (0xffffffff: none)
  uint2 *runs;                               // -dump flavour: [chunk][nbins_pad] (entry index of the run's start, length)
  unsigned long long *prof;                  // SW_PROFILE builds: [8] cycles per phase of the scatter pass, summed over CTAs
};

// keeps the compiler from hashing all eight windows of a group before the first atomic (which needs
// more live registers than there are): a scheduling fence every SC_FENCE_EVERY windows
#ifndef SC_FENCE_EVERY
#define SC_FENCE_EVERY 2
#endif
#ifndef SC_FENCE_KIND
#define SC_FENCE_KIND 1
#endif
#if SC_FENCE_KIND == 0
#define SC_SCHED_FENCE(u) do { } while (0)
#elif SC_FENCE_KIND == 1
#define SC_SCHED_FENCE(u) do { if (((u) + 1) % SC_FENCE_EVERY == 0) asm volatile("" ::: "memory"); } while (0)
#else
#define SC_SCHED_FENCE(u) do { if (((u) + 1) % SC_FENCE_EVERY == 0) __syncwarp(); } while (0)
#endif

// rank = live ? atomicAdd(shared counter at byte address saddr, 1) : 0, as ONE predicated instruction
// (a C++ `if` around the atomic becomes a branch per window and splits the straight-line hash code)
__device__ __forceinline__ uint32_t sc_rank_atomic(uint32_t saddr, uint32_t live) {
  uint32_t rank;
  asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %2, 0;\n mov.u32 %0, 0;\n @p atom.shared.add.u32 %0, [%1], 1;\n}"
               : "=r"(rank) : "r"(saddr), "r"(live) : "memory");
  return rank;
}

#ifdef SW_PROFILE
#define SW_PROF_MARK(i) do { if (tid == 0) { const long long t_ = clock64(); atomicAdd(&P.prof[i], (unsigned long long)(t_ - prof_t)); prof_t = t_; } } while (0)
#else
#define SW_PROF_MARK(i) do { } while (0)
#endif

// ---- the k-mer as two 32-bit halves (16 < k < 32): every 64-bit logic op or shift is two ALU
// instructions, and half of those would work on bits that are known to be zero ---------------
template <int K> struct Kmer2 {
  static constexpr uint32_t HB = 2 * K - 32;                    // bits in the high half
  static constexpr uint32_t HMASK = (1u << HB) - 1u;
  // x ^= x >> K: the high half has fewer than K bits, so only the low half changes
  static __device__ __forceinline__ void xorshift(uint32_t hi, uint32_t &lo) { lo ^= __funnelshift_r(lo, hi, K); }
  // x = (x * M) mod 2^2K
  static __device__ __forceinline__ void mul(uint32_t &hi, uint32_t &lo, uint64_t M) {
    const uint32_t ml = (uint32_t)M, mh = (uint32_t)(M >> 32);
    const uint64_t p = (uint64_t)lo * ml;
    hi = ((uint32_t)(p >> 32) + lo * mh + hi * ml) & HMASK;
    lo = (uint32_t)p;
  }
  // mfn_mix(c, 0, kmask, K) on halves
  static __device__ __forceinline__ void mix(uint32_t &hi, uint32_t &lo) {
    xorshift(hi, lo);
    mul(hi, lo, 0xff51afd7ed558ccdull);
    xorshift(hi, lo);
    mul(hi, lo, 0xc4ceb9fe1a85ec53ull);
    xorshift(hi, lo);
  }
};

// ------------------------------------------------------------------------------------------
// pass 1.  KK: mer size known at compile time (0 = read it from the table)
//          WIDE: 64-bit low part (tables with 64-bit slots whose tags do not fit a narrow entry)
//          GEO: 0 = any table divisor D = d << t, narrow low part of 32..34 bits
//               1 = D is a power of two and the low part is exactly 34 bits: line, tag, bin and low
//                   part are bit fields of the hash, no division and no re-assembly
// ------------------------------------------------------------------------------------------
template <int KK, bool WIDE, int GEO>
__global__ void __launch_bounds__(SC_THREADS, WIDE ? 1 : SC_CTAS_PER_SM) k_sweep_scatter(const __grid_constant__ SweepParams P) {
  typedef typename std::conditional<WIDE, unsigned long long, uint32_t>::type low_t;
  constexpr bool HALVES = (KK > 16 && KK < 32) && !WIDE;              // hash on 32-bit halves
  static_assert(GEO == 0 || HALVES, "the bit-field geometry is implemented for 16 < k < 32");
  extern __shared__ __align__(16) unsigned char sw_smem[];            // dynamic: SW_CHUNK * 8 + nbins_pad * 12
  unsigned long long *s_ent = (unsigned long long *)sw_smem;           // [SW_CHUNK] entries sorted by bin
  unsigned long long *s_dst = s_ent + SW_CHUNK;                        // [nbins_pad] where the chunk's sorted position 0 would go, per bin (byte address)
  uint32_t *s_off = (uint32_t *)(s_dst + P.nbins_pad);                 // [nbins_pad] counts, then exclusive offsets
  __shared__ uint32_t s_warp[SC_THREADS / 32];
  __shared__ unsigned long long s_chunk[2];
  __shared__ uint32_t s_total;
  // the chunk's 2-bit words and validity bits (incl. a 64-base halo), staged by TMA one chunk ahead
  __shared__ __align__(16) unsigned long long s_seq[2][SC_SEQ_WORDS];
  __shared__ __align__(16) unsigned long long s_ok[2][SC_OK_WORDS];
  __shared__ __align__(8) unsigned long long s_bar[2];

  const uint32_t tid = threadIdx.x;
  const uint32_t s_off_sa = smem_u32(s_off);
  const uint32_t grp = tid / (MFN_TILE / WPT), t8 = tid % (MFN_TILE / WPT);   // tile of the round, window group in it
  const uint32_t k = KK ? (uint32_t)KK : P.J.k;
  const uint64_t kmask = KK ? (KK == 32 ? ~0ull : ((1ull << (2 * (KK & 31))) - 1ull)) : P.J.kmask;
  const uint64_t kbits = (1ull << k) - 1;
  const uint64_t cmask = 0xaaaaaaaaaaaaaaaaull & kmask;
  const uint64_t need = (1ull << (WPT + k - 1)) - 1;                   // validity bits behind a thread's 8 windows
  const bool even_k = (k & 1u) == 0;
  const uint32_t lmask = (1u << P.lbits) - 1u;
  const uint32_t xbits = P.lowbits > 32 && !WIDE ? P.lowbits - 32 : 0; // bits of a narrow low part beyond 32
  const uint64_t nchunks = ceil_div_u64(P.tile_end - P.tile_begin, SW_CHUNK_TILES);
  const uint32_t per_thread_bins = P.nbins_pad / SC_THREADS;
  unsigned long long *const rec_acc = P.acc + MFN_ACC_HEADER + MFN_DENSE_R * MFN_DENSE_A;
#ifdef SW_PROFILE
  long long prof_t = clock64();
#endif

  unsigned long long my_valid = 0;
  uint32_t cur_rec = 0, n_ok0 = 0;                           // record of the current chunk's first window
  uint64_t cur_end = P.nseq ? P.rec_start[1] : ~0ull;        // chunks arrive in increasing order: the record only moves forward
  auto flush_rec = [&]() {
    uint32_t a = n_ok0;
    for (int o = 16; o > 0; o >>= 1) a += __shfl_down_sync(0xffffffffu, a, o);
    if ((tid & 31) == 0 && cur_rec < P.nseq && a) atomicAdd(&rec_acc[2 * cur_rec], (unsigned long long)a);
    my_valid += n_ok0;
    n_ok0 = 0;
  };

  // thread 0 claims chunks two rounds ahead and starts the TMA of the next chunk's sequence one
  // round ahead, so neither the atomic's round trip nor the loads sit on a round's critical path
  auto stage = [&](unsigned long long c, int buf) {
    s_chunk[buf] = c;
    if (c >= nchunks) return;
    const uint64_t t0 = P.tile_begin + c * SW_CHUNK_TILES;
    const uint64_t w2 = t0 * (MFN_TILE / 32), wo = t0 * (MFN_TILE / 64);
    const uint32_t n2 = (uint32_t)min((uint64_t)SC_SEQ_WORDS, P.seq_words - w2) & ~1u;     // multiples of 16 bytes
    const uint32_t no = (uint32_t)min((uint64_t)SC_OK_WORDS, P.ok_words - wo) & ~1u;
    mbar_expect_tx(&s_bar[buf], (n2 + no) * 8);
    tma_load_1d(s_seq[buf], P.seq2 + w2, n2 * 8, &s_bar[buf]);
    tma_load_1d(s_ok[buf], P.ok + wo, no * 8, &s_bar[buf]);
  };
  unsigned long long pre = 0;
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    stage(atomicAdd(P.counter, 1ull), 0);
    pre = atomicAdd(P.counter, 1ull);
  }

  for (uint32_t it = 0;; it++) {
    const int buf = it & 1;
    for (uint32_t i = tid; i < P.nbins_pad; i += SC_THREADS) s_off[i] = 0;
    __syncthreads();                                         // s_chunk[buf], the zeroed counters; everyone is done with the other buffer
    const uint64_t chunk = s_chunk[buf];
    if (chunk >= nchunks) break;
    if (tid == 0) { stage(pre, buf ^ 1); pre = atomicAdd(P.counter, 1ull); }
    const uint64_t tile0 = P.tile_begin + chunk * SW_CHUNK_TILES;
    SW_PROF_MARK(0);

    // record of the chunk's first window (uniform); does the chunk stay inside it?
    {
      const uint64_t gt = tile0 * MFN_TILE;
      if (gt >= cur_end) {
        flush_rec();
        while (gt >= cur_end) { cur_rec++; cur_end = (cur_rec < P.nseq) ? P.rec_start[cur_rec + 1] : ~0ull; }
      }
    }
    const uint32_t r0 = cur_rec;
    const uint64_t r0_end = cur_end;
    mbar_wait(&s_bar[buf], (it >> 1) & 1);
    const bool one_record = (tile0 + SW_CHUNK_TILES) * MFN_TILE <= r0_end;

    // ---- hash this thread's windows; rank them inside their bins ---------------------------
    low_t low[SC_NW];
    uint32_t meta[SC_NW];
    uint32_t live = 0;                                                // bit idx: window idx goes to a bin
#pragma unroll
    for (int q = 0; q < SC_TPT; q++) {
      const uint64_t tile = tile0 + q * (SC_THREADS / (MFN_TILE / WPT)) + grp;
      const uint64_t g0 = tile * MFN_TILE + t8 * WPT;
      uint64_t lo = 0, hi = 0, okb = 0;
      if (tile < P.tile_end) {
        const uint32_t j0 = (uint32_t)(tile - tile0) * MFN_TILE + t8 * WPT;      // position in the staged chunk
        const uint32_t w = j0 >> 5, sh = (j0 & 31) * 2;
        const uint64_t a = s_seq[buf][w], b = s_seq[buf][w + 1];
        lo = sh ? ((a >> sh) | (b << (64 - sh))) : a;
        hi = b >> sh;
        const uint32_t wo = j0 >> 6, so = j0 & 63;
        const uint64_t c = s_ok[buf][wo], d = s_ok[buf][wo + 1];
        okb = so ? ((c >> so) | (d << (64 - so))) : c;
      }
      // which of the 8 windows are valid (kmers-iterator.H:93-95): all of them almost always
      uint32_t vmask = 0xffu;
      if ((okb & need) != need) {
        vmask = 0;
#pragma unroll
        for (int u = 0; u < WPT; u++) if (((okb >> u) & kbits) == kbits) vmask |= 1u << u;
      }
      // record of each window: the chunk's first record almost always (delta 0)
      uint64_t rdpack = 0;                                       // 7 bits per window
      if (one_record || vmask == 0) n_ok0 += __popc(vmask);
      else {
        uint32_t rec = r0;
        uint64_t next_start = r0_end;
        for (int u = 0; u < WPT; u++) {
          if (!((vmask >> u) & 1u)) continue;
          const uint64_t g = g0 + u;
          while (g >= next_start) { rec++; next_start = (rec < P.nseq) ? P.rec_start[rec + 1] : ~0ull; }
          const uint32_t rd = rec - r0;
          if (rd == 0) { n_ok0++; continue; }
          atomicAdd(&rec_acc[2 * rec], 1ull);
          my_valid++;
          if (!WIDE && rd <= SW_MAX_RD) { rdpack |= (uint64_t)rd << (7 * u); continue; }
          // a chunk spanning more records than the staging word can tell apart (wide entries: any
          // record but the chunk's first): this window goes to the side list
          vmask &= ~(1u << u);
          const uint64_t x = (u ? ((lo >> (2 * u)) | (hi << (64 - 2 * u))) : lo) & kmask;
          const uint64_t r = x ^ cmask, f = mfn_rev2(x) >> (64 - 2 * k);
          uint64_t line, tag;
          joint_split(mfn_mix(f < r ? f : r, 0, kmask, k), P.J.fs_t, P.J.fs_d, P.J.fs_m, P.J.fs_s, line, tag);
          unsigned long long o = atomicAdd(P.over_n, 1ull);
          if (o < P.over_cap) { P.over_lt[o] = line * P.J.D + tag; P.over_rec[o] = rec | ((even_k && f == r) ? 0x80000000u : 0u); }
          else atomicAdd(&P.acc[2], 1ull);
        }
      }
      live |= vmask << (q * WPT);
      // Forward k-mer F (first base most significant, kmers-tiny.H:82): bit-reversal for the first
      // window, then one base shifted in per window; the stream-order slice x (base g in the low 2
      // bits) gives the reverse complement for free (x ^ cmask).  Straight-line code: windows that
      // are not valid are hashed like the others and only their shared-memory atomic is predicated off.
      if (HALVES) {
        typedef Kmer2<KK ? KK : 21> H;
        const uint32_t w0 = (uint32_t)lo, w1 = (uint32_t)(lo >> 32), w2 = (uint32_t)hi;
        const uint64_t F0 = mfn_rev2(lo & kmask) >> (64 - 2 * k);
        uint32_t Fl = (uint32_t)F0, Fh = (uint32_t)(F0 >> 32);
#pragma unroll
        for (int u = 0; u < WPT; u++) {
          const int idx = q * WPT + u;
          const uint32_t xl = u ? __funnelshift_r(w0, w1, 2 * u) : w0;
          const uint32_t xf = u ? __funnelshift_r(w1, w2, 2 * u) : w1;           // high half before masking
          if (u) {
            const uint32_t nb = (xf >> (H::HB - 2)) & 3u;                        // the window's last base
            Fh = __funnelshift_l(Fl, Fh, 2) & H::HMASK;
            Fl = Fl * 4u + nb;
          }
          const uint32_t rl = xl ^ (uint32_t)cmask, rh = (xf & H::HMASK) ^ (uint32_t)(cmask >> 32);
          const bool lt = Fh < rh || (Fh == rh && Fl < rl);
          uint32_t cl = lt ? Fl : rl, ch = lt ? Fh : rh;
          const uint32_t twice = (even_k && Fh == rh && Fl == rl) ? (1u << SC_TWICE_BIT) : 0u;
          H::mix(ch, cl);
          uint32_t binoff, m0;                                                   // byte offset of the bin's counter; meta without the rank
          if (GEO == 1) {
            // 34-bit low part = the low 34 bits of the hash; bin = the bits above: the high half IS
            // bits 32..33 of the low part | bin << 2
            low[idx] = (low_t)cl;
            m0 = ch;
            binoff = ch & ~3u;
          } else {
            uint64_t line, tag;
            joint_split(((uint64_t)ch << 32) | cl, P.J.fs_t, P.J.fs_d, P.J.fs_m, P.J.fs_s, line, tag);
            const uint64_t lp = ((uint64_t)((uint32_t)line & lmask) << P.tbits) | tag;
            low[idx] = (low_t)lp;
            binoff = ((uint32_t)line >> P.lbits) << SC_BIN_SHIFT;
            m0 = binoff | (uint32_t)(lp >> 32);
          }
          const uint32_t rank = sc_rank_atomic(s_off_sa + binoff, (vmask >> u) & 1u);
          meta[idx] = m0 | (rank << SC_RANK_SHIFT) | (((uint32_t)(rdpack >> (7 * u)) & 0x7fu) << SC_RD_SHIFT) | twice;
          SC_SCHED_FENCE(u);
        }
      } else {
        uint64_t F = mfn_rev2(lo & kmask) >> (64 - 2 * k);
#pragma unroll
        for (int u = 0; u < WPT; u++) {
          const int idx = q * WPT + u;
          uint64_t x = u ? ((lo >> (2 * u)) | (hi << (64 - 2 * u))) : lo;
          x &= kmask;                                             // base g in the low 2 bits
          if (u) F = ((F << 2) | (x >> (2 * k - 2))) & kmask;
          const uint64_t r = x ^ cmask;
          const uint32_t twice = (even_k && F == r) ? (1u << SC_TWICE_BIT) : 0u;
          uint64_t line, tag;
          joint_split(mfn_mix(F < r ? F : r, 0, kmask, k), P.J.fs_t, P.J.fs_d, P.J.fs_m, P.J.fs_s, line, tag);
          const uint32_t bin = (uint32_t)line >> P.lbits;
          const uint32_t rank = sc_rank_atomic(s_off_sa + bin * 4u, (vmask >> u) & 1u);
          const uint64_t lp = (((uint64_t)((uint32_t)line & lmask)) << P.tbits) | tag;
          low[idx] = (low_t)lp;
          meta[idx] = (bin << SC_BIN_SHIFT) | (rank << SC_RANK_SHIFT) | (((uint32_t)(rdpack >> (7 * u)) & 0x7fu) << SC_RD_SHIFT) | twice |
                      (WIDE ? 0u : (uint32_t)(lp >> 32));
        }
      }
    }
    SW_PROF_MARK(1);
    __syncthreads();
    SW_PROF_MARK(2);

    // ---- exclusive scan of the bin counts; reserve space in the global bins ----------------
    constexpr uint32_t BPT = SW_MAX_BINS / SC_THREADS;
    const uint32_t b0 = tid * per_thread_bins;
    uint32_t cnt[BPT];
    unsigned long long got[BPT];
    {
      uint32_t sum = 0;
#pragma unroll
      for (uint32_t i = 0; i < BPT; i++) { cnt[i] = i < per_thread_bins ? s_off[b0 + i] : 0; sum += cnt[i]; }
      // all atomics of this thread in flight together; their results are first needed after the sort
#pragma unroll
      for (uint32_t i = 0; i < BPT; i++) {
        got[i] = 0;
        if (cnt[i]) got[i] = atomicAdd(&P.bin_count[(size_t)(b0 + i) * P.cstride], (unsigned long long)cnt[i]);
      }
      uint32_t incl = sum;
      for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= (uint32_t)o) incl += v; }
      if ((tid & 31) == 31) s_warp[tid >> 5] = incl;
      __syncthreads();
      uint32_t wbase = 0;
      for (uint32_t w = 0; w < (tid >> 5); w++) wbase += s_warp[w];
      uint32_t run = wbase + incl - sum;
#pragma unroll
      for (uint32_t i = 0; i < BPT; i++) {
        if (i < per_thread_bins) { s_off[b0 + i] = run; run += cnt[i]; }
      }
      if (tid == SC_THREADS - 1) s_total = run;
    }
    __syncthreads();
    SW_PROF_MARK(3);

    // ---- counting sort into shared memory (the reservations are still in flight) -----------
#pragma unroll
    for (int idx = 0; idx < SC_NW; idx++) {
      const uint32_t m = meta[idx];
      const uint32_t p = *(const uint32_t *)((const char *)s_off + (m & SC_BIN_FIELD)) + (m >> SC_RANK_SHIFT);
      // staged: narrow = meta without the rank << 32 | low; wide = twice (63) | bin (52..62) | low (0..51)
      if ((live >> idx) & 1u) {
        if (WIDE) s_ent[p] = ((unsigned long long)((m >> SC_TWICE_BIT) & 1u) << 63) | ((unsigned long long)((m & SC_BIN_FIELD) >> SC_BIN_SHIFT) << 52) | low[idx];
        else s_ent[p] = ((unsigned long long)(m & SC_KEEP_MASK) << 32) | low[idx];
      }
    }
    // where each bin's run goes: byte address of the chunk's sorted position 0 in the bin's run
    // (entries + (bin * cap + reservation - offset) * 8; the address of a live entry never has bit 63)
#pragma unroll
    for (uint32_t i = 0; i < BPT; i++) {
      if (i < per_thread_bins) {
        unsigned long long dst = (unsigned long long)(P.entries + ((unsigned long long)(b0 + i) * P.cap + got[i] - s_off[b0 + i])) & 0x7fffffffffffffffull;
        // bit 63: the run does not fit its bin (its windows go to the side list)
        if (cnt[i] && got[i] + cnt[i] > P.cap) { atomicMin(&P.bin_valid[b0 + i], got[i]); dst |= 1ull << 63; }
        s_dst[b0 + i] = dst;
      }
    }
    SW_PROF_MARK(4);
    __syncthreads();
    SW_PROF_MARK(5);

    // ---- coalesced copy-out: consecutive threads write consecutive entries of a bin's run --
    {
      const uint32_t total = s_total;
      const uint32_t r0s = r0 << xbits;
      // the usual chunk (one record, k-mers cannot be their own reverse complement): the entry's high
      // word is the chunk's record next to the staged bits 32..33 of the low part
      const bool simple = !WIDE && one_record && !even_k;
#pragma unroll 4
      for (uint32_t i = tid; i < total; i += SC_THREADS) {
        const unsigned long long e = s_ent[i];
        const uint32_t hs = (uint32_t)(e >> 32);
        const unsigned long long dst = WIDE ? s_dst[(uint32_t)(e >> 52) & 0x7ffu]
                                            : *(const unsigned long long *)((const char *)s_dst + ((hs & SC_BIN_FIELD) << 1));
        if (simple && (long long)dst >= 0) {
          *(uint2 *)(dst + (unsigned long long)i * 8) = make_uint2((uint32_t)e, (hs & 3u) | r0s);
          continue;
        }
        uint32_t bin, rec;
        unsigned long long lowpart, out;
        if (WIDE) {
          bin = (uint32_t)(e >> 52) & 0x7ffu;
          lowpart = e & 0x000fffffffffffffull;
          rec = r0 | ((uint32_t)(e >> 32) & 0x80000000u);
          out = lowpart | ((unsigned long long)r0 << P.lowbits) | (e & 0x8000000000000000ull);
        } else {
          bin = (hs & SC_BIN_FIELD) >> SC_BIN_SHIFT;
          const uint32_t tw = ((hs >> SC_TWICE_BIT) & 1u) << 31;
          rec = (r0 + ((hs >> SC_RD_SHIFT) & SC_RD_MASK)) | tw;
          lowpart = (unsigned long long)(uint32_t)e | ((unsigned long long)(hs & 3u) << 32);
          out = lowpart | ((unsigned long long)(rec & 0x7fffffffu) << P.lowbits) | ((unsigned long long)tw << 32);
        }
        if ((long long)dst >= 0) {
          *(unsigned long long *)(dst + (unsigned long long)i * 8) = out;
        } else {
          unsigned long long o = atomicAdd(P.over_n, 1ull);
          if (o < P.over_cap) {
            const uint64_t line = ((uint64_t)bin << P.lbits) | (lowpart >> P.tbits);
            P.over_lt[o] = line * P.J.D + (lowpart & ((1ull << P.tbits) - 1ull));
            P.over_rec[o] = rec;
          } else atomicAdd(&P.acc[2], 1ull);               // reported as MFN_ERR_CAPACITY at fetch
        }
      }
    }
    SW_PROF_MARK(6);
  }
  flush_rec();
  for (int o = 16; o > 0; o >>= 1) my_valid += __shfl_down_sync(0xffffffffu, my_valid, o);
  if ((tid & 31) == 0 && my_valid) atomicAdd(&P.acc[0], my_valid);
}

// ------------------------------------------------------------------------------------------
// pass 2
// ------------------------------------------------------------------------------------------

// Everything the home sector alone does not settle (spilled or absent k-mer, saturated value,
// v_read == 0, pair outside the shared-memory tier, k-mer that counts twice): complete lookup and
// the general accumulation, as in k_scan.
__device__ __forceinline__ void sweep_slow(const SweepParams &P, uint32_t *s_hist, uint32_t line, uint64_t tag,
                                           uint32_t rec_flag) {
  uint32_t vr, va;                                       // side tables are keyed by k-mer
  const uint64_t c = mfn_unmix((uint64_t)line * P.J.D + tag, P.J.seed, P.J.kmask, P.J.k);
  joint_full(P.J, c, line, tag, vr, va);
  // a restricted table stores no k-mer with both values 0: (0, 0) means "not held here", and the
  // per-database tables answer (a sequence other than the one the table was restricted to)
  if (P.restricted && (vr | va) == 0) { vr = mfn_value(P.R, c); va = mfn_value(P.A, c); }
  if (rec_flag & 0x80000000u) { vr += vr; va += va; }    // value(f) + value(r), f == r (Q2)
  const uint32_t rec = rec_flag & 0x7fffffffu;
  bool miss = (vr == 0);
  if (vr != 0 && vr < P.zlen) miss = (__ldg(&P.zmap[vr >> 6]) >> (vr & 63)) & 1;
  if (miss) atomicAdd(&P.acc[MFN_ACC_HEADER + MFN_DENSE_R * MFN_DENSE_A + 2 * rec + 1], 1ull);
  if (vr < SH_R && va < SH_A) atomicAdd(&s_hist[va * SH_R + vr], 1u);
  else if (vr < MFN_DENSE_R && va < MFN_DENSE_A) atomicAdd(&P.acc[MFN_ACC_HEADER + va * MFN_DENSE_R + vr], 1ull);
  else sparse_add(P.sparse, P.acc, vr, va, 1ull);
}

// a 32-byte sector of the table slice the grid is working on.  SW_TABLE_EVICT_LAST (experiment): ask the
// L2 to keep these lines over the entry stream (which is already loaded evict_first)
__device__ __forceinline__ void sweep_load_sector(const char *p, Bucket &b) {
#ifdef SW_TABLE_EVICT_LAST
  asm volatile("ld.global.nc.L1::no_allocate.L2::evict_last.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#else
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
#endif
               : "=r"(b.w[0]), "=r"(b.w[1]), "=r"(b.w[2]), "=r"(b.w[3]),
                 "=r"(b.w[4]), "=r"(b.w[5]), "=r"(b.w[6]), "=r"(b.w[7])
               : "l"(p));
}

// 32-byte home sector -> slot payload (both value fields), or ~0 when the tag is not in it.
// A stored slot is  tag << vb | values  with values != 0, so  slot - (tag << vb) - 1 < 2^vb - 1
// exactly for the matching slot; the minimum over the 8 slots finds it (VIADDMNMX / VIMNMX3).
__device__ __forceinline__ uint32_t sweep_match32(const Bucket &b, uint32_t nwant, uint32_t lim) {
  const uint32_t m0 = min(b.w[0] + nwant, b.w[1] + nwant), m1 = min(b.w[2] + nwant, b.w[3] + nwant);
  const uint32_t m2 = min(b.w[4] + nwant, b.w[5] + nwant), m3 = min(b.w[6] + nwant, b.w[7] + nwant);
  const uint32_t m = min(min(m0, m1), min(m2, m3));
  return m < lim ? m + 1u : 0xffffffffu;
}

template <bool W64>
__global__ void __launch_bounds__(SW_THREADS, 3) k_sweep_probe(const __grid_constant__ SweepParams P) {
  __shared__ uint32_t s_hist[SH_R * SH_A];
  __shared__ unsigned long long s_q[SW_THREADS / 32][SW_QCAP];    // per-warp queue of slow lookups
  __shared__ uint32_t s_qn[SW_THREADS / 32];
  const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (uint32_t i = tid; i < SH_R * SH_A; i += SW_THREADS) s_hist[i] = 0;
  if (tid < SW_THREADS / 32) s_qn[tid] = 0;
  __syncthreads();

  const uint64_t slices_per_bin = ceil_div_u64(P.cap, SW_SLICE);
  const uint64_t nitems = (uint64_t)P.nbins * slices_per_bin;
  const uint32_t tbits = P.tbits;
  const uint64_t tmask64 = (1ull << tbits) - 1ull;
  // entry fields: the low part is lowbits wide (32..34 narrow, up to 52 wide), the record above it
  const uint32_t lowbits = P.lowbits;
  const uint64_t lowmask = (1ull << lowbits) - 1ull;
  const uint32_t himask = (uint32_t)(lowmask >> 32);               // bits of the low part in the entry's high word
  const uint32_t recmask = (uint32_t)((1ull << (63 - lowbits)) - 1ull);
  auto recflag_of = [&](unsigned long long e) { return ((uint32_t)(e >> lowbits) & recmask) | ((uint32_t)(e >> 32) & 0x80000000u); };
  const uint32_t vb = P.J.rbits + P.J.abits, abits = P.J.abits, amax = P.J.amax;
  // pairs the fast path takes: rlo <= v_read < min(rmax, SH_R), v_asm < min(amax, SH_A); everything
  // else (v_read == 0 or zeroed by -prob, saturated fields, big values) goes to the queue
  const uint32_t rlo = P.zhi + 1u;
  const uint32_t rhi = min(P.J.rmax, SH_R), rspan = rhi > rlo ? rhi - rlo : 0u;
  const uint32_t alim = min(amax, SH_A);
  unsigned long long *const q = s_q[wid];
  uint32_t *const qn_p = &s_qn[wid];

  // one entry whose home sector is in registers
  auto settle = [&](unsigned long long e, const Bucket &b) {
    const uint32_t lw = (uint32_t)e, hw = (uint32_t)(e >> 32);   // bit 31 of hw: counts twice
    bool fast, settled;                           // settled: the home sector holds the k-mer
    uint32_t vr, va;
    if (!W64) {
      const uint32_t pay = sweep_match32(b, ~((lw << vb) & P.pr_wmask), P.pr_lim);
      vr = pay >> abits; va = pay & amax;
      settled = pay != 0xffffffffu;
      fast = settled && (vr - rlo) < rspan && va < alim && (int32_t)hw >= 0;
    } else {
      // 64-bit slots (4 per sector): same test on 64-bit words
      const uint64_t nwant = ~(((e & lowmask) & tmask64) << vb), lim64 = (1ull << vb) - 1ull;
      uint64_t pay = ~0ull;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const uint64_t d = (((uint64_t)b.w[2 * i + 1] << 32) | b.w[2 * i]) + nwant;
        if (d < lim64) pay = d + 1ull;
      }
      const uint64_t r64 = pay >> abits;
      vr = (uint32_t)r64; va = (uint32_t)pay & amax;
      settled = pay != ~0ull && r64 < P.J.rmax;
      fast = settled && r64 < rhi && vr >= rlo && va < alim && (int32_t)hw >= 0;
    }
    if (fast) atomicAdd(&s_hist[va * SH_R + vr], 1u);
    else if (settled && vr >= rlo && vr < min(P.J.rmax, MFN_DENSE_R) && va < min(amax, MFN_DENSE_A) && (int32_t)hw >= 0)
      atomicAdd(&P.acc[MFN_ACC_HEADER + va * MFN_DENSE_R + vr], 1ull);   // found, just outside the shared-memory tier: global tier, fire and forget
    else q[atomicAdd(qn_p, 1u)] = e;
  };

  // work items are handed out per WARP (no CTA barrier in the loop)
  while (true) {
    uint64_t item = 0;
    if (lane == 0) item = atomicAdd(P.counter, 1ull);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= nitems) break;
    const uint64_t bin = item / slices_per_bin, s = item % slices_per_bin;
    const uint64_t cnt = min(min(P.bin_count[bin * P.cstride], P.bin_valid[bin]), (unsigned long long)P.cap);
    if (s * SW_SLICE >= cnt) continue;
    const unsigned long long *src = P.entries + bin * P.cap + s * SW_SLICE;
    const uint32_t n = (uint32_t)min((uint64_t)SW_SLICE, cnt - s * SW_SLICE);
    const uint32_t line0 = (uint32_t)bin << P.lbits;
    const char *slice = (const char *)P.J.lines + (uint64_t)line0 * 128;

    // each lane: batches of 4 consecutive entries (32 B per lane, 1 KB per warp, coalesced);
    // full batches first, then one partial batch
    const uint32_t nfull = n >> 7;                                          // uniform
    auto load_entries = [&](uint32_t bt, unsigned long long (&e)[4]) {
      // streamed once: keep it from evicting the table slice
      asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.u64 {%0,%1,%2,%3}, [%4];"
                   : "=l"(e[0]), "=l"(e[1]), "=l"(e[2]), "=l"(e[3]) : "l"(src + bt * 128u + lane * 4u));
    };
    auto sector_of = [&](unsigned long long e) {
      if (W64) {
        const uint64_t lp = e & lowmask;
        return slice + (((lp >> tbits) << 2) | (lp & 3u)) * 32;
      }
      const uint32_t lw = (uint32_t)e;
      const uint32_t lib = __funnelshift_r(lw, (uint32_t)(e >> 32) & himask, tbits);   // line in bin (tbits < 32)
      return slice + (((uint64_t)lib << 2) | (lw & 3u)) * 32;
    };
    // the few lookups the home sector cannot settle are finished 32 at a time with every lane
    // busy, instead of dragging the whole warp through the slow path once per entry
    auto drain = [&](uint32_t thr) {
      __syncwarp();
      uint32_t qn = *(volatile uint32_t *)qn_p;
      while (qn >= thr) {
        const uint32_t take = min(qn, 32u);
        if (lane < take) {
          const unsigned long long x = q[qn - take + lane];
          const uint64_t lp = x & lowmask;
          bool done = false;
          if (!W64) {
            // Most of the queue are k-mers that spilled out of a full home sector into another sector
            // of the SAME line (2-5 % of the k-mers at 60-70 % load): the three other sectors (L2 hits,
            // the line has just been read) are searched here with the same tag test; only what they do
            // not settle either takes the complete lookup below.
            const uint32_t lw = (uint32_t)x, hw = (uint32_t)(x >> 32);
            const uint32_t lib = __funnelshift_r(lw, hw & himask, tbits);
            const char *linep = slice + (uint64_t)lib * 128;
            const uint32_t home = lw & 3u;
            Bucket b1, b2, b3;
            sweep_load_sector(linep + ((home + 1) & 3u) * 32, b1);
            sweep_load_sector(linep + ((home + 2) & 3u) * 32, b2);
            sweep_load_sector(linep + ((home + 3) & 3u) * 32, b3);
            const uint32_t nwant = ~((lw << vb) & P.pr_wmask);
            const uint32_t pay = min(sweep_match32(b1, nwant, P.pr_lim), min(sweep_match32(b2, nwant, P.pr_lim), sweep_match32(b3, nwant, P.pr_lim)));
            const uint32_t vr = pay >> abits, va = pay & amax;
            if (pay != 0xffffffffu && (vr - rlo) < rspan && va < alim && (int32_t)hw >= 0) { atomicAdd(&s_hist[va * SH_R + vr], 1u); done = true; }
          }
          if (!done) sweep_slow(P, s_hist, line0 + (uint32_t)(lp >> tbits), lp & tmask64, recflag_of(x));
        }
        qn -= take;
        __syncwarp();
        if (lane == 0) *qn_p = qn;
        __syncwarp();
      }
    };
    unsigned long long enext[4];
    if (nfull) load_entries(0, enext);
#pragma unroll 1
    for (uint32_t bt = 0; bt < nfull; bt++) {
      unsigned long long e[4];
#pragma unroll
      for (int u = 0; u < 4; u++) e[u] = enext[u];
      if (bt + 1 < nfull) load_entries(bt + 1, enext);         // in flight while this batch probes
      Bucket b[4];
#pragma unroll
      for (int u = 0; u < 4; u++) sweep_load_sector(sector_of(e[u]), b[u]);
#pragma unroll
      for (int u = 0; u < 4; u++) settle(e[u], b[u]);
      drain(32u);
    }
    // the tail: < 128 entries, one per lane per round
    for (uint32_t i = (nfull << 7) + lane; i < n; i += 32) {
      const unsigned long long e = src[i];
      Bucket b;
      sweep_load_sector(sector_of(e), b);
      settle(e, b);
    }
    drain(1u);                                                 // the queue is emptied before the bin changes
  }
  __syncthreads();
  for (uint32_t i = tid; i < SH_R * SH_A; i += SW_THREADS) {
    uint32_t c = s_hist[i];
    if (c) atomicAdd(&P.acc[MFN_ACC_HEADER + (i / SH_R) * MFN_DENSE_R + (i % SH_R)], (unsigned long long)c);
  }
}

// Windows that found no room in their bin (or whose chunk spanned too many records): probed
// from the side list, one complete lookup each.  Normally empty.
template <int UNUSED>
__global__ void __launch_bounds__(SW_THREADS) k_sweep_over(const __grid_constant__ SweepParams P) {
  __shared__ uint32_t s_hist[SH_R * SH_A];
  const uint32_t tid = threadIdx.x;
  for (uint32_t i = tid; i < SH_R * SH_A; i += SW_THREADS) s_hist[i] = 0;
  __syncthreads();
  const uint64_t n = min(*P.over_n, (unsigned long long)P.over_cap);
  for (uint64_t i = blockIdx.x * (uint64_t)SW_THREADS + tid; i < n; i += (uint64_t)gridDim.x * SW_THREADS) {
    uint64_t line, tag;
    joint_split(P.over_lt[i], P.J.fs_t, P.J.fs_d, P.J.fs_m, P.J.fs_s, line, tag);
    sweep_slow(P, s_hist, (uint32_t)line, tag, P.over_rec[i]);
  }
  __syncthreads();
  for (uint32_t i = tid; i < SH_R * SH_A; i += SW_THREADS) {
    uint32_t c = s_hist[i];
    if (c) atomicAdd(&P.acc[MFN_ACC_HEADER + (i / SH_R) * MFN_DENSE_R + (i % SH_R)], (unsigned long long)c);
  }
}

// One translation unit per probe entry (sweep_probe32.cu, sweep_probe64.cu, sweep_over.cu): ptxas
// 12.9 crashes when two entries that share an out-of-line device function are compiled in one module.
void mfn_sweep_launch_probe32(const SweepParams &P, int sm_count, cudaStream_t stream, int cap = 0);
void mfn_sweep_launch_probe64(const SweepParams &P, int sm_count, cudaStream_t stream);
void mfn_sweep_launch_over(const SweepParams &P, int grid, cudaStream_t stream);
