// -dump / path scoring through the range-partitioned scan (see sweep_kernels.cuh, sweep_scatter3.cuh).
//
// The direct scan (k_scan<DUMP>) pays one random 128-byte HBM line per window: 34 Gbases/s, 0.38 of the
// 72.25 B/base roofline.  -dump needs its answers in window order, which the two-pass scan does not give --
// but it can put them back:
//   pass 1  k_sweep_scatter3<.., DUMP>  as for -hist, and where[window] = index of the window's entry
//   pass 2  k_sweep_probe_dump          bins in order; every entry is OVERWRITTEN by its (v_read, v_asm) pair
//   pass 3  k_sweep_gather              windows in order: out[position] = entries[where[window]]
// Pass 3's 8-byte gathers are DRAM-friendly: the 8192 windows of a chunk left their entries in one run of
// ~128 bytes per bin, and the CTA that gathers the chunk reads all of them.  DRAM traffic per window:
// 8 + 4 written, 8 + table/windows read, 8 written, 4 + 8 read, 9 written -- about 55 bytes instead of 128+.
// (own translation unit: the probe kernel calls the out-of-line complete lookup; ptxas 12.9 crashes when two
// entries of a module share it)
#include "sweep_scatter3.cuh"

// (v_read, v_asm) of one entry whose home sector does not settle it: the other sectors of the line, then the
// complete lookup; a table restricted to other sequences falls back to the per-database tables
static __device__ __forceinline__ void dump_slow(const SweepParams &P, uint32_t line, uint64_t tag, uint32_t &vr, uint32_t &va) {
  const uint64_t c = mfn_unmix((uint64_t)line * P.J.D + tag, P.J.seed, P.J.kmask, P.J.k);
  joint_full(P.J, c, line, tag, vr, va);
  if (P.restricted && (vr | va) == 0) { vr = mfn_value(P.R, c); va = mfn_value(P.A, c); }
}

__global__ void __launch_bounds__(SW_THREADS, 3) k_sweep_probe_dump(const __grid_constant__ SweepParams P) {
  // per-warp queue of the entries the home sector does not settle (index in the work item): finished 32 at a
  // time with every lane busy, as in k_sweep_probe -- inline, every batch of 128 entries would wait for the
  // two or three dependent loads of its 2-4 % spilled k-mers
  __shared__ uint16_t s_q[SW_THREADS / 32][SW_QCAP];
  __shared__ uint32_t s_qn[SW_THREADS / 32];
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) s_qn[wid] = 0;
  __syncwarp();
  uint16_t *const q = s_q[wid];
  uint32_t *const qn_p = &s_qn[wid];
  const uint64_t slices_per_bin = ceil_div_u64(P.cap, SW_SLICE);
  const uint64_t nitems = (uint64_t)P.nbins * slices_per_bin;
  const uint32_t tbits = P.tbits;
  const uint64_t tmask64 = (1ull << tbits) - 1ull;
  const uint64_t lowmask = (1ull << P.lowbits) - 1ull;
  const uint32_t himask = (uint32_t)(lowmask >> 32);
  const uint32_t vb = P.J.rbits + P.J.abits, abits = P.J.abits, amax = P.J.amax, rmax = P.J.rmax;

  while (true) {
    uint64_t item = 0;
    if (lane == 0) item = atomicAdd(P.counter, 1ull);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= nitems) break;
    const uint64_t bin = item / slices_per_bin, s = item % slices_per_bin;
    const uint64_t cnt = min(P.bin_count[bin * P.cstride], (unsigned long long)P.cap);
    if (s * SW_SLICE >= cnt) continue;
    unsigned long long *src = P.entries + bin * P.cap + s * SW_SLICE;
    const uint32_t n = (uint32_t)min((uint64_t)SW_SLICE, cnt - s * SW_SLICE);
    const uint32_t line0 = (uint32_t)bin << P.lbits;
    const char *slice = (const char *)P.J.lines + (uint64_t)line0 * 128;

    auto line_in_bin = [&](unsigned long long e) { return __funnelshift_r((uint32_t)e, (uint32_t)(e >> 32) & himask, tbits); };
    auto sector_of = [&](unsigned long long e) { return slice + (((uint64_t)line_in_bin(e) << 2) | ((uint32_t)e & 3u)) * 32; };
    // entry i of the item + its home sector: the pair (v_asm << 32 | v_read) is stored over the entry, or the
    // entry is queued
    // -> true and the pair, or false (the entry has to take the queue)
    auto lookup = [&](unsigned long long e, const Bucket &b, unsigned long long &pair) -> bool {
      const uint32_t lw = (uint32_t)e;
      const uint32_t pay = sweep_match32(b, ~((lw << vb) & P.pr_wmask), P.pr_lim);
      const uint32_t vr = pay >> abits, va = pay & amax;
      pair = ((unsigned long long)va << 32) | vr;
      return pay != 0xffffffffu && vr < rmax && va < amax;
    };
    auto settle = [&](uint32_t i, unsigned long long e, const Bucket &b) {
      unsigned long long pair;
      if (lookup(e, b, pair)) src[i] = pair;
      else q[atomicAdd(qn_p, 1u)] = (uint16_t)i;
    };
    auto drain = [&](uint32_t thr) {
      __syncwarp();
      uint32_t qn = *(volatile uint32_t *)qn_p;
      while (qn >= thr) {
        const uint32_t take = min(qn, 32u);
        if (lane < take) {
          const uint32_t i = q[qn - take + lane];
          const unsigned long long e = src[i];
          const uint64_t lp = e & lowmask;
          uint32_t vr, va;
          dump_slow(P, line0 + (uint32_t)(lp >> tbits), lp & tmask64, vr, va);
          src[i] = ((unsigned long long)va << 32) | vr;
        }
        qn -= take;
        __syncwarp();
        if (lane == 0) *qn_p = qn;
        __syncwarp();
      }
    };
    const uint32_t nfull = n >> 7;
#pragma unroll 1
    for (uint32_t bt = 0; bt < nfull; bt++) {
      unsigned long long e[4];
      const uint32_t i0 = bt * 128u + lane * 4u;
      asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(e[0]), "=l"(e[1]), "=l"(e[2]), "=l"(e[3]) : "l"(src + i0));
      Bucket b[4];
#pragma unroll
      for (int u = 0; u < 4; u++) sweep_load_sector(sector_of(e[u]), b[u]);
      // the usual lane has all four settled: one 32-byte store (an L2 request per entry costs as much as the probe itself)
      unsigned long long r[4];
      bool okk[4];
#pragma unroll
      for (int u = 0; u < 4; u++) okk[u] = lookup(e[u], b[u], r[u]);
      if (okk[0] && okk[1] && okk[2] && okk[3]) {
        asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" :: "l"(src + i0), "l"(r[0]), "l"(r[1]), "l"(r[2]), "l"(r[3]) : "memory");
      } else {
#pragma unroll
        for (int u = 0; u < 4; u++) {
          if (okk[u]) src[i0 + u] = r[u];
          else q[atomicAdd(qn_p, 1u)] = (uint16_t)(i0 + u);
        }
      }
      drain(32u);
    }
    for (uint32_t i = (nfull << 7) + lane; i < n; i += 32) {
      const unsigned long long e = src[i];
      Bucket b;
      sweep_load_sector(sector_of(e), b);
      settle(i, e, b);
    }
    drain(1u);                                                 // the queue is emptied before the item changes
  }
}

// pass 3: windows in order, one CTA per chunk of the scatter pass.  The chunk's pairs sit in one run of ~16 entries
// per bin; the CTA copies whole runs into shared memory (a warp per run: consecutive lanes on consecutive entries,
// i.e. DRAM reads of whole sectors and lines -- a first version gathered 8 bytes per window straight from global
// memory and ran at the random-sector rate of the DRAM, 10.3 ms per Gbp), then every window picks its pair by
// (bin, rank).  Positions are reported in the caller's coordinates (one separator slot per earlier record, as
// k_scan<DUMP> does).
template <int TILES>
__global__ void __launch_bounds__(TILES * 128) k_sweep_gather(const uint32_t *__restrict__ where, const uint2 *__restrict__ runs,
                                                              const unsigned long long *__restrict__ entries, uint32_t nbins_pad,
                                                              const unsigned long long *__restrict__ rec_start, uint32_t nseq,
                                                              uint64_t g_begin, uint64_t g_end, uint32_t *out_vr, uint32_t *out_va, uint8_t *out_valid) {
  constexpr uint32_t THREADS = TILES * 128, CHUNK = TILES * MFN_TILE, NW = CHUNK / THREADS;
  extern __shared__ __align__(16) unsigned char gsm[];
  unsigned long long *s_ent = (unsigned long long *)gsm;                // [CHUNK] the chunk's pairs, bin by bin
  uint32_t *s_off = (uint32_t *)(s_ent + CHUNK);                        // [nbins_pad] where each bin's run starts in s_ent
  uint32_t *s_start = s_off + nbins_pad;                                // [nbins_pad] ... and in P.entries
  __shared__ uint32_t s_warp[THREADS / 32];
  const uint32_t tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint64_t chunk = blockIdx.x;
  const uint2 *const run = runs + chunk * nbins_pad;
  const uint64_t t0 = chunk * CHUNK;                                    // first window of the chunk, relative to g_begin
  const uint64_t gc = g_begin + t0;
  // this thread's windows (tid, tid + THREADS, ...): their (bin, rank) words are on their way while the runs are copied
  uint32_t wv[NW];
#pragma unroll
  for (uint32_t u = 0; u < NW; u++) wv[u] = gc + tid + THREADS * u < g_end ? __ldg(where + t0 + tid + THREADS * u) : 0xfffffffeu;
  // exclusive scan of the run lengths (nbins_pad is a multiple of THREADS: 1 or 2 bins per thread)
  const uint32_t bpt = nbins_pad / THREADS, b0 = tid * bpt;
  const uint2 r0 = run[b0], r1 = bpt > 1 ? run[b0 + 1] : make_uint2(0u, 0u);
  const uint32_t c0 = r0.y, c1 = r1.y;
  s_start[b0] = r0.x;
  if (bpt > 1) s_start[b0 + 1] = r1.x;
  {
    const uint32_t sum = c0 + c1;
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += v; }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    uint32_t wbase = 0;
    for (uint32_t x = 0; x < wid; x++) wbase += s_warp[x];
    const uint32_t at = wbase + incl - sum;
    s_off[b0] = at;
    if (bpt > 1) s_off[b0 + 1] = at + c0;
  }
  __syncthreads();
  // the chunk's entries, bin by bin: thread j copies sorted positions j, j + THREADS, ... (consecutive lanes on
  // consecutive entries of a run; all of a thread's loads in flight together -- a warp per run with its
  // data-dependent inner loop waited for one DRAM access per run, 32 in a row: 37 us per CTA)
  {
    const uint32_t total = s_off[nbins_pad - 1] + run[nbins_pad - 1].y;
    unsigned long long v[NW];
    uint32_t nbits = 0;
    while ((1u << nbits) < nbins_pad) nbits++;
#pragma unroll
    for (uint32_t u = 0; u < NW; u++) {
      const uint32_t p = tid + THREADS * u;
      v[u] = 0;
      if (p < total) {
        uint32_t b = 0;                                                   // largest b with s_off[b] <= p
        for (uint32_t step = 1u << (nbits - 1); step; step >>= 1) if (b + step < nbins_pad && s_off[b + step] <= p) b += step;
        v[u] = __ldg(entries + s_start[b] + (p - s_off[b]));
      }
    }
#pragma unroll
    for (uint32_t u = 0; u < NW; u++) { const uint32_t p = tid + THREADS * u; if (p < total) s_ent[p] = v[u]; }
  }
  __syncthreads();
  uint32_t rec;
  {
    uint32_t a = 0, b = nseq;
    while (a < b) { const uint32_t m = (a + b + 1) >> 1; if (rec_start[m] <= gc) a = m; else b = m - 1; }
    rec = a;
  }
  uint64_t next_start = rec < nseq ? rec_start[rec + 1] : ~0ull;
#pragma unroll
  for (uint32_t u = 0; u < NW; u++) {
    const uint32_t w = wv[u];
    if (w == 0xfffffffeu) break;                                        // past the end of the range
    const uint64_t g = gc + tid + THREADS * u;
    while (g >= next_start) { rec++; next_start = rec < nseq ? rec_start[rec + 1] : ~0ull; }
    if (rec >= nseq) break;
    if (g + 1 == next_start) continue;                                  // the separator slot: no position of the caller's
    const bool ok = w != 0xffffffffu;
    const unsigned long long r = ok ? s_ent[s_off[w >> 13] + (w & 0x1fffu)] : 0ull;
    const uint64_t sidx = g - rec;                                      // every position of the range is written: no memset needed
    out_vr[sidx] = (uint32_t)r; out_va[sidx] = (uint32_t)(r >> 32); out_valid[sidx] = ok ? 1 : 0;
  }
}

// Host side.  *done_tiles = how many leading tiles of the sequence set got their outputs from the
// range-partitioned path (0: it does not apply); the caller runs the direct scan over the rest.
int mfn_sweep_dump(mfn_ctx *ctx, const mfn_table *R, const mfn_table *A, const mfn_seq *seq, uint32_t *d_vr, uint32_t *d_va,
                   uint8_t *d_valid, uint64_t *done_tiles) {
  *done_tiles = 0;
  const char *env = getenv("MFN_DUMP_SWEEP");
  if (env && atoi(env) == 0) return MFN_OK;
  const bool force = env && atoi(env) == 1;
  if (!force && seq->ntiles * MFN_TILE < (1ull << 24)) return MFN_OK;
  if (!mfn_sweep_dump_applicable(R, A, seq->ntiles, force)) return MFN_OK;
  // pieces of at most ~1.5e9 windows: entry indices stay below 2^32
  const uint64_t piece_tiles = 720000;
  for (uint64_t tb = 0; tb < seq->ntiles; tb += piece_tiles) {
    const uint64_t te = std::min(seq->ntiles, tb + piece_tiles);
    int ok = 0;
    MFN_TRY(mfn_sweep_dump_piece(ctx, R, A, seq, tb, te, d_vr, d_va, d_valid, &ok));
    if (!ok) break;                                          // a bin overflowed: the direct scan takes it from here
    *done_tiles = te;
  }
  return MFN_OK;
}

void mfn_sweep_dump_launch_probe(const SweepParams &P, int sm_count, cudaStream_t stream) {
  k_sweep_probe_dump<<<sm_count * 3, SW_THREADS, 0, stream>>>(P);
}

void mfn_sweep_dump_launch_gather(const SweepParams &P, int tiles, uint64_t nchunks, uint64_t g_begin, uint64_t g_end,
                                  uint32_t *out_vr, uint32_t *out_va, uint8_t *out_valid, cudaStream_t stream) {
  const size_t smem = (size_t)tiles * MFN_TILE * 8 + (size_t)P.nbins_pad * 8;
  if (tiles == 2) {
    cudaFuncSetAttribute(k_sweep_gather<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_sweep_gather<2><<<(unsigned)nchunks, 256, smem, stream>>>(P.where, P.runs, P.entries, P.nbins_pad, P.rec_start, P.nseq, g_begin, g_end, out_vr, out_va, out_valid);
  } else {
    cudaFuncSetAttribute(k_sweep_gather<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_sweep_gather<4><<<(unsigned)nchunks, 512, smem, stream>>>(P.where, P.runs, P.entries, P.nbins_pad, P.rec_start, P.nseq, g_begin, g_end, out_vr, out_va, out_valid);
  }
}
