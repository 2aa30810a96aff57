// Range-partitioned -hist scan: host side and the two main passes (see sweep_kernels.cuh).
#include "sweep_scatter3.cuh"

// ------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------
static uint32_t bits_of(uint64_t x) { uint32_t n = 0; while (x) { n++; x >>= 1; } return n; }

// The joint table a sweep of (R, A) runs against: the one restricted to the scanned sequences when
// there is one (mfn_tables_restrict), else the full fused table.
static const mfn_joint *sweep_table(const mfn_table *R, const mfn_table *A) {
  const char *env = getenv("MFN_NO_RESTRICTED");
  if (R->hot && R->hot->asm_ == A && !(env && atoi(env))) return R->hot;
  if (R->joint && R->joint->asm_ == A) return R->joint;
  return nullptr;
}

// Bins of 2^lbits lines: a slice of the table of at most MFN_SWEEP_SLICE_MB (default 32 MB: two of them in
// flight fit the L2 next to the entry stream), at most SW_MAX_BINS bins.  The entry's low part holds
// (line in bin) << tbits | tag: up to 34 bits in a narrow entry, else up to 52 bits for tables with 64-bit
// slots (wide entries, fewer bits left for the record).  Measured at the 3 Gbp benchmark (restricted table of
// 2^27 lines; profiles/r02_call12_summary.txt): 512 bins of 32 MB = scatter 12.8 ms (8192-window chunks) +
// probe 12.3 ms; 256 bins of 64 MB = scatter 10.65 ms (4096-window chunks, 4 CTAs per SM) + probe 14.5 ms
// (two slices in flight no longer fit the L2: the table is read 2.3 x).  Same sum, but the host-buffer path
// exposes the last probe pass, so the shorter probe wins end to end (39.2 against 42.3 ms).
static bool sweep_geometry(const JointDev &J, uint32_t &lbits, uint32_t &tbits, uint32_t &nbins, uint32_t &lowbits) {
  tbits = std::max(1u, bits_of(J.D - 1));
  if (J.nlines >= (1ull << 32)) return false;
  const char *env = getenv("MFN_SWEEP_SLICE_MB");
  uint32_t slice_mb = env ? (uint32_t)atoi(env) : 32u;
  if (slice_mb < 1 || slice_mb > 256) slice_mb = 32;
  const uint32_t lb_want = bits_of((uint64_t)slice_mb * (1u << 20) / 128) - 1;     // 32 MB -> 18
  for (int wide = 0; wide < 2; wide++) {
    const uint32_t room = wide ? 52u : (uint32_t)SW_MAX_LOWBITS_NARROW;
    if (wide && !J.w64) break;
    if (tbits + 1 > room) continue;
    for (uint32_t lb = std::min(lb_want, room - tbits); lb <= room - tbits && lb <= 24; lb++) {
      lbits = lb;
      nbins = (uint32_t)ceil_div_u64(J.nlines, 1ull << lbits);
      if (nbins <= SW_MAX_BINS) { lowbits = wide ? lbits + tbits : std::max(32u, lbits + tbits); return true; }
    }
  }
  return false;
}

// Can (and should) this scan use the sweep?
bool mfn_sweep_applicable(const mfn_table *R, const mfn_table *A, uint64_t ntiles, uint32_t nseq) {
  if (R->streamed) {
    // no per-database table to fall back on: the range-partitioned scan of the restricted table or nothing
    const mfn_joint *T = R->hot && R->hot->asm_ == A ? R->hot : nullptr;
    uint32_t lbits, tbits, nbins, lowbits;
    return T && sweep_geometry(T->dev, lbits, tbits, nbins, lowbits) && (uint64_t)nseq + 1 < (1ull << (63 - lowbits));
  }
  const int mode = R->ctx ? R->ctx->sweep_mode : -1;
  if (mode == 0) return false;
  const char *env = getenv("MFN_SWEEP");
  if (env && atoi(env) == 0) return false;
  const mfn_joint *T = sweep_table(R, A);
  if (!T) return false;
  uint32_t lbits, tbits, nbins, lowbits;
  if (!sweep_geometry(T->dev, lbits, tbits, nbins, lowbits)) return false;
  if ((uint64_t)nseq + 1 >= (1ull << (63 - lowbits))) return false;       // record ids must fit the entry
  if (mode == 1 || (env && atoi(env) == 1)) return true;
  // wide entries have no room for a record delta: windows of a chunk that are not in the chunk's
  // first record take the side list, which is only acceptable when records are long
  if (lowbits > SW_MAX_LOWBITS_NARROW && ntiles * MFN_TILE < (uint64_t)nseq * 16 * SW_CHUNK) return false;
  // worth it when the windows revisit the table lines: >= 1 window per 4 lines, and a big batch
  const uint64_t windows = ntiles * MFN_TILE;
  return windows >= (1ull << 24) && windows * 4 >= T->dev.nlines;
}

// A sweep over a sequence set: begin (workspace sized for `windows`), scatter any number of tile
// ranges as their bases become resident, finish (probe every bin once).
struct mfn_sweep {
  SweepParams P{};
  unsigned long long *ctl = nullptr, *counters = nullptr;
  uint32_t n_scatter = 0;
};

int mfn_sweep_begin(mfn_ctx *ctx, const mfn_table *R, const mfn_table *A, const mfn_seq *seq, uint64_t windows, mfn_accum *acc,
                    const unsigned long long *zmap, uint32_t zlen, uint32_t zhi, mfn_sweep **out) {

  const mfn_joint *T = sweep_table(R, A);
  if (!T) { mfn_set_error("sweep: no fused table"); return MFN_ERR_INVALID; }
  const JointDev &J = T->dev;
  mfn_sweep *w = new mfn_sweep();
  SweepParams &P = w->P;
  P.J = J;
  P.restricted = T->restricted ? 1u : 0u;
  if (T->restricted) { P.R = T->reads->dev; P.A = T->asm_->dev; }
  P.seq2 = seq->d_seq2; P.ok = seq->d_ok; P.rec_start = seq->d_rec_start; P.nseq = seq->nseq;
  {
    const uint64_t gpad = (seq->ntiles + 1) * MFN_TILE + 256;            // as allocated by mfn_seq_alloc
    P.seq_words = gpad / 32 + 8; P.ok_words = gpad / 64 + 8;
  }
  if (!sweep_geometry(J, P.lbits, P.tbits, P.nbins, P.lowbits)) { delete w; mfn_set_error("sweep: table geometry not supported"); return MFN_ERR_UNSUPPORTED; }
  if ((uint64_t)seq->nseq + 1 >= (1ull << (63 - P.lowbits))) { delete w; mfn_set_error("sweep: %u records do not fit the entry", seq->nseq); return MFN_ERR_UNSUPPORTED; }
  P.nbins_pad = (P.nbins + SC_THREADS - 1) / SC_THREADS * SC_THREADS;
  P.cap = ((uint64_t)((double)windows / P.nbins * 1.15) + 2048 + 31) & ~31ull;
  P.over_cap = std::max<uint64_t>(1 << 20, windows / 16);
  P.acc = acc->d_buf; P.sparse = acc->sparse; P.zmap = zmap; P.zlen = zlen; P.zhi = zhi;
  {
    const uint32_t vb = (J.rbits + J.abits) & 31u;
    P.pr_lim = (1u << vb) - 1u;
    P.pr_wmask = ((1u << P.tbits) - 1u) << vb;
  }
  // workspace from the stream-ordered pool: reused across steps without cudaMalloc cost
  const size_t ebytes = (size_t)P.nbins * P.cap * 8;
  {
    const char *es = getenv("MFN_BIN_STRIDE");
    P.cstride = es ? (uint32_t)atoi(es) : 1u;
    if (P.cstride < 1 || P.cstride > 64) P.cstride = 1;
  }
  const size_t cbins = (size_t)P.nbins_pad * P.cstride;               // counters, then bin_valid
  const size_t cwords = cbins + (size_t)P.nbins_pad + 4 + 4096 + 8;      // bins, side-list count, pass counters (one per scatter pass), phase cycles
  cudaError_t e = cudaMallocAsync(&P.entries, ebytes, ctx->stream);
  if (e == cudaSuccess) e = cudaMallocAsync(&w->ctl, cwords * 8, ctx->stream);
  if (e == cudaSuccess) e = cudaMallocAsync(&P.over_lt, P.over_cap * 8, ctx->stream);
  if (e == cudaSuccess) e = cudaMallocAsync(&P.over_rec, P.over_cap * 4, ctx->stream);
  if (e != cudaSuccess) { mfn_set_error("sweep workspace (%zu bytes): %s", ebytes, cudaGetErrorString(e)); delete w; return MFN_ERR_NOMEM; }
  P.bin_count = w->ctl; P.bin_valid = w->ctl + cbins; P.over_n = w->ctl + cbins + P.nbins_pad;
  w->counters = w->ctl + cbins + P.nbins_pad + 1;
  P.prof = w->ctl + cbins + P.nbins_pad + 4 + 4096;
  MFN_CUDA(cudaMemsetAsync(w->ctl, 0, cwords * 8, ctx->stream));
  MFN_CUDA(cudaMemsetAsync(P.bin_valid, 0xff, (size_t)P.nbins_pad * 8, ctx->stream));
  *out = w;
  return MFN_OK;
}

int mfn_sweep_scatter(mfn_ctx *ctx, mfn_sweep *w, uint64_t tile_begin, uint64_t tile_end) {
  if (tile_begin >= tile_end) return MFN_OK;
  if (w->n_scatter >= 4000) { mfn_set_error("sweep: too many scatter passes"); return MFN_ERR_INVALID; }
  SweepParams P = w->P;
  P.tile_begin = tile_begin; P.tile_end = tile_end;
  const size_t sc_smem = (size_t)SW_CHUNK * 8 + (size_t)P.nbins_pad * 12;
  P.counter = w->counters + 2 + w->n_scatter++;
  const uint64_t nchunks = ceil_div_u64(tile_end - tile_begin, SW_CHUNK_TILES);
  // the mer size is a compile-time constant for the common k (shifts and masks fold away)
  auto launch = [&](auto kernel) -> int {
    int per_sm = 0;
    MFN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc_smem));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, SC_THREADS, sc_smem);
    const int grid = (int)std::min<uint64_t>((uint64_t)ctx->sm_count * std::max(per_sm, 1), nchunks);
    kernel<<<grid, SC_THREADS, sc_smem, ctx->stream>>>(P);
    return MFN_OK;
  };
  // bit-field geometry: power-of-two divisor, 34-bit low part, bin = the hash bits above it
  const bool geo1 = P.J.fs_d == 1 && P.lowbits == 34 && P.lbits + P.tbits == 34 && P.tbits == P.J.fs_t &&
                    2 * P.J.k >= 34 && 2 * P.J.k - 34 <= SC_BIN_BITS && P.nbins == (1u << (2 * P.J.k - 34));
  // scatter3 (sweep_scatter3.cuh): k = 21, narrow entries, few bins, entry indices below 2^32
  const char *env3 = getenv("MFN_SCATTER");
  const bool s3 = !(env3 && atoi(env3) == 2) && P.J.k == 21 && P.lowbits <= SW_MAX_LOWBITS_NARROW && P.nbins <= S3_MAX_BINS &&
                  (uint64_t)P.nbins * P.cap + 8192 < (1ull << 32);
  if (s3) {
    const char *envt = getenv("MFN_S3_TILES");
    const int tiles = envt && (atoi(envt) == 2 || atoi(envt) == 4) ? atoi(envt) : (P.nbins <= 256 ? 2 : 4);
    const uint32_t threads3 = (uint32_t)tiles * 128;
    P.nbins_pad = (P.nbins + threads3 - 1) / threads3 * threads3;
    const size_t smem3 = s3_smem_bytes(P.nbins_pad, tiles);
    const uint64_t nchunks3 = ceil_div_u64(tile_end - tile_begin, (uint64_t)tiles);
    auto launch3 = [&](auto kernel) -> int {
      int per_sm = 0;
      MFN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, (int)threads3, smem3);
      const int grid = (int)std::min<uint64_t>((uint64_t)ctx->sm_count * std::max(per_sm, 1), nchunks3);
      kernel<<<grid, threads3, smem3, ctx->stream>>>(P);
      return MFN_OK;
    };
    // the same with a 33-bit low part (twice the bins of half the size)
    const bool geo2 = P.J.fs_d == 1 && P.lowbits == 33 && P.lbits + P.tbits == 33 && P.tbits == P.J.fs_t &&
                      2 * P.J.k - 33 <= SC_BIN_BITS && P.nbins == (1u << (2 * P.J.k - 33));
    const int geo = geo1 ? 1 : geo2 ? 2 : 0;
    if (tiles == 2) {
      if (geo == 1) MFN_TRY(launch3(k_sweep_scatter3<21, 1, 2>)); else if (geo == 2) MFN_TRY(launch3(k_sweep_scatter3<21, 2, 2>)); else MFN_TRY(launch3(k_sweep_scatter3<21, 0, 2>));
    } else {
      if (geo == 1) MFN_TRY(launch3(k_sweep_scatter3<21, 1, 4>)); else if (geo == 2) MFN_TRY(launch3(k_sweep_scatter3<21, 2, 4>)); else MFN_TRY(launch3(k_sweep_scatter3<21, 0, 4>));
    }
  } else if (P.lowbits > SW_MAX_LOWBITS_NARROW) {
    switch (P.J.k) {
      case 31: MFN_TRY(launch(k_sweep_scatter<31, true, 0>)); break;
      default: MFN_TRY(launch(k_sweep_scatter<0, true, 0>)); break;
    }
  } else {
    switch (P.J.k) {
      case 21: if (geo1) MFN_TRY(launch(k_sweep_scatter<21, false, 1>)); else MFN_TRY(launch(k_sweep_scatter<21, false, 0>)); break;
      case 31: MFN_TRY(launch(k_sweep_scatter<31, false, 0>)); break;
      default: MFN_TRY(launch(k_sweep_scatter<0, false, 0>)); break;
    }
  }
  ctx->launches++;
  MFN_CUDA(cudaGetLastError());
  return MFN_OK;
}

// side_stream: the probe pass (and the release of the workspace) goes to the ctx's second compute stream, ordered
// after everything queued on ctx->stream so far, and ctx->stream does NOT wait for it: the scatter passes of the
// next sub-sweep run next to it (mfn_hist_host; the caller joins the streams at the end with mfn_sweep_join)
int mfn_sweep_finish(mfn_ctx *ctx, mfn_sweep *w, bool side_stream) {
  SweepParams P = w->P;
  cudaStream_t st = ctx->stream;
  if (side_stream) {
    st = ctx->probe_stream;
    cudaEventRecord(ctx->ev_probe[0], ctx->stream);
    cudaStreamWaitEvent(st, ctx->ev_probe[0], 0);
  }
  P.counter = w->counters;
  if (P.J.w64) mfn_sweep_launch_probe64(P, ctx->sm_count, st);
  else mfn_sweep_launch_probe32(P, ctx->sm_count, st, side_stream ? ctx->side_probe_ctas : 0);
  P.counter = w->counters + 1;
  mfn_sweep_launch_over(P, ctx->sm_count, st);
  ctx->launches += 2;
  cudaError_t e = cudaGetLastError();
  if (side_stream) {
    cudaFreeAsync(P.entries, st);
    cudaFreeAsync(w->ctl, st);
    cudaFreeAsync(P.over_lt, st);
    cudaFreeAsync(P.over_rec, st);
    delete w;
    if (e != cudaSuccess) { mfn_set_error("sweep: %s", cudaGetErrorString(e)); return MFN_ERR_CUDA; }
    return MFN_OK;
  }
#ifdef SW_PROFILE
  {
    unsigned long long h[8];
    cudaMemcpyAsync(h, P.prof, sizeof h, cudaMemcpyDeviceToHost, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    unsigned long long t = 0;
    for (int i = 0; i < 7; i++) t += h[i];
    fprintf(stderr, "[scatter phases, %% of CTA cycles] zero+claim %.1f  hash %.1f  barrier %.1f  scan %.1f  sort %.1f  barrier %.1f  copy-out %.1f  (bins %u, lowbits %u)\n",
            100.0 * h[0] / t, 100.0 * h[1] / t, 100.0 * h[2] / t, 100.0 * h[3] / t, 100.0 * h[4] / t, 100.0 * h[5] / t, 100.0 * h[6] / t, P.nbins, P.lowbits);
  }
#endif
  // a side list that overflowed is reported through the accumulator header (MFN_ERR_CAPACITY at fetch)
  cudaFreeAsync(P.entries, ctx->stream);
  cudaFreeAsync(w->ctl, ctx->stream);
  cudaFreeAsync(P.over_lt, ctx->stream);
  cudaFreeAsync(P.over_rec, ctx->stream);
  delete w;
  if (e != cudaSuccess) { mfn_set_error("sweep: %s", cudaGetErrorString(e)); return MFN_ERR_CUDA; }
  return MFN_OK;
}

int mfn_sweep_hist(mfn_ctx *ctx, const mfn_table *R, const mfn_table *A, const mfn_seq *seq, const mfn_kmodel *km,
                   uint64_t tile_begin, uint64_t tile_end, mfn_accum *acc, const unsigned long long *zmap, uint32_t zlen,
                   uint32_t zhi) {
  (void)km;
  if (tile_begin >= tile_end) return MFN_OK;
  mfn_sweep *w = nullptr;
  MFN_TRY(mfn_sweep_begin(ctx, R, A, seq, (tile_end - tile_begin) * MFN_TILE, acc, zmap, zlen, zhi, &w));
  if (ctx->time_passes) mfn_ctx_mark(ctx);
  int st = mfn_sweep_scatter(ctx, w, tile_begin, tile_end);
  if (ctx->time_passes) mfn_ctx_mark(ctx);
  int st2 = mfn_sweep_finish(ctx, w, false);
  if (ctx->time_passes) mfn_ctx_mark(ctx);
  return st != MFN_OK ? st : st2;
}

// ------------------------------------------------------------------------------------------
// -dump / path scoring through the range-partitioned scan (sweep_dump.cu)
// ------------------------------------------------------------------------------------------
void mfn_sweep_dump_launch_probe(const SweepParams &P, int sm_count, cudaStream_t stream);
void mfn_sweep_dump_launch_gather(const SweepParams &P, int tiles, uint64_t nchunks, uint64_t g_begin, uint64_t g_end,
                                  uint32_t *out_vr, uint32_t *out_va, uint8_t *out_valid, cudaStream_t stream);

// k = 21 (scatter3), canonical 32-bit-slot table, narrow entries, at most 512 bins
bool mfn_sweep_dump_applicable(const mfn_table *R, const mfn_table *A, uint64_t ntiles, bool force) {
  if (R->streamed) return false;
  const mfn_joint *T = sweep_table(R, A);
  if (!T || T->dev.k != 21 || T->dev.w64) return false;
  uint32_t lbits, tbits, nbins, lowbits;
  if (!sweep_geometry(T->dev, lbits, tbits, nbins, lowbits)) return false;
  if (lowbits > SW_MAX_LOWBITS_NARROW || nbins > S3_MAX_BINS) return false;
  const uint64_t windows = ntiles * MFN_TILE;
  return force || windows * 4 >= T->dev.nlines;               // worth it when the windows revisit the table lines
}

int mfn_sweep_dump_piece(mfn_ctx *ctx, const mfn_table *R, const mfn_table *A, const mfn_seq *seq, uint64_t tile_begin, uint64_t tile_end,
                         uint32_t *d_vr, uint32_t *d_va, uint8_t *d_valid, int *ok) {
  *ok = 0;
  const mfn_joint *T = sweep_table(R, A);
  const JointDev &J = T->dev;
  SweepParams P{};
  P.J = J;
  P.restricted = T->restricted ? 1u : 0u;
  if (T->restricted) { P.R = T->reads->dev; P.A = T->asm_->dev; }
  P.seq2 = seq->d_seq2; P.ok = seq->d_ok; P.rec_start = seq->d_rec_start; P.nseq = seq->nseq;
  {
    const uint64_t gpad = (seq->ntiles + 1) * MFN_TILE + 256;
    P.seq_words = gpad / 32 + 8; P.ok_words = gpad / 64 + 8;
  }
  if (!sweep_geometry(J, P.lbits, P.tbits, P.nbins, P.lowbits)) return MFN_OK;
  const uint64_t windows = (tile_end - tile_begin) * MFN_TILE;
  P.cap = ((uint64_t)((double)windows / P.nbins * 1.15) + 2048 + 31) & ~31ull;
  if ((uint64_t)P.nbins * P.cap + 8192 >= (1ull << 32)) return MFN_OK;
  P.cstride = 1;
  P.tile_begin = tile_begin; P.tile_end = tile_end;
  {
    const uint32_t vb = (J.rbits + J.abits) & 31u;
    P.pr_lim = (1u << vb) - 1u;
    P.pr_wmask = ((1u << P.tbits) - 1u) << vb;
  }
  const int tiles = P.nbins <= 256 ? 2 : 4;
  const uint32_t threads3 = (uint32_t)tiles * 128;
  P.nbins_pad = (P.nbins + threads3 - 1) / threads3 * threads3;
  const uint64_t nchunks = ceil_div_u64(tile_end - tile_begin, (uint64_t)tiles);
  const size_t nwhere = (size_t)nchunks * tiles * MFN_TILE;
  unsigned long long *ctl = nullptr;
  const size_t cwords = (size_t)P.nbins_pad + 8;                       // bin counts, overflow count, two work counters
  cudaError_t e = cudaMallocAsync(&P.entries, (size_t)P.nbins * P.cap * 8, ctx->stream);
  if (e == cudaSuccess) e = cudaMallocAsync(&P.where, nwhere * 4, ctx->stream);
  if (e == cudaSuccess) e = cudaMallocAsync(&P.runs, (size_t)nchunks * P.nbins_pad * sizeof(uint2), ctx->stream);
  if (e == cudaSuccess) e = cudaMallocAsync(&ctl, cwords * 8, ctx->stream);
  if (e != cudaSuccess) {
    // no room for the workspace: not an error, the direct scan needs none
    cudaGetLastError();
    if (P.entries) cudaFreeAsync(P.entries, ctx->stream);
    if (P.where) cudaFreeAsync(P.where, ctx->stream);
    if (P.runs) cudaFreeAsync(P.runs, ctx->stream);
    return MFN_OK;
  }
  MFN_CUDA(cudaMemsetAsync(ctl, 0, cwords * 8, ctx->stream));
  P.bin_count = ctl; P.over_n = ctl + P.nbins_pad; P.counter = ctl + P.nbins_pad + 1;
  P.over_cap = 0;
  const bool geo1 = J.fs_d == 1 && P.lowbits == 34 && P.lbits + P.tbits == 34 && P.tbits == J.fs_t && P.nbins == (1u << (2 * J.k - 34));
  const bool geo2 = J.fs_d == 1 && P.lowbits == 33 && P.lbits + P.tbits == 33 && P.tbits == J.fs_t && P.nbins == (1u << (2 * J.k - 33));
  const size_t smem3 = s3_smem_bytes(P.nbins_pad, tiles);
  auto launch3 = [&](auto kernel) -> int {
    int per_sm = 0;
    MFN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, (int)threads3, smem3);
    const int grid = (int)std::min<uint64_t>((uint64_t)ctx->sm_count * std::max(per_sm, 1), nchunks);
    kernel<<<grid, threads3, smem3, ctx->stream>>>(P);
    return MFN_OK;
  };
  const int geo = geo1 ? 1 : geo2 ? 2 : 0;
  if (tiles == 2) {
    if (geo == 1) MFN_TRY(launch3(k_sweep_scatter3<21, 1, 2, true>)); else if (geo == 2) MFN_TRY(launch3(k_sweep_scatter3<21, 2, 2, true>)); else MFN_TRY(launch3(k_sweep_scatter3<21, 0, 2, true>));
  } else {
    if (geo == 1) MFN_TRY(launch3(k_sweep_scatter3<21, 1, 4, true>)); else if (geo == 2) MFN_TRY(launch3(k_sweep_scatter3<21, 2, 4, true>)); else MFN_TRY(launch3(k_sweep_scatter3<21, 0, 4, true>));
  }
  ctx->launches++;
  // MFN_DUMP_DEBUG=1: milliseconds of the three passes to stderr
  const bool dbg = getenv("MFN_DUMP_DEBUG") != nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  if (dbg) { for (auto &x : ev) cudaEventCreate(&x); cudaEventRecord(ev[1], ctx->stream); }
  // did every run fit its bin?  (the outputs are only written by the last pass)
  unsigned long long over = 0;
  MFN_CUDA(cudaMemcpyAsync(&over, P.over_n, 8, cudaMemcpyDeviceToHost, ctx->stream));
  MFN_CUDA(cudaStreamSynchronize(ctx->stream));
  if (over == 0) {
    P.counter = ctl + P.nbins_pad + 2;
    mfn_sweep_dump_launch_probe(P, ctx->sm_count, ctx->stream);
    if (dbg) cudaEventRecord(ev[2], ctx->stream);
    mfn_sweep_dump_launch_gather(P, tiles, nchunks, tile_begin * MFN_TILE, tile_end * MFN_TILE, d_vr, d_va, d_valid, ctx->stream);
    ctx->launches += 2;
    *ok = 1;
    if (dbg) {
      cudaEventRecord(ev[3], ctx->stream);
      cudaStreamSynchronize(ctx->stream);
      float b = 0, c = 0;
      cudaEventElapsedTime(&b, ev[1], ev[2]); cudaEventElapsedTime(&c, ev[2], ev[3]);
      fprintf(stderr, "[dump sweep] %llu windows, %u bins, lowbits %u, %d-tile chunks: probe %.2f ms, gather %.2f ms (scatter = the rest of the call)\n",
              (unsigned long long)windows, P.nbins, P.lowbits, tiles, b, c);
    }
  }
  if (dbg) for (auto &x : ev) if (x) cudaEventDestroy(x);
  cudaError_t e2 = cudaGetLastError();
  cudaFreeAsync(P.entries, ctx->stream);
  cudaFreeAsync(P.where, ctx->stream);
  cudaFreeAsync(P.runs, ctx->stream);
  cudaFreeAsync(ctl, ctx->stream);
  if (e2 != cudaSuccess) { mfn_set_error("dump sweep: %s", cudaGetErrorString(e2)); return MFN_ERR_CUDA; }
  return MFN_OK;
}

// ctx->stream waits for whatever mfn_sweep_finish(..., true) put on the second compute stream
void mfn_sweep_join(mfn_ctx *ctx) {
  cudaEventRecord(ctx->ev_probe[1], ctx->probe_stream);
  cudaStreamWaitEvent(ctx->stream, ctx->ev_probe[1], 0);
}
