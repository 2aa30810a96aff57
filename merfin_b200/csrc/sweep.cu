// Range-partitioned -hist scan: host side and the two main passes (see sweep_kernels.cuh).
#include "sweep_kernels.cuh"

// ------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------
static uint32_t bits_of(uint64_t x) { uint32_t n = 0; while (x) { n++; x >>= 1; } return n; }

// Can (and should) this scan use the sweep?  Needs the joint table and h low bits that fit 32 bits.
bool mfn_sweep_applicable(const mfn_table *R, const mfn_table *A, uint64_t ntiles) {
  const char *env = getenv("MFN_SWEEP");
  if (env && atoi(env) == 0) return false;
  if (!R->joint || R->joint->asm_ != A) return false;
  const JointDev &J = R->joint->dev;
  if (2 * J.k > 32 + 11 || J.k < 6) return false;            // low bits of h must fit 32 bits; >= 256 bins
  if (env && atoi(env) == 1) return true;
  // worth it when the windows revisit the table lines: >= 1 window per 4 lines, and a big batch
  const uint64_t windows = ntiles * MFN_TILE;
  return windows >= (1ull << 24) && windows * 4 >= J.nlines;
}

int mfn_sweep_hist(mfn_ctx *ctx, const mfn_table *R, const mfn_table *A, const mfn_seq *seq, const mfn_kmodel *km,
                   uint64_t tile_begin, uint64_t tile_end, mfn_accum *acc, const unsigned long long *zmap, uint32_t zlen) {
  (void)km; (void)A;
  if (tile_begin >= tile_end) return MFN_OK;
  const JointDev &J = R->joint->dev;
  SweepParams P{};
  P.J = J;
  P.seq2 = seq->d_seq2; P.ok = seq->d_ok; P.rec_start = seq->d_rec_start; P.nseq = seq->nseq;
  P.tile_begin = tile_begin; P.tile_end = tile_end;
  // bins: table slice per bin <= ~24 MB, low bits of h must fit 32 bits, nbins multiple of 256
  const uint64_t table_bytes = J.nlines * 128;
  uint32_t b = bits_of(table_bytes / (24ull << 20));
  const uint32_t bmin = 2 * J.k > 32 ? 2 * J.k - 32 : 0;
  b = std::max(std::max(b, bmin), 8u);
  b = std::min(b, std::min(11u, 2 * J.k));
  if (2 * J.k - b > 32) { mfn_set_error("sweep: k=%u needs more than 2048 bins", J.k); return MFN_ERR_UNSUPPORTED; }
  P.bbits = b; P.lowbits = 2 * J.k - b; P.nbins = 1u << b;
  const uint64_t windows = (tile_end - tile_begin) * MFN_TILE;
  P.cap = ((uint64_t)((double)windows / P.nbins * 1.15) + 2048 + 31) & ~31ull;
  P.over_cap = std::max<uint64_t>(1 << 20, windows / 16);
  P.acc = acc->d_buf; P.sparse = acc->sparse; P.zmap = zmap; P.zlen = zlen;

  // workspace (stream-ordered pool: reused across steps without cudaMalloc cost)
  const size_t ebytes = (size_t)P.nbins * P.cap * 8;
  unsigned long long *ctl = nullptr;
  MFN_CUDA(cudaMallocAsync(&P.entries, ebytes, ctx->stream));
  MFN_CUDA(cudaMallocAsync(&ctl, (2 * (size_t)P.nbins + 4) * 8, ctx->stream));
  MFN_CUDA(cudaMallocAsync(&P.over_h, P.over_cap * 8, ctx->stream));
  MFN_CUDA(cudaMallocAsync(&P.over_rec, P.over_cap * 4, ctx->stream));
  P.bin_count = ctl; P.bin_valid = ctl + P.nbins; P.over_n = ctl + 2 * P.nbins;
  unsigned long long *counters = ctl + 2 * P.nbins + 1;       // [3]
  MFN_CUDA(cudaMemsetAsync(ctl, 0, (2 * (size_t)P.nbins + 4) * 8, ctx->stream));
  MFN_CUDA(cudaMemsetAsync(P.bin_valid, 0xff, (size_t)P.nbins * 8, ctx->stream));

  int per_sm = 0;
  const size_t sc_smem = (2 * (size_t)SW_CHUNK + 3 * (size_t)SW_MAX_BINS) * 4;
  MFN_CUDA(cudaFuncSetAttribute(k_sweep_scatter<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc_smem));
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sweep_scatter<0>, SW_THREADS, sc_smem);
  P.counter = counters;
  k_sweep_scatter<0><<<ctx->sm_count * std::max(per_sm, 1), SW_THREADS, sc_smem, ctx->stream>>>(P);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sweep_probe<false>, SW_THREADS, 0);
  P.counter = counters + 1;
  k_sweep_probe<false><<<ctx->sm_count * std::max(per_sm, 1), SW_THREADS, 0, ctx->stream>>>(P);
  P.counter = counters + 2;
  mfn_sweep_launch_over(P, ctx->sm_count, ctx->stream);
  ctx->launches += 3;
  MFN_CUDA(cudaGetLastError());
  // the side list must not have overflowed: checked lazily at fetch through the header
  MFN_CUDA(cudaFreeAsync(P.entries, ctx->stream));
  MFN_CUDA(cudaFreeAsync(ctl, ctx->stream));
  MFN_CUDA(cudaFreeAsync(P.over_h, ctx->stream));
  MFN_CUDA(cudaFreeAsync(P.over_rec, ctx->stream));
  return MFN_OK;
}
