// Host side of the scan kernels: launches, accumulators, fetch.
#include "scan_kernel.cuh"

#include <algorithm>

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------

static int check_tables(const mfn_table *R, const mfn_table *A, bool per_position = false) {
  if (!R || !A) return MFN_ERR_INVALID;
  if (per_position && R->streamed) { mfn_set_error("this read table was streamed (mfn_tables_restrict_stream): -dump and path scoring need the database resident (mfn_table_load)"); return MFN_ERR_UNSUPPORTED; }
  if (R->dev.k != A->dev.k) { mfn_set_error("mer size mismatch: reads k=%u, assembly k=%u", R->dev.k, A->dev.k); return MFN_ERR_K_MISMATCH; }
  return MFN_OK;
}

// tile range [begin, end) of shard s out of n
static void shard_tiles(uint64_t ntiles, uint32_t s, uint32_t n, uint64_t &b, uint64_t &e) {
  b = ntiles * s / n;
  e = ntiles * (s + 1) / n;
}

struct ScanScratch { unsigned long long *counter = nullptr; unsigned long long *zmap = nullptr; };

int mfn_scan_grid_0(mfn_ctx *); int mfn_scan_grid_1(mfn_ctx *); int mfn_scan_grid_2(mfn_ctx *); int mfn_scan_grid_3(mfn_ctx *);
void mfn_scan_launch_0(const ScanParams &, int, cudaStream_t); void mfn_scan_launch_1(const ScanParams &, int, cudaStream_t);
void mfn_scan_launch_2(const ScanParams &, int, cudaStream_t); void mfn_scan_launch_3(const ScanParams &, int, cudaStream_t);

int mfn_scan_grid(mfn_ctx *ctx, int mode) {
  switch (mode) {
    case 0: return mfn_scan_grid_0(ctx);
    case 1: return mfn_scan_grid_1(ctx);
    case 2: return mfn_scan_grid_2(ctx);
    default: return mfn_scan_grid_3(ctx);
  }
}

void mfn_scan_launch(const ScanParams &P, int mode, int grid, cudaStream_t stream) {
  switch (mode) {
    case 0: mfn_scan_launch_0(P, grid, stream); break;
    case 1: mfn_scan_launch_1(P, grid, stream); break;
    case 2: mfn_scan_launch_2(P, grid, stream); break;
    default: mfn_scan_launch_3(P, grid, stream); break;
  }
}

// The fused table is used whenever it has been built for exactly this pair of databases.
static bool use_joint(const mfn_table *R, const mfn_table *A, ScanParams &P) {
  const char *env = getenv("MFN_NO_JOINT");
  if (env && atoi(env)) return false;
  if (!R->joint || R->joint->asm_ != A) return false;
  P.J = R->joint->dev;
  return true;
}

static int launch_scan(mfn_ctx *ctx, ScanParams &P, bool dump, bool joint) {
  if (P.tile_begin >= P.tile_end) return MFN_OK;
  unsigned long long *counter;
  MFN_CUDA(cudaMallocAsync(&counter, 8, ctx->stream));
  MFN_CUDA(cudaMemsetAsync(counter, 0, 8, ctx->stream));
  P.tile_counter = counter;
  const int mode = (dump ? SCAN_DUMP : 0) | (joint ? SCAN_JOINT : 0);
  int grid = (int)std::min<uint64_t>((uint64_t)mfn_scan_grid(ctx, mode), P.tile_end - P.tile_begin);
  mfn_scan_launch(P, mode, grid, ctx->stream);
  ctx->launches++;
  MFN_CUDA(cudaGetLastError());
  MFN_CUDA(cudaFreeAsync(counter, ctx->stream));
  return MFN_OK;
}

// bitmap of read multiplicities whose readK is 0 under the -prob table (merfin-globals.C:93-97)
int mfn_build_zmap(mfn_ctx *ctx, const mfn_kmodel *km, unsigned long long **d_zmap, uint32_t *zlen, uint32_t *zhi) {
  *d_zmap = nullptr; *zlen = 0; *zhi = 0;
  if (!km || km->nprob == 0) return MFN_OK;
  uint32_t n = km->nprob + 1;                    // values 0 .. nprob
  std::vector<unsigned long long> bits((n + 63) / 64, 0);
  bits[0] |= 1;                                  // v == 0
  for (uint32_t v = 1; v <= km->nprob; v++)
    if (km->probK[v - 1] == 0) { bits[v >> 6] |= 1ull << (v & 63); *zhi = v; }
  MFN_CUDA(cudaMallocAsync(d_zmap, bits.size() * 8, ctx->stream));
  MFN_CUDA(cudaMemcpyAsync(*d_zmap, bits.data(), bits.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  MFN_CUDA(cudaStreamSynchronize(ctx->stream));  // 'bits' is a local
  *zlen = n;
  return MFN_OK;
}

int mfn_hist_tiles(mfn_ctx *ctx, const mfn_table *reads, const mfn_table *asm_, const mfn_seq *seq,
                   const mfn_kmodel *km, uint64_t tile_begin, uint64_t tile_end, mfn_accum *acc) {
  if (!ctx || !seq || !acc) return MFN_ERR_INVALID;
  MFN_TRY(check_tables(reads, asm_));
  if (acc->nseq < seq->nseq) { mfn_set_error("accumulator has %u record slots, sequence set has %u", acc->nseq, seq->nseq); return MFN_ERR_INVALID; }
  MFN_CUDA(cudaSetDevice(ctx->device));
  if (reads->wide || asm_->wide) return mfn_wide_hist_tiles(ctx, reads, asm_, seq, km, tile_begin, tile_end, acc);     // k > 32
  ScanParams P{};
  P.R = reads->dev; P.A = asm_->dev;
  P.seq2 = seq->d_seq2; P.ok = seq->d_ok; P.rec_start = seq->d_rec_start; P.nseq = seq->nseq;
  P.tile_begin = tile_begin; P.tile_end = std::min(tile_end, seq->ntiles);
  P.acc = acc->d_buf; P.sparse = acc->sparse;
  unsigned long long *zmap = nullptr;
  uint32_t zhi = 0;
  MFN_TRY(mfn_build_zmap(ctx, km, &zmap, &P.zlen, &zhi));
  P.zmap = zmap;
  int st;
  const bool sweep = mfn_sweep_applicable(reads, asm_, P.tile_end > P.tile_begin ? P.tile_end - P.tile_begin : 0, seq->nseq);
  if (reads->streamed && !sweep) {
    if (zmap) cudaFreeAsync(zmap, ctx->stream);
    mfn_set_error("this read table was streamed (mfn_tables_restrict_stream): it serves -hist of the sequences it was restricted to, "
                  "with this assembly table, through the range-partitioned scan only");
    return MFN_ERR_UNSUPPORTED;
  }
  if (sweep)
    st = mfn_sweep_hist(ctx, reads, asm_, seq, km, P.tile_begin, P.tile_end, acc, zmap, P.zlen, zhi);
  else
    st = launch_scan(ctx, P, false, use_joint(reads, asm_, P));
  if (zmap) cudaFreeAsync(zmap, ctx->stream);
  return st;
}

extern "C" int mfn_hist_resident(mfn_ctx *ctx, const mfn_table *reads, const mfn_table *asm_, const mfn_seq *seq,
                                 const mfn_kmodel *km, uint32_t shard, uint32_t nshards, mfn_accum *acc) {
  if (!seq || nshards == 0 || shard >= nshards) return MFN_ERR_INVALID;
  uint64_t tb, te;
  shard_tiles(seq->ntiles, shard, nshards, tb, te);
  return mfn_hist_tiles(ctx, reads, asm_, seq, km, tb, te, acc);
}

extern "C" int mfn_dump_resident(mfn_ctx *ctx, const mfn_table *reads, const mfn_table *asm_, const mfn_seq *seq,
                                 uint32_t *d_v_read, uint32_t *d_v_asm, uint8_t *d_valid) {
  if (!ctx || !seq || !d_v_read || !d_v_asm || !d_valid) return MFN_ERR_INVALID;
  MFN_TRY(check_tables(reads, asm_, true));
  MFN_CUDA(cudaSetDevice(ctx->device));
  if (reads->wide || asm_->wide) return mfn_wide_dump(ctx, reads, asm_, seq, d_v_read, d_v_asm, d_valid);              // k > 32
  ScanParams P{};
  P.R = reads->dev; P.A = asm_->dev;
  P.seq2 = seq->d_seq2; P.ok = seq->d_ok; P.rec_start = seq->d_rec_start; P.nseq = seq->nseq;
  // big batches over a fused table: the range-partitioned path (sweep_dump.cu); the direct scan does the rest
  uint64_t done_tiles = 0;
  MFN_TRY(mfn_sweep_dump(ctx, reads, asm_, seq, d_v_read, d_v_asm, d_valid, &done_tiles));
  // the range-partitioned path writes every position of its tiles; the direct scan relies on zeroed outputs
  {
    const uint64_t g = std::min(done_tiles * MFN_TILE, seq->G);
    const uint64_t rec = (uint64_t)(std::upper_bound(seq->rec_start.begin(), seq->rec_start.end(), g) - seq->rec_start.begin()) - 1;
    // first caller position not written yet: g - rec, or the next record's first base when g is a separator slot or beyond
    uint64_t s0 = rec >= seq->nseq ? seq->total_bases : std::min<uint64_t>(g - rec, seq->rec_start[rec + 1] - 1 - rec);
    if (done_tiles == 0) s0 = 0;
    if (s0 < seq->total_bases) {
      MFN_CUDA(cudaMemsetAsync(d_v_read + s0, 0, (seq->total_bases - s0) * 4, ctx->stream));
      MFN_CUDA(cudaMemsetAsync(d_v_asm + s0, 0, (seq->total_bases - s0) * 4, ctx->stream));
      MFN_CUDA(cudaMemsetAsync(d_valid + s0, 0, seq->total_bases - s0, ctx->stream));
    }
  }
  P.tile_begin = done_tiles; P.tile_end = seq->ntiles;
  P.out_vr = d_v_read; P.out_va = d_v_asm; P.out_valid = d_valid;
  return launch_scan(ctx, P, true, use_joint(reads, asm_, P));
}

// ---------------------------------------------------------------------------
// accumulators
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sparse_clear(PairTable t) {
  if (*t.local_n == 0) return;                 // nothing was inserted since the last clear (the usual case)
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i <= t.mask; i += (uint64_t)gridDim.x * blockDim.x) {
    t.keys[i] = MFN_EMPTY64;
    t.counts[i] = 0;
  }
}

extern "C" int mfn_accum_create(mfn_ctx *ctx, uint32_t nseq, mfn_accum **out) {
  if (!ctx || !out) return MFN_ERR_INVALID;
  MFN_CUDA(cudaSetDevice(ctx->device));
  mfn_accum *a = new mfn_accum();
  a->ctx = ctx; a->device = ctx->device; a->nseq = nseq;
  a->nwords = MFN_ACC_HEADER + (uint64_t)MFN_DENSE_R * MFN_DENSE_A + 2ull * nseq;
  a->sparse.mask = MFN_SPARSE_CAP - 1;
  cudaError_t e = cudaMalloc(&a->d_buf, a->nwords * 8);
  if (e == cudaSuccess) e = cudaMallocHost(&a->h_buf, a->nwords * 8);
  if (e == cudaSuccess) e = cudaMalloc(&a->sparse.local_n, 8);
  if (e == cudaSuccess) e = cudaMalloc(&a->sparse.keys, (size_t)MFN_SPARSE_CAP * 8);
  if (e == cudaSuccess) e = cudaMalloc(&a->sparse.counts, (size_t)MFN_SPARSE_CAP * 8);
  if (e != cudaSuccess) { mfn_set_error("cudaMalloc(accumulator): %s", cudaGetErrorString(e)); mfn_accum_free(a); return MFN_ERR_NOMEM; }
  // first clear: unconditional
  if (e == cudaSuccess) e = cudaMemsetAsync(a->sparse.keys, 0xff, (size_t)MFN_SPARSE_CAP * 8, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(a->sparse.counts, 0, (size_t)MFN_SPARSE_CAP * 8, ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(a->sparse.local_n, 0, 8, ctx->stream);
  if (e != cudaSuccess) { mfn_set_error("accumulator: %s", cudaGetErrorString(e)); mfn_accum_free(a); return MFN_ERR_CUDA; }
  int st = mfn_accum_reset(ctx, a);
  if (st != MFN_OK) { mfn_accum_free(a); return st; }
  *out = a;
  return MFN_OK;
}

// dst[i] += src[i] over the flat accumulator buffer (src: a staging copy on dst's device)
__global__ void __launch_bounds__(256) k_accum_add(unsigned long long *dst, const unsigned long long *src, uint64_t n) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const unsigned long long v = src[i];
    if (v) dst[i] += v;
  }
}

// One process driving several devices (the CLI's -devices): add the flat buffer of `src` (another
// device's accumulator) into `dst`.  The copy goes device to device (NVLink when the devices are
// peers); the caller has synchronised src's ctx.  Equivalent to one step of the all-reduce that
// one-process-per-GPU callers run with NCCL.
extern "C" int mfn_accum_add_from(mfn_ctx *ctx, mfn_accum *dst, const mfn_accum *src) {
  if (!ctx || !dst || !src || dst->nwords != src->nwords) return MFN_ERR_INVALID;
  MFN_CUDA(cudaSetDevice(ctx->device));
  unsigned long long *tmp = nullptr;
  MFN_CUDA(cudaMallocAsync(&tmp, dst->nwords * 8, ctx->stream));
  MFN_CUDA(cudaMemcpyPeerAsync(tmp, ctx->device, src->d_buf, src->ctx->device, dst->nwords * 8, ctx->stream));
  k_accum_add<<<ctx->sm_count * 2, 256, 0, ctx->stream>>>(dst->d_buf, tmp, dst->nwords);
  ctx->launches++;
  MFN_CUDA(cudaGetLastError());
  MFN_CUDA(cudaFreeAsync(tmp, ctx->stream));
  return MFN_OK;
}

extern "C" int mfn_accum_reset(mfn_ctx *ctx, mfn_accum *a) {
  if (!ctx || !a) return MFN_ERR_INVALID;
  // the 64 MB side table is cleared only when something was put into it (its counter word tells, on the device)
  MFN_CUDA(cudaMemsetAsync(a->d_buf, 0, a->nwords * 8, ctx->stream));
  k_sparse_clear<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(a->sparse);
  MFN_CUDA(cudaMemsetAsync(a->sparse.local_n, 0, 8, ctx->stream));
  return MFN_OK;
}

extern "C" void mfn_accum_free(mfn_accum *a) {
  if (!a) return;
  cudaSetDevice(a->device);
  cudaFree(a->d_buf); cudaFree(a->sparse.keys); cudaFree(a->sparse.counts); cudaFree(a->sparse.local_n);
  if (a->h_buf) cudaFreeHost(a->h_buf);
  delete a;
}

extern "C" int mfn_accum_device_buffer(const mfn_accum *a, void **dptr, uint64_t *nwords) {
  if (!a) return MFN_ERR_INVALID;
  if (dptr) *dptr = a->d_buf;
  if (nwords) *nwords = a->nwords;
  return MFN_OK;
}

// ---- accumulator layout (host side, no GPU needed) --------------------------------------------
extern "C" int mfn_accum_layout(uint32_t nseq, uint32_t *header_words, uint32_t *dense_r, uint32_t *dense_a,
                                uint64_t *nwords) {
  if (header_words) *header_words = MFN_ACC_HEADER;
  if (dense_r) *dense_r = MFN_DENSE_R;
  if (dense_a) *dense_a = MFN_DENSE_A;
  if (nwords) *nwords = MFN_ACC_HEADER + (uint64_t)MFN_DENSE_R * MFN_DENSE_A + 2ull * nseq;
  return MFN_OK;
}

extern "C" int mfn_shard_tiles(uint64_t total_bases, uint32_t nseq, uint32_t shard, uint32_t nshards,
                               uint64_t *g_begin, uint64_t *g_end) {
  if (nshards == 0 || shard >= nshards || !g_begin || !g_end) return MFN_ERR_INVALID;
  uint64_t G = total_bases + nseq, ntiles = std::max<uint64_t>(1, ceil_div_u64(G, MFN_TILE)), b, e;
  shard_tiles(ntiles, shard, nshards, b, e);
  *g_begin = b * MFN_TILE; *g_end = std::min<uint64_t>(e * MFN_TILE, std::max<uint64_t>(G, 1));
  if (shard + 1 == nshards) *g_end = G;
  return MFN_OK;
}

// Dense tier and per-record counters of a copy of the flat accumulator buffer.
static int decode_dense(const uint64_t *h, uint32_t nseq, mfn_seq_counts *per_seq, mfn_pair_count *pairs, uint64_t cap,
                        uint64_t *npairs, bool *full_out) {
  if (h[2]) { mfn_set_error("joint-histogram side table overflowed (%llu pairs lost)", (unsigned long long)h[2]); return MFN_ERR_CAPACITY; }
  uint64_t n = 0;
  bool full = false;
  const uint64_t *d = h + MFN_ACC_HEADER;
  for (uint32_t va = 0; va < MFN_DENSE_A; va++)
    for (uint32_t vr = 0; vr < MFN_DENSE_R; vr++) {
      uint64_t c = d[(uint64_t)va * MFN_DENSE_R + vr];
      if (!c) continue;
      if (pairs && n < cap) { pairs[n].v_read = vr; pairs[n].v_asm = va; pairs[n].count = c; } else full = true;
      n++;
    }
  if (h[3]) {
    if (pairs && n < cap) { pairs[n].v_read = 0xffffffffu; pairs[n].v_asm = 0xffffffffu; pairs[n].count = h[3]; } else full = true;
    n++;
  }
  if (per_seq)
    for (uint32_t i = 0; i < nseq; i++) {
      per_seq[i].kasm = d[(uint64_t)MFN_DENSE_R * MFN_DENSE_A + 2 * i];
      per_seq[i].kmissing = d[(uint64_t)MFN_DENSE_R * MFN_DENSE_A + 2 * i + 1];
    }
  *npairs = n;
  *full_out = full;
  return MFN_OK;
}

// Decode a (possibly all-reduced) copy of the flat accumulator buffer.  Pairs outside the dense tier
// are NOT in the buffer (header word [1] counts them over all ranks): a buffer that has any is refused
// here -- gather them and use mfn_accum_decode_gathered.
extern "C" int mfn_accum_decode(const uint64_t *h, uint32_t nseq, mfn_seq_counts *per_seq,
                                mfn_pair_count *pairs, uint64_t cap, uint64_t *npairs) {
  if (!h || !npairs) return MFN_ERR_INVALID;
  if (h[1]) {
    mfn_set_error("%llu (v_read, v_asm) pairs outside the dense tier live in rank-local side tables that the all-reduce does not carry: "
                  "gather them with mfn_accum_sparse_fetch on every rank and decode with mfn_accum_decode_gathered",
                  (unsigned long long)h[1]);
    return MFN_ERR_INVALID;
  }
  bool full = false;
  MFN_TRY(decode_dense(h, nseq, per_seq, pairs, cap, npairs, &full));
  return (pairs && full) ? MFN_ERR_CAPACITY : MFN_OK;
}

extern "C" int mfn_accum_decode_gathered(const uint64_t *h, uint32_t nseq, const mfn_pair_count *sparse, uint64_t n_sparse,
                                         mfn_seq_counts *per_seq, mfn_pair_count *pairs, uint64_t cap, uint64_t *npairs) {
  if (!h || !npairs || (n_sparse && !sparse)) return MFN_ERR_INVALID;
  if (h[1] != n_sparse) {
    mfn_set_error("the accumulator header counts %llu side-table pairs over all ranks, %llu were supplied",
                  (unsigned long long)h[1], (unsigned long long)n_sparse);
    return MFN_ERR_INVALID;
  }
  uint64_t n = 0;
  bool full = false;
  MFN_TRY(decode_dense(h, nseq, per_seq, pairs, cap, &n, &full));
  // equal pairs of different ranks are summed; ascending (v_read, v_asm), as the single-rank fetch lists them
  std::vector<std::pair<unsigned long long, unsigned long long>> v;
  v.reserve(n_sparse);
  for (uint64_t i = 0; i < n_sparse; i++) v.emplace_back(((unsigned long long)sparse[i].v_read << 32) | sparse[i].v_asm, sparse[i].count);
  std::sort(v.begin(), v.end());
  for (size_t i = 0; i < v.size();) {
    size_t j = i;
    unsigned long long c = 0;
    while (j < v.size() && v[j].first == v[i].first) c += v[j++].second;
    if (pairs && n < cap) { pairs[n].v_read = (uint32_t)(v[i].first >> 32); pairs[n].v_asm = (uint32_t)v[i].first; pairs[n].count = c; } else full = true;
    n++;
    i = j;
  }
  *npairs = n;
  return (pairs && full) ? MFN_ERR_CAPACITY : MFN_OK;
}

// D2H of the flat buffer into the accumulator's pinned mirror
static int read_flat(mfn_ctx *ctx, const mfn_accum *a) {
  MFN_CUDA(cudaSetDevice(ctx->device));
  MFN_CUDA(cudaMemcpyAsync(a->h_buf, a->d_buf, a->nwords * 8, cudaMemcpyDeviceToHost, ctx->stream));
  MFN_CUDA(cudaStreamSynchronize(ctx->stream));
  return MFN_OK;
}

extern "C" int mfn_accum_read(mfn_ctx *ctx, const mfn_accum *a, uint64_t *host_buf) {
  if (!ctx || !a || !host_buf) return MFN_ERR_INVALID;
  MFN_TRY(read_flat(ctx, a));
  memcpy(host_buf, a->h_buf, a->nwords * 8);
  return MFN_OK;
}

// this rank's side table of pairs outside the dense tier, ascending (v_read, v_asm)
static int read_sparse(mfn_ctx *ctx, const mfn_accum *a, std::vector<std::pair<unsigned long long, unsigned long long>> &v) {
  std::vector<unsigned long long> keys(MFN_SPARSE_CAP), counts(MFN_SPARSE_CAP);
  MFN_CUDA(cudaMemcpyAsync(keys.data(), a->sparse.keys, (size_t)MFN_SPARSE_CAP * 8, cudaMemcpyDeviceToHost, ctx->stream));
  MFN_CUDA(cudaMemcpyAsync(counts.data(), a->sparse.counts, (size_t)MFN_SPARSE_CAP * 8, cudaMemcpyDeviceToHost, ctx->stream));
  MFN_CUDA(cudaStreamSynchronize(ctx->stream));
  for (uint32_t i = 0; i < MFN_SPARSE_CAP; i++) if (keys[i] != MFN_EMPTY64) v.emplace_back(keys[i], counts[i]);
  std::sort(v.begin(), v.end());
  return MFN_OK;
}

extern "C" int mfn_accum_sparse_fetch(mfn_ctx *ctx, const mfn_accum *a, mfn_pair_count *pairs, uint64_t cap, uint64_t *npairs) {
  if (!ctx || !a || !npairs) return MFN_ERR_INVALID;
  MFN_CUDA(cudaSetDevice(ctx->device));
  // the count of entries this rank inserted: one word, read before the (32 MB) table itself
  unsigned long long any = 0;
  MFN_CUDA(cudaMemcpyAsync(&any, a->sparse.local_n, 8, cudaMemcpyDeviceToHost, ctx->stream));
  MFN_CUDA(cudaStreamSynchronize(ctx->stream));
  *npairs = any;
  if (!any) return MFN_OK;
  if (!pairs || any > cap) return MFN_ERR_CAPACITY;
  std::vector<std::pair<unsigned long long, unsigned long long>> v;
  MFN_TRY(read_sparse(ctx, a, v));
  if (v.size() != any) { mfn_set_error("side table holds %zu pairs, its counter says %llu", v.size(), any); return MFN_ERR_CUDA; }
  for (size_t i = 0; i < v.size(); i++) { pairs[i].v_read = (uint32_t)(v[i].first >> 32); pairs[i].v_asm = (uint32_t)v[i].first; pairs[i].count = v[i].second; }
  return MFN_OK;
}

extern "C" int mfn_accum_fetch(mfn_ctx *ctx, const mfn_accum *a, int include_dense, int include_sparse,
                               mfn_seq_counts *per_seq, mfn_pair_count *pairs, uint64_t cap, uint64_t *npairs) {
  if (!ctx || !a || !npairs) return MFN_ERR_INVALID;
  MFN_TRY(read_flat(ctx, a));
  const unsigned long long *h = a->h_buf;
  uint64_t n = 0;
  bool full = false;
  if (h[2]) { mfn_set_error("joint-histogram side table overflowed (%llu pairs lost)", h[2]); return MFN_ERR_CAPACITY; }
  if (include_dense) MFN_TRY(decode_dense((const uint64_t *)h, a->nseq, per_seq, pairs, cap, &n, &full));
  auto push = [&](uint32_t vr, uint32_t va, uint64_t c) {
    if (pairs && n < cap) { pairs[n].v_read = vr; pairs[n].v_asm = va; pairs[n].count = c; } else full = true;
    n++;
  };
  if (include_sparse && h[1]) {
    std::vector<std::pair<unsigned long long, unsigned long long>> v;
    MFN_TRY(read_sparse(ctx, a, v));
    // a buffer that was all-reduced counts the side-table pairs of EVERY rank; this call only has its own
    if (include_dense && v.size() != h[1]) {
      mfn_set_error("the accumulator header counts %llu side-table pairs, this rank holds %zu: the buffer was all-reduced -- "
                    "gather with mfn_accum_sparse_fetch and decode with mfn_accum_decode_gathered", h[1], v.size());
      return MFN_ERR_INVALID;
    }
    for (auto &kv : v) push((uint32_t)(kv.first >> 32), (uint32_t)kv.first, kv.second);
  }
  *npairs = n;
  return (pairs && full) ? MFN_ERR_CAPACITY : MFN_OK;
}
