// C ABI glue: context, errors, the host-buffer entry points and -completeness.
#include "scan_kernel.cuh"

#include <algorithm>
#include <map>
#include <mutex>


// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local char g_err[1024] = "";

void mfn_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

extern "C" const char *mfn_last_error(void) { return g_err; }
extern "C" const char *mfn_version(void) { return "merfin_b200 0.1 (sm_100a)"; }

extern "C" const char *mfn_strerror(int s) {
  switch (s) {
    case MFN_OK: return "ok";
    case MFN_ERR_NO_DEVICE: return "no usable CUDA device (this library has no CPU path)";
    case MFN_ERR_CUDA: return "CUDA error";
    case MFN_ERR_IO: return "I/O error";
    case MFN_ERR_FORMAT: return "not a meryl database / corrupt data";
    case MFN_ERR_UNSUPPORTED: return "unsupported input";
    case MFN_ERR_INVALID: return "invalid argument";
    case MFN_ERR_CAPACITY: return "buffer or table capacity exceeded";
    case MFN_ERR_K_MISMATCH: return "mer size mismatch";
    case MFN_ERR_NOMEM: return "out of memory";
  }
  return "unknown error";
}

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
extern "C" int mfn_ctx_create(mfn_ctx **out, int device) {
  if (!out) return MFN_ERR_INVALID;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    mfn_set_error("no CUDA device: %s", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    return MFN_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= n) { mfn_set_error("device %d out of range (0..%d)", device, n - 1); return MFN_ERR_INVALID; }
  MFN_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  MFN_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) { mfn_set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor); return MFN_ERR_NO_DEVICE; }
  mfn_ctx *c = new mfn_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  MFN_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  MFN_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  MFN_CUDA(cudaStreamCreateWithFlags(&c->pack_stream, cudaStreamNonBlocking));
  MFN_CUDA(cudaStreamCreateWithFlags(&c->probe_stream, cudaStreamNonBlocking));
  for (auto &ev : c->ev_probe) MFN_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  for (auto &ev : c->ev) MFN_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  c->stream = c->own_stream;
  // keep stream-ordered workspace (cudaMallocAsync) cached between calls
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  *out = c;
  return MFN_OK;
}

extern "C" void mfn_ctx_destroy(mfn_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  mfn_ingest_free(c->ingest);
  mfn_seq_free(c->scratch_seq);
  mfn_accum_free(c->scratch_acc);
  for (int i = 0; i < MFN_STAGE_BUFS; i++) cudaFree(c->d_ascii[i]);
  if (c->h_pinned) cudaFreeHost(c->h_pinned);
  for (auto &ev : c->ev) if (ev) cudaEventDestroy(ev);
  for (auto &ev : c->pass_ev) cudaEventDestroy(ev);
  cudaStreamDestroy(c->own_stream);
  cudaStreamDestroy(c->copy_stream);
  cudaStreamDestroy(c->pack_stream);
  cudaStreamDestroy(c->probe_stream);
  for (auto &ev : c->ev_probe) if (ev) cudaEventDestroy(ev);
  delete c;
}

extern "C" int mfn_ctx_device(const mfn_ctx *c) { return c ? c->device : -1; }

extern "C" int mfn_ctx_set_stream(mfn_ctx *c, void *stream) {
  if (!c) return MFN_ERR_INVALID;
  c->stream = stream ? (cudaStream_t)stream : c->own_stream;
  return MFN_OK;
}

extern "C" int mfn_ctx_sync(mfn_ctx *c) {
  if (!c) return MFN_ERR_INVALID;
  MFN_CUDA(cudaStreamSynchronize(c->stream));
  return MFN_OK;
}

extern "C" int mfn_ctx_set_sweep(mfn_ctx *c, int mode) {
  if (!c || mode < -1 || mode > 1) return MFN_ERR_INVALID;
  c->sweep_mode = mode;
  return MFN_OK;
}

extern "C" uint64_t mfn_ctx_launch_count(const mfn_ctx *c) { return c ? c->launches.load() : 0; }

extern "C" uint64_t mfn_ctx_h2d_bytes(const mfn_ctx *c) { return c ? c->h2d_bytes.load() : 0; }

extern "C" int mfn_ctx_time_passes(mfn_ctx *c, int enable) {
  if (!c) return MFN_ERR_INVALID;
  c->time_passes = enable != 0;
  c->pass_ev_used = 0;
  return MFN_OK;
}

// next timing event of the ctx (grown on demand), recorded on the launch stream
cudaEvent_t mfn_ctx_mark(mfn_ctx *c) {
  if (c->pass_ev_used == c->pass_ev.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
    c->pass_ev.push_back(e);
  }
  cudaEvent_t e = c->pass_ev[c->pass_ev_used++];
  cudaEventRecord(e, c->stream);
  return e;
}

extern "C" int mfn_ctx_pass_times(mfn_ctx *c, double *scatter_ms, double *probe_ms, uint32_t *nscans) {
  if (!c) return MFN_ERR_INVALID;
  MFN_CUDA(cudaStreamSynchronize(c->stream));
  double a = 0, b = 0;
  uint32_t n = 0;
  for (size_t i = 0; i + 2 < c->pass_ev_used; i += 3) {
    float t1 = 0, t2 = 0;
    if (cudaEventElapsedTime(&t1, c->pass_ev[i], c->pass_ev[i + 1]) != cudaSuccess) break;
    if (cudaEventElapsedTime(&t2, c->pass_ev[i + 1], c->pass_ev[i + 2]) != cudaSuccess) break;
    a += t1; b += t2; n++;
  }
  c->pass_ev_used = 0;
  if (scatter_ms) *scatter_ms = a;
  if (probe_ms) *probe_ms = b;
  if (nscans) *nscans = n;
  return MFN_OK;
}

extern "C" int mfn_host_alloc(void **p, size_t bytes) {
  if (!p) return MFN_ERR_INVALID;
  cudaError_t e = cudaMallocHost(p, bytes ? bytes : 1);
  if (e != cudaSuccess) { mfn_set_error("cudaMallocHost(%zu): %s", bytes, cudaGetErrorString(e)); return MFN_ERR_NOMEM; }
  return MFN_OK;
}

extern "C" void mfn_host_free(void *p) { if (p) cudaFreeHost(p); }

// ---------------------------------------------------------------------------
// host-buffer entry points
// ---------------------------------------------------------------------------

// Scratch sequence set kept in the ctx between host-buffer calls: a caller that streams batches
// of the same shape (the benchmark, a pipeline over chromosomes) does not pay cudaMalloc per call.
static int scratch_seq(mfn_ctx *ctx, const uint64_t *seq_off, uint32_t nseq, mfn_seq **out) {
  mfn_seq *s = ctx->scratch_seq;
  bool reuse = s && s->nseq == nseq;
  for (uint32_t i = 0; reuse && i <= nseq; i++) reuse = s->rec_start[i] == (seq_off[i] - seq_off[0]) + i;
  if (reuse) {
    // the packed arrays belong to the pack stream; scans of the previous call may still read them
    uint64_t gpad = (s->ntiles + 1) * MFN_TILE + 256;
    MFN_CUDA(cudaEventRecord(ctx->ev[MFN_STAGE_BUFS + 2], ctx->stream));
    MFN_CUDA(cudaStreamWaitEvent(ctx->pack_stream, ctx->ev[MFN_STAGE_BUFS + 2], 0));
    MFN_CUDA(cudaMemsetAsync(s->d_seq2, 0, (gpad / 32 + 8) * 8, ctx->pack_stream));
    MFN_CUDA(cudaMemsetAsync(s->d_ok, 0, (gpad / 64 + 8) * 8, ctx->pack_stream));
  } else {
    mfn_seq_free(s);
    ctx->scratch_seq = nullptr;
    MFN_TRY(mfn_seq_alloc(ctx, seq_off, nseq, &s));
    ctx->scratch_seq = s;
  }
  *out = s;
  return MFN_OK;
}

// H2D + pack + scan of one shard, pipelined: the shard's tile range is cut into pieces; while
// the scan kernel of piece i runs, the copy stream is already uploading piece i+1.
extern "C" int mfn_hist_host(mfn_ctx *ctx, const mfn_table *reads, const mfn_table *asm_, const char *bases,
                             const uint64_t *seq_off, uint32_t nseq, const mfn_kmodel *km,
                             uint32_t shard, uint32_t nshards, mfn_accum *acc) {
  if (!ctx || !reads || !asm_ || !seq_off || !acc || nshards == 0 || shard >= nshards) return MFN_ERR_INVALID;
  if (nseq == 0) return MFN_OK;
  mfn_seq *seq = nullptr;
  MFN_TRY(scratch_seq(ctx, seq_off, nseq, &seq));
  const uint64_t tb = seq->ntiles * shard / nshards, te = seq->ntiles * (shard + 1) / nshards;
  const uint64_t pieces = std::max<uint64_t>(1, std::min<uint64_t>(16, (te - tb) / 8192));
  if (reads->dev.k != asm_->dev.k) { mfn_set_error("mer size mismatch: reads k=%u, assembly k=%u", reads->dev.k, asm_->dev.k); return MFN_ERR_K_MISMATCH; }
  if (reads->streamed && !mfn_sweep_applicable(reads, asm_, te - tb, seq->nseq)) {
    mfn_set_error("this read table was streamed (mfn_tables_restrict_stream): it serves -hist of the sequences it was restricted to only");
    return MFN_ERR_UNSUPPORTED;
  }
  if (mfn_sweep_applicable(reads, asm_, te - tb, seq->nseq)) {
    // Range-partitioned scan.  The shard is cut into sub-sweeps of ~1.5e9 windows: each scatters its
    // pieces as soon as they are packed and then probes its bins, so the probe pass of one
    // sub-sweep runs under the ingest of the next and only the last one is exposed (every extra
    // sub-sweep streams the table once more, which is why there are few of them).
    if (acc->nseq < seq->nseq) { mfn_set_error("accumulator has %u record slots, sequence set has %u", acc->nseq, seq->nseq); return MFN_ERR_INVALID; }
    unsigned long long *zmap = nullptr; uint32_t zlen = 0, zhi = 0;
    MFN_TRY(mfn_build_zmap(ctx, km, &zmap, &zlen, &zhi));
    const char *env = getenv("MFN_SUBSWEEPS");
    uint64_t nsw = env ? (uint64_t)atoll(env) : ((te - tb) * MFN_TILE + 750000000ull) / 1500000000ull;
    nsw = std::max<uint64_t>(1, std::min<uint64_t>(nsw, 16));
    // pieces of <= 32768 tiles (64 MB of ASCII); two ingest lanes fill them (ingest.cu): raw upload
    // + k_pack, and host-packed words
    const uint64_t piece_tiles = std::max<uint64_t>(1024, std::min<uint64_t>(32768, ceil_div_u64(te - tb, 8 * nsw)));
    uint32_t np = 0;
    int st = mfn_ingest_begin(ctx, seq, bases + seq_off[0], tb, te, piece_tiles, &np);
    nsw = std::min<uint64_t>(nsw, std::max<uint32_t>(np, 1));
    std::vector<char> waited(np, 0);
    auto wait_piece = [&](uint32_t i) -> int {
      if (i >= np || waited[i]) return MFN_OK;
      waited[i] = 1;
      return mfn_ingest_wait(ctx, i);
    };
    // Only the LAST sub-sweep's probe pass is exposed (nothing is left to ingest under it), so with two
    // sub-sweeps the first one takes the bigger share of the pieces (MFN_SUBSWEEP_FIRST, default 0.5 = even)
    const char *ea = getenv("MFN_ASYNC_PROBE");
    const bool async_probe = ea && atoi(ea) != 0;
    if (async_probe) ctx->side_probe_ctas = std::max(1, std::min(3, atoi(ea)));
    double first_share = 0.5;
    if (const char *ef = getenv("MFN_SUBSWEEP_FIRST")) { first_share = atof(ef); if (!(first_share > 0.05 && first_share < 0.95)) first_share = 0.5; }
    auto bound = [&](uint64_t sw) -> uint32_t {
      if (nsw == 2 && sw == 1) return (uint32_t)std::min<uint64_t>(np, std::max<uint64_t>(1, (uint64_t)(np * first_share + 0.5)));
      return (uint32_t)(np * sw / nsw);
    };
    for (uint64_t sw = 0; sw < nsw && st == MFN_OK; sw++) {
      const uint32_t pa = bound(sw), pb = sw + 1 == nsw ? np : bound(sw + 1);
      if (pa >= pb) continue;
      const uint64_t sa = tb + (uint64_t)pa * piece_tiles, sb = std::min(te, tb + (uint64_t)pb * piece_tiles);
      mfn_sweep *w = nullptr;
      st = mfn_sweep_begin(ctx, reads, asm_, seq, (sb - sa) * MFN_TILE, acc, zmap, zlen, zhi, &w);
      for (uint32_t p = pa; p < pb && st == MFN_OK; p++) {
        // lane G runs a few pieces ahead of the scan; the windows of a piece read k-1 bases of the next one
        st = mfn_ingest_feed(ctx, p + 3);
        if (st == MFN_OK) st = wait_piece(p);
        if (st == MFN_OK) st = wait_piece(p + 1);
        const uint64_t a = tb + (uint64_t)p * piece_tiles, b = std::min(te, a + piece_tiles);
        if (st == MFN_OK) st = mfn_sweep_scatter(ctx, w, a, b);
      }
      // every probe pass but the last runs on the second compute stream, next to the following sub-sweep's scatters
      if (w) { int st2 = mfn_sweep_finish(ctx, w, async_probe && sw + 1 < nsw); if (st == MFN_OK) st = st2; }
    }
    if (async_probe) mfn_sweep_join(ctx);
    mfn_ingest_end(ctx);
    if (zmap) cudaFreeAsync(zmap, ctx->stream);
    return st;
  }
  for (uint64_t p = 0; p < pieces; p++) {
    uint64_t a = tb + (te - tb) * p / pieces, b = tb + (te - tb) * (p + 1) / pieces;
    MFN_TRY(mfn_seq_pack_range(seq, bases + seq_off[0], a * MFN_TILE, b * MFN_TILE + 128));
    MFN_TRY(mfn_hist_tiles(ctx, reads, asm_, seq, km, a, b, acc));
  }
  return MFN_OK;
}

extern "C" int mfn_hist(mfn_ctx *ctx, const mfn_table *reads, const mfn_table *asm_, const char *bases,
                        const uint64_t *seq_off, uint32_t nseq, const mfn_kmodel *km, mfn_seq_counts *per_seq,
                        mfn_pair_count *pairs, uint64_t cap, uint64_t *npairs) {
  if (!ctx || !reads || !asm_ || !seq_off || !npairs) return MFN_ERR_INVALID;
  mfn_accum *acc = ctx->scratch_acc;
  if (acc && acc->nseq < nseq) { mfn_accum_free(acc); acc = ctx->scratch_acc = nullptr; }
  if (!acc) { MFN_TRY(mfn_accum_create(ctx, nseq, &acc)); ctx->scratch_acc = acc; }
  else MFN_TRY(mfn_accum_reset(ctx, acc));
  MFN_TRY(mfn_hist_host(ctx, reads, asm_, bases, seq_off, nseq, km, 0, 1, acc));
  // The range-partitioned scan sizes its bins for k-mers that spread over the table; a batch
  // dominated by a handful of k-mers (one bin taking a large share of all windows) can run out of
  // side-list room, which the accumulator header records.  Redo such a batch with the direct scan.
  unsigned long long lost = 0;
  MFN_CUDA(cudaMemcpyAsync(&lost, acc->d_buf + 2, 8, cudaMemcpyDeviceToHost, ctx->stream));
  MFN_CUDA(cudaStreamSynchronize(ctx->stream));
  if (lost && ctx->sweep_mode != 0) {
    const int keep_mode = ctx->sweep_mode;
    ctx->sweep_mode = 0;
    int st2 = mfn_accum_reset(ctx, acc);
    if (st2 == MFN_OK) st2 = mfn_hist_host(ctx, reads, asm_, bases, seq_off, nseq, km, 0, 1, acc);
    ctx->sweep_mode = keep_mode;
    if (st2 != MFN_OK) return st2;
  }
  uint32_t keep = acc->nseq;
  acc->nseq = nseq;                      // fetch exactly the caller's records
  int st = mfn_accum_fetch(ctx, acc, 1, 1, per_seq, pairs, cap, npairs);
  acc->nseq = keep;
  return st;
}

// mfn_dump / mfn_score_paths: pack the records, one direct scan that writes the pair per position, D2H.
// Several host threads may call this on one ctx: the device part holds ctx->dump_mu (the scratch
// streams and staging buffers of the ctx are shared), the callers' own work overlaps with it.
// Output arrays come from the stream-ordered pool (no cudaMalloc / cudaFree per call).
static int dump_host(mfn_ctx *ctx, const mfn_table *reads, const mfn_table *asm_, const char *bases,
                     const uint64_t *off, uint64_t nrec, uint32_t *v_read, uint32_t *v_asm, uint8_t *valid) {
  if (!ctx || !reads || !asm_ || !off) return MFN_ERR_INVALID;
  if (nrec > 0xfffffff0ull) return MFN_ERR_INVALID;
  if (nrec == 0 || off[nrec] == off[0]) return MFN_OK;
  std::lock_guard<std::mutex> lk(ctx->dump_mu);
  MFN_CUDA(cudaSetDevice(ctx->device));
  mfn_seq *seq = nullptr;
  MFN_TRY(mfn_seq_create_in(ctx, bases, off, (uint32_t)nrec, &seq, true));
  uint64_t n = seq->total_bases;
  uint32_t *dr = nullptr, *da = nullptr; uint8_t *dv = nullptr;
  int st = MFN_OK;
  if (cudaMallocAsync(&dr, n * 4, ctx->stream) != cudaSuccess || cudaMallocAsync(&da, n * 4, ctx->stream) != cudaSuccess ||
      cudaMallocAsync(&dv, n, ctx->stream) != cudaSuccess) {
    mfn_set_error("device memory for the dump output (%llu positions)", (unsigned long long)n);
    cudaGetLastError();
    st = MFN_ERR_NOMEM;
  }
  // (mfn_ctx_time_passes: "scatter" = the scan kernel incl. its output memsets, "probe" = the D2H of the result)
  if (ctx->time_passes) mfn_ctx_mark(ctx);
  if (st == MFN_OK) st = mfn_dump_resident(ctx, reads, asm_, seq, dr, da, dv);
  if (ctx->time_passes) mfn_ctx_mark(ctx);
  if (st == MFN_OK) {
    cudaMemcpyAsync(v_read, dr, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(v_asm, da, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
    cudaMemcpyAsync(valid, dv, n, cudaMemcpyDeviceToHost, ctx->stream);
  }
  if (ctx->time_passes) mfn_ctx_mark(ctx);
  if (dr) cudaFreeAsync(dr, ctx->stream);
  if (da) cudaFreeAsync(da, ctx->stream);
  if (dv) cudaFreeAsync(dv, ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  if (st == MFN_OK && e != cudaSuccess) { mfn_set_error("dump: %s", cudaGetErrorString(e)); st = MFN_ERR_CUDA; }
  mfn_seq_free(seq);
  return st;
}

extern "C" int mfn_dump(mfn_ctx *ctx, const mfn_table *reads, const mfn_table *asm_, const char *bases,
                        const uint64_t *seq_off, uint32_t nseq, uint32_t *v_read, uint32_t *v_asm, uint8_t *valid) {
  return dump_host(ctx, reads, asm_, bases, seq_off, nseq, v_read, v_asm, valid);
}

extern "C" int mfn_score_paths(mfn_ctx *ctx, const mfn_table *reads, const mfn_table *asm_, const char *paths,
                               const uint64_t *path_off, uint64_t npaths, uint32_t *v_read, uint32_t *v_asm,
                               uint8_t *valid) {
  return dump_host(ctx, reads, asm_, paths, path_off, npaths, v_read, v_asm, valid);
}

// ---------------------------------------------------------------------------
// -completeness: stream the read DB, probe the (raw) assembly table
// ---------------------------------------------------------------------------
// The read database holds ~2 error k-mers per genomic one, and they all land in the cells (1, 0) and
// (2, 0): with one global atomic per k-mer the pass ran at the speed of the L2's atomic unit on two
// addresses.  So: equal pairs of a warp are combined first (match_any), the pairs go to a per-CTA
// shared-memory tier [v_asm < 16][v_read < 512] that is flushed once, and each thread has four
// independent lookups in flight.
__global__ void __launch_bounds__(256) k_join(DevTable A, const unsigned long long *kmers, const uint32_t *values,
                                              uint64_t n, unsigned long long *acc, PairTable sparse) {
  __shared__ uint32_t s_hist[SH_R * SH_A];
  for (uint32_t i = threadIdx.x; i < SH_R * SH_A; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  // whole warps stay in the loop together (match_any needs the full mask): the bound is rounded up per warp
  for (uint64_t i0 = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; (i0 - lane) < n; i0 += 4 * stride) {
    uint32_t vr[4], va[4];
    bool live[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const uint64_t i = i0 + u * stride;
      live[u] = i < n;
      vr[u] = live[u] ? values[i] : 0u;
      va[u] = live[u] ? mfn_value(A, kmers[i]) : 0u;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const bool small = live[u] && vr[u] < SH_R && va[u] < SH_A;
      const uint32_t key = small ? va[u] * SH_R + vr[u] : 0xffffffffu - lane;     // others: unique keys, no grouping
      const uint32_t m = __match_any_sync(0xffffffffu, key);
      if (small) {
        if (lane == (uint32_t)__ffs(m) - 1u) atomicAdd(&s_hist[key], (uint32_t)__popc(m));
      } else if (live[u]) {
        if (vr[u] < MFN_DENSE_R && va[u] < MFN_DENSE_A) atomicAdd(&acc[MFN_ACC_HEADER + va[u] * MFN_DENSE_R + vr[u]], 1ull);
        else sparse_add(sparse, acc, vr[u], va[u], 1ull);         // same side table as the scan kernel
      }
    }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < SH_R * SH_A; i += blockDim.x) {
    const uint32_t c = s_hist[i];
    if (c) atomicAdd(&acc[MFN_ACC_HEADER + (i / SH_R) * MFN_DENSE_R + (i % SH_R)], (unsigned long long)c);
  }
}

// One shard of -completeness: data files [64 s / n, 64 (s+1) / n) of the read database (the unit
// the reference's OpenMP loop hands out, merfin-completeness.C:65) are streamed and joined with the
// RAW assembly table into `acc`; ranks sum-all-reduce the accumulator buffers afterwards.
extern "C" int mfn_completeness_shard(mfn_ctx *ctx, const char *reads_dir, const mfn_table *asm_raw, uint32_t shard,
                                      uint32_t nshards, mfn_accum *acc) {
  if (!ctx || !reads_dir || !asm_raw || !acc || nshards == 0 || shard >= nshards) return MFN_ERR_INVALID;
  if (asm_raw->value_offset != 0) { mfn_set_error("-completeness joins raw values: load the assembly table with minV = 0"); return MFN_ERR_INVALID; }
  MerylIndex idx;
  MFN_TRY(meryl_open(reads_dir, idx));
  if (idx.k != asm_raw->dev.k) { mfn_set_error("mer size mismatch: reads k=%u, assembly k=%u", idx.k, asm_raw->dev.k); return MFN_ERR_K_MISMATCH; }
  const uint32_t nfiles = 1u << idx.numFilesBits;
  const uint32_t fb = (uint32_t)((uint64_t)nfiles * shard / nshards), fe = (uint32_t)((uint64_t)nfiles * (shard + 1) / nshards);
  if (fb >= fe) return MFN_OK;
  if (asm_raw->wide) return mfn_wide_completeness(ctx, idx, asm_raw, fb, fe, acc);                                    // k > 32
  return meryl_stream(ctx, idx, [&](const unsigned long long *dk, const uint32_t *dv, uint64_t n) {
    int grid = (int)std::min<uint64_t>(ceil_div_u64(n, 256), (uint64_t)ctx->sm_count * 8);
    k_join<<<grid, 256, 0, ctx->stream>>>(asm_raw->dev, dk, dv, n, acc->d_buf, acc->sparse);
    ctx->launches++;
    return cudaGetLastError() == cudaSuccess ? MFN_OK : MFN_ERR_CUDA;
  }, fb, fe);
}

extern "C" int mfn_completeness(mfn_ctx *ctx, const char *reads_dir, const char *asm_dir,
                                mfn_pair_count *pairs, uint64_t cap, uint64_t *npairs) {
  if (!ctx || !reads_dir || !asm_dir || !npairs) return MFN_ERR_INVALID;
  mfn_table *A = nullptr;
  MFN_TRY(mfn_table_load(ctx, asm_dir, 0, 0xffffffffu, &A));      // raw values: offset 0
  mfn_accum *acc = nullptr;
  int st = mfn_accum_create(ctx, 0, &acc);
  if (st == MFN_OK) st = mfn_completeness_shard(ctx, reads_dir, A, 0, 1, acc);
  if (st == MFN_OK) st = mfn_accum_fetch(ctx, acc, 1, 1, nullptr, pairs, cap, npairs);
  mfn_accum_free(acc);
  mfn_table_free(A);
  return st;
}
