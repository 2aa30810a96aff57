// The restricted fused table built from a read database that is STREAMED through the device and never
// resident (BASELINE configs[4]: a read database of twice the HBM).  A -hist scan only asks for k-mers of
// the sequences it scans, so the table the probe pass needs holds at most one entry per distinct
// assembly k-mer whatever the size of the read database: the database's (k-mer, value) blocks are
// decoded chunk by chunk (meryl_gpu.cu) and semi-joined against the k-mers of the sequence shard
// (`keys`, counted on the device); everything else -- read-only error k-mers, two thirds of a 30x
// database -- is dropped on the spot.  kmers-exact.C:95-146 sizes one table for the whole database and
// gives up when it does not fit -memory; this is the B200 answer to the same situation.
// (own translation unit: ptxas 12.9 crashes when these kernels share a module with the other
// joint-table kernels)
#include "joint_build.cuh"
#include "synth.cuh"

#include <algorithm>
#include <chrono>

// one decoded chunk of the database: keep what the shard can ask for
__global__ void __launch_bounds__(256) k_joint_fill_stream(JointDev J, DevTable keys, const unsigned long long *dk, const uint32_t *dv,
                                                          uint64_t n, uint32_t minV, uint32_t maxV, uint32_t voff, int keep_existing,
                                                          unsigned long long *kept) {
  JointCounts cnt = {0, 0, 0, 0};
  unsigned long long mine = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t v = dv[i];
    if (v < minV || v > maxV) continue;
    const uint64_t c = dk[i];
    if (mfn_revcomp(c, J.k) < c) { atomicAdd(kept + 1, 1ull); continue; }     // not a canonical database (Q2): the caller loads it instead
    if (mfn_value(keys, c) == 0) continue;                        // not a k-mer of the shard
    if (keep_existing) { uint32_t vr, va; joint_value(J, c, vr, va); if (vr) continue; }
    if (J.w64) joint_insert<true>(J, c, 0, v - voff, cnt); else joint_insert<false>(J, c, 0, v - voff, cnt);
    mine++;
  }
  for (int o = 16; o > 0; o >>= 1) {
    cnt.home += __shfl_down_sync(0xffffffffu, cnt.home, o);
    cnt.spill += __shfl_down_sync(0xffffffffu, cnt.spill, o);
    cnt.stash += __shfl_down_sync(0xffffffffu, cnt.stash, o);
    cnt.big += __shfl_down_sync(0xffffffffu, cnt.big, o);
    mine += __shfl_down_sync(0xffffffffu, mine, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (cnt.home) atomicAdd(&J.stats->n_home, (unsigned long long)cnt.home);
    if (cnt.spill) atomicAdd(&J.stats->n_spill, (unsigned long long)cnt.spill);
    if (cnt.stash) atomicAdd(&J.stats->n_stash, (unsigned long long)cnt.stash);
    if (cnt.big) atomicAdd(&J.stats->n_big, (unsigned long long)cnt.big);
    if (mine) atomicAdd(kept, mine);
  }
}

// joint_stream_asm.cu (one entry per translation unit that uses the out-of-line lookups: ptxas 12.9)
void mfn_joint_fill_asm_launch(const JointDev &J, const DevTable &keys, const DevTable &A, int synth, double cov, uint64_t seed,
                               int grid, cudaStream_t stream);

// synthetic database blocks: error k-mers [i0, i0 + n) of k_synth_errors (table.cu), as (k-mer, value) arrays
__global__ void __launch_bounds__(256) k_synth_error_chunk(uint64_t kmask, uint32_t k, uint64_t i0, uint64_t n, uint64_t seed,
                                                          unsigned long long *dk, uint32_t *dv) {
  for (uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; j < n; j += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i = i0 + j;
    const uint64_t x = mfn_rng(i ^ (seed * 0x100000001b3ull)) & kmask;
    const uint64_t rc = mfn_revcomp(x, k);
    dk[j] = x < rc ? x : rc;
    dv[j] = 1u + (uint32_t)((mfn_rng(x + seed) >> 40) & 1u) * ((mfn_rng(x) >> 41) & 1u);
  }
}

// an empty read table that stands for a database that was streamed, with the restricted table attached
static int make_streamed_handle(mfn_ctx *ctx, uint32_t k, mfn_table **out) {
  mfn_table *t = nullptr;
  MFN_TRY(mfn_table_alloc(ctx, k, 1024, &t));
  t->min_value = 1; t->max_value = 0xffffffffu; t->value_offset = 0;
  int st = mfn_table_finalize(t, false);
  if (st != MFN_OK) { mfn_table_free(t); return st; }
  t->streamed = true;
  *out = t;
  return MFN_OK;
}

static int restrict_begin(mfn_ctx *ctx, uint32_t k, const mfn_table *asm_, const mfn_seq *seq, uint32_t shard, uint32_t nshards,
                          mfn_table **reads, mfn_table **keys, mfn_joint **j) {
  if (asm_->dev.k != k) { mfn_set_error("mer size mismatch: reads k=%u, assembly k=%u", k, asm_->dev.k); return MFN_ERR_K_MISMATCH; }
  if (asm_->dev.probe_both) { mfn_set_error("the restricted table needs canonical databases"); return MFN_ERR_UNSUPPORTED; }
  MFN_CUDA(cudaSetDevice(ctx->device));
  const uint64_t tb = seq->ntiles * shard / nshards, te = seq->ntiles * (shard + 1) / nshards;
  MFN_TRY(make_streamed_handle(ctx, k, reads));
  int st = mfn_table_count_seq_range(ctx, k, seq, tb * MFN_TILE, te * MFN_TILE, keys);
  if (st == MFN_OK) {
    const uint64_t nk = (*keys)->n_kmers;
    st = mfn_joint_alloc(ctx, *reads, asm_, nk, true, [&](uint32_t, uint32_t) { return nk / 64 + 4096; }, j);
  }
  if (st != MFN_OK) { mfn_table_free(*keys); mfn_table_free(*reads); *keys = nullptr; *reads = nullptr; }
  return st;
}

static int restrict_end(mfn_ctx *ctx, mfn_table *reads, mfn_table *keys, mfn_joint *j, int st) {
  if (st == MFN_OK) st = mfn_joint_finish(ctx, j, false);
  mfn_table_free(keys);
  if (st != MFN_OK) { mfn_joint_free(j); mfn_table_free(reads); return st; }
  j->restricted = true;
  reads->hot = j;
  return MFN_OK;
}

extern "C" int mfn_tables_restrict_stream(mfn_ctx *ctx, const char *reads_dir, uint32_t minV, uint32_t maxV, const mfn_table *asm_,
                                          const mfn_seq *seq, uint32_t shard, uint32_t nshards, mfn_table **reads_out,
                                          mfn_stream_stats *stats) {
  if (!ctx || !reads_dir || !asm_ || !seq || !reads_out || nshards == 0 || shard >= nshards) return MFN_ERR_INVALID;
  MerylIndex idx;
  MFN_TRY(meryl_open(reads_dir, idx));
  if (idx.k > 32) { mfn_set_error("'%s': k=%u; the streamed (restricted) table holds 64-bit k-mers: load the database with mfn_table_load", reads_dir, idx.k); return MFN_ERR_UNSUPPORTED; }
  MFN_REJECT_WIDE(asm_, "mfn_tables_restrict_stream");
  if (idx.flags & 1) { mfn_set_error("'%s' is a multiset database; not supported", reads_dir); return MFN_ERR_UNSUPPORTED; }
  if (minV == 0) minV = 1;                                                 // kmers-exact.C:43
  if (maxV == 0xffffffffu && !idx.histValue.empty()) maxV = (uint32_t)idx.histValue.back();
  const auto t0 = std::chrono::steady_clock::now();
  mfn_table *reads = nullptr, *keys = nullptr;
  mfn_joint *j = nullptr;
  MFN_TRY(restrict_begin(ctx, idx.k, asm_, seq, shard, nshards, &reads, &keys, &j));
  reads->min_value = minV; reads->max_value = maxV; reads->value_offset = minV - 1;
  unsigned long long *d_kept = nullptr, kept[2] = {0, 0};      // [0] k-mers kept, [1] non-canonical k-mers met
  uint64_t streamed = 0;
  int st = cudaMalloc(&d_kept, 16) == cudaSuccess ? MFN_OK : MFN_ERR_NOMEM;
  if (st == MFN_OK) cudaMemsetAsync(d_kept, 0, 16, ctx->stream);
  const int grid = ctx->sm_count * 8;
  if (st == MFN_OK)
    st = meryl_stream(ctx, idx, [&](const unsigned long long *dk, const uint32_t *dv, uint64_t n) {
      k_joint_fill_stream<<<grid, 256, 0, ctx->stream>>>(j->dev, keys->dev, dk, dv, n, minV, maxV, minV - 1, 0, d_kept);
      ctx->launches++;
      streamed += n;
      return cudaGetLastError() == cudaSuccess ? MFN_OK : MFN_ERR_CUDA;
    });
  if (st == MFN_OK) {
    mfn_joint_fill_asm_launch(j->dev, keys->dev, asm_->dev, 0, 0.0, 0, grid, ctx->stream);
    ctx->launches++;
    cudaMemcpyAsync(kept, d_kept, 16, cudaMemcpyDeviceToHost, ctx->stream);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = MFN_ERR_CUDA;
    if (st == MFN_OK && kept[1]) { mfn_set_error("'%s' is not a canonical database (%llu k-mers): load it with mfn_table_load", reads_dir, kept[1]); st = MFN_ERR_UNSUPPORTED; }
  }
  const uint64_t nkeys = keys ? keys->n_kmers : 0;
  st = restrict_end(ctx, reads, keys, j, st);
  cudaFree(d_kept);
  if (st != MFN_OK) return st;
  if (stats) {
    stats->db_kmers = streamed; stats->kept_kmers = kept[0]; stats->shard_kmers = nkeys;
    stats->table_bytes = reads->hot->device_bytes;
    stats->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  *reads_out = reads;
  return MFN_OK;
}

extern "C" int mfn_tables_restrict_synth(mfn_ctx *ctx, const mfn_table *asm_, const mfn_seq *seq, uint32_t shard, uint32_t nshards,
                                         double cov, uint64_t n_error, uint64_t seed, uint64_t chunk_kmers, mfn_table **reads_out,
                                         mfn_stream_stats *stats) {
  if (!ctx || !asm_ || !seq || !reads_out || nshards == 0 || shard >= nshards) return MFN_ERR_INVALID;
  MFN_REJECT_WIDE(asm_, "mfn_tables_restrict_synth");
  const auto t0 = std::chrono::steady_clock::now();
  mfn_table *reads = nullptr, *keys = nullptr;
  mfn_joint *j = nullptr;
  MFN_TRY(restrict_begin(ctx, asm_->dev.k, asm_, seq, shard, nshards, &reads, &keys, &j));
  const int grid = ctx->sm_count * 8;
  // the genomic part of the synthetic database: a value for every k-mer the assembly holds (only those of
  // the shard can ever be asked for), together with the assembly multiplicities
  mfn_joint_fill_asm_launch(j->dev, keys->dev, asm_->dev, 1, cov, seed, grid, ctx->stream);
  ctx->launches++;
  // the error k-mers, in database-block-sized chunks through the same semi-join a real database takes
  if (chunk_kmers == 0) chunk_kmers = 1ull << 28;
  chunk_kmers = std::min<uint64_t>(chunk_kmers, std::max<uint64_t>(n_error, 1));
  unsigned long long *dk[2] = {nullptr, nullptr}, *d_kept = nullptr, kept = 0;
  uint32_t *dv[2] = {nullptr, nullptr};
  int st = MFN_OK;
  for (int b = 0; b < 2 && st == MFN_OK; b++)
    if (cudaMalloc(&dk[b], chunk_kmers * 8) != cudaSuccess || cudaMalloc(&dv[b], chunk_kmers * 4) != cudaSuccess) { mfn_set_error("synthetic stream buffers"); st = MFN_ERR_NOMEM; }
  if (st == MFN_OK && cudaMalloc(&d_kept, 16) != cudaSuccess) st = MFN_ERR_NOMEM;
  if (st == MFN_OK) cudaMemsetAsync(d_kept, 0, 16, ctx->stream);
  int b = 0;
  for (uint64_t i0 = 0; i0 < n_error && st == MFN_OK; i0 += chunk_kmers, b ^= 1) {
    const uint64_t n = std::min<uint64_t>(chunk_kmers, n_error - i0);
    k_synth_error_chunk<<<grid, 256, 0, ctx->stream>>>(asm_->dev.kmask, asm_->dev.k, i0, n, seed, dk[b], dv[b]);
    k_joint_fill_stream<<<grid, 256, 0, ctx->stream>>>(j->dev, keys->dev, dk[b], dv[b], n, 1, 0xffffffffu, 0, 1, d_kept);
    ctx->launches += 2;
    if (cudaGetLastError() != cudaSuccess) st = MFN_ERR_CUDA;
  }
  if (st == MFN_OK) cudaMemcpyAsync(&kept, d_kept, 8, cudaMemcpyDeviceToHost, ctx->stream);
  const uint64_t nkeys = keys->n_kmers;
  st = restrict_end(ctx, reads, keys, j, st);
  for (int x = 0; x < 2; x++) { cudaFree(dk[x]); cudaFree(dv[x]); }
  cudaFree(d_kept);
  if (st != MFN_OK) return st;
  if (stats) {
    stats->db_kmers = n_error + asm_->n_kmers; stats->kept_kmers = kept; stats->shard_kmers = nkeys;
    stats->table_bytes = reads->hot->device_bytes;
    stats->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  *reads_out = reads;
  return MFN_OK;
}
