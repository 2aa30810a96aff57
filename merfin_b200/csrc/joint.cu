// Building the joint (read, assembly) table from two per-database tables (see joint.cuh).
#include "joint_build.cuh"

#include <algorithm>
#include <cmath>
#include <functional>
#include <type_traits>

template <bool W64>
__global__ void __launch_bounds__(256) k_joint_fill(JointDev J, DevTable T, int which) {
  JointCounts cnt = {0, 0, 0, 0};
  mfn_foreach_entry(T, [&](uint64_t c, uint32_t v) { joint_insert<W64>(J, c, which, v, cnt); });
  for (int o = 16; o > 0; o >>= 1) {
    cnt.home += __shfl_down_sync(0xffffffffu, cnt.home, o);
    cnt.spill += __shfl_down_sync(0xffffffffu, cnt.spill, o);
    cnt.stash += __shfl_down_sync(0xffffffffu, cnt.stash, o);
    cnt.big += __shfl_down_sync(0xffffffffu, cnt.big, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (cnt.home) atomicAdd(&J.stats->n_home, (unsigned long long)cnt.home);
    if (cnt.spill) atomicAdd(&J.stats->n_spill, (unsigned long long)cnt.spill);
    if (cnt.stash) atomicAdd(&J.stats->n_stash, (unsigned long long)cnt.stash);
    if (cnt.big) atomicAdd(&J.stats->n_big, (unsigned long long)cnt.big);
  }
}

// k-mers of A that R does not hold (sizes the union)
__global__ void __launch_bounds__(256) k_count_absent(DevTable A, DevTable R, unsigned long long *count) {
  unsigned long long n = 0;
  mfn_foreach_entry(A, [&](uint64_t c, uint32_t) { if (mfn_value(R, c) == 0) n++; });
  for (int o = 16; o > 0; o >>= 1) n += __shfl_down_sync(0xffffffffu, n, o);
  if ((threadIdx.x & 31) == 0 && n) atomicAdd(count, n);
}

__global__ void __launch_bounds__(256) k_count_above(DevTable T, uint32_t fmax, unsigned long long *count) {
  unsigned long long n = 0;
  mfn_foreach_entry(T, [&](uint64_t, uint32_t v) { if (v > fmax) n++; });
  for (int o = 16; o > 0; o >>= 1) n += __shfl_down_sync(0xffffffffu, n, o);
  if ((threadIdx.x & 31) == 0 && n) atomicAdd(count, n);
}

__global__ void __launch_bounds__(256) k_joint_lookup(JointDev J, const unsigned long long *kmers, uint64_t n,
                                                      uint32_t *vr, uint32_t *va) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t a = 0, b = 0;
    if (!(kmers[i] & ~J.kmask)) joint_value(J, kmers[i], a, b);
    vr[i] = a; va[i] = b;
  }
}

static uint32_t bits_of(uint64_t x) { uint32_t n = 0; while (x) { n++; x >>= 1; } return n; }
static uint64_t next_pow2(uint64_t x) { uint64_t p = 1; while (p < x) p <<= 1; return p; }

void mfn_joint_free(mfn_joint *j) {
  if (!j) return;
  for (void *p : j->allocs) cudaFree(p);
  delete j;
}

// Geometry and device memory of a joint table for n k-mers (values not yet filled in).  nbig: upper
// bound of the stored values that saturate their field, given the field maxima (sizes the excess table).
int mfn_joint_alloc(mfn_ctx *ctx, const mfn_table *R, const mfn_table *A, uint64_t n_kmers, bool want_pow2,
                       const std::function<uint64_t(uint32_t, uint32_t)> &count_big, mfn_joint **out) {
  const uint32_t k = R->dev.k;
  const uint64_t n = std::max<uint64_t>(n_kmers, 1024);
  mfn_joint *j = new mfn_joint();
  j->reads = R; j->asm_ = A; j->n_kmers = n_kmers;
  JointDev &J = j->dev;
  J.k = k; J.kmask = R->dev.kmask; J.seed = 0;            // the scatter pass hashes with a literal 0 seed
  const char *env = getenv("MFN_JOINT_LOAD");
  double load = env ? atof(env) : 0.70;
  if (!(load > 0.05 && load < 0.99)) load = 0.70;
  for (int w64 = 0; w64 < 2; w64++) {
    const int slots = w64 ? 16 : 32;
    // lines wanted for this load; the divisor actually used is the largest D' = d << t <= D with
    // d in {3, 5, 9, 15} (steps of at most 1.25x), so the table is never fuller than asked
    const uint64_t want_lines = std::max<uint64_t>(16, (uint64_t)std::ceil(n / (slots * load)));
    unsigned __int128 space = (unsigned __int128)1 << (2 * k);
    unsigned __int128 D = space / want_lines;
    if (D >> 63) D = ((unsigned __int128)1 << 63);
    static const uint32_t FS[5][3] = {{3, 0xAAAAAAABu, 1}, {5, 0xCCCCCCCDu, 2}, {9, 0x38E38E39u, 1}, {15, 0x88888889u, 3}, {1, 0, 0}};
    const uint32_t tmin = 2 * k > 32 ? 2 * k - 32 : 0;                    // h >> t must fit 32 bits
    int bf = 0; uint32_t bt = tmin;
    uint64_t best = 0;
    for (int f = 0; f < 5; f++)
      for (uint32_t t = tmin; t < 60; t++) {
        const uint64_t cand = (uint64_t)FS[f][0] << t;
        if ((unsigned __int128)cand > D) break;
        if (cand > best) { best = cand; bf = f; bt = t; }
      }
    // A power-of-two divisor makes line, tag and the sweep's bin plain bit fields of the hash (no
    // division in the scatter pass): the one whose load falls in (hi / 2, hi], hi = MFN_JOINT_LOAD_POW2 (0.62: at most 2.5 % of the k-mers spill out of their home sector)
    if (want_pow2) {
      const char *e2 = getenv("MFN_JOINT_LOAD_POW2");
      double hi_load = e2 ? atof(e2) : 0.62;
      if (!(hi_load > 0.1 && hi_load < 0.99)) hi_load = 0.62;
      for (uint32_t t = std::max(tmin, 1u); t < 60; t++) {
        const double lines = std::ldexp(1.0, 2 * (int)k - (int)t);
        const double ld = (double)n / (lines * slots);
        if (ld <= hi_load && ld > hi_load / 2) { best = 1ull << t; bf = 4; bt = t; break; }
        if (ld > hi_load) break;                                              // fewer lines only get fuller
      }
    }
    if (best == 0) best = (uint64_t)FS[0][0] << tmin;                     // more lines wanted than the geometry allows
    J.fs_d = FS[bf][0]; J.fs_m = FS[bf][1]; J.fs_s = FS[bf][2]; J.fs_t = bt;
    J.D = best;
    J.nlines = (uint64_t)((space + J.D - 1) / J.D);
    const uint32_t T = std::max(1u, bits_of(J.D - 1));
    J.w64 = w64;
    if (!w64) {
      if (T > 20) continue;                       // 32-bit slots need room for both value fields
      J.abits = T > 16 ? 5 : 6; J.rbits = std::min(20u, 32u - T - J.abits);
    } else {
      const uint32_t avail = 64 - std::min(T, 58u);
      J.abits = std::min(20u, avail / 3);
      J.rbits = std::min(32u, avail - J.abits);
    }
    j->tag_bits = T;
    load = std::max(load, (double)n / ((double)J.nlines * slots));       // the side tables are sized for the load actually reached
    break;
  }
  J.rmax = J.rbits >= 32 ? 0xffffffffu : ((1u << J.rbits) - 1u);
  J.amax = (1u << J.abits) - 1u;

  auto dalloc = [&](void **p, size_t bytes, int fill) -> int {
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) { mfn_set_error("cudaMalloc(joint table, %zu bytes): %s", bytes, cudaGetErrorString(e)); return MFN_ERR_NOMEM; }
    j->allocs.push_back(*p);
    j->device_bytes += bytes;
    MFN_CUDA(cudaMemsetAsync(*p, fill, bytes, ctx->stream));
    return MFN_OK;
  };
  int st = dalloc(&J.lines, J.nlines * 128, 0);
  OaTable2 *oas[2] = {&J.stash, &J.big};
  // side-table capacity from the expected overflow of Poisson(slots * load) keys sharing a line
  double lambda = (J.w64 ? 16 : 32) * load, pr = std::exp(-lambda), over = 0.0;
  for (int x = 1; x < 400; x++) { pr *= lambda / x; if (x > (J.w64 ? 16 : 32)) over += (x - (J.w64 ? 16 : 32)) * pr; }
  const uint64_t spill = (uint64_t)(4.0 * over / lambda * n) + 8192;
  const uint64_t nbig = count_big(J.rmax, J.amax);
  uint64_t caps[2] = {next_pow2(2 * spill), next_pow2(2 * nbig + 4096)};
  for (int i = 0; i < 2 && st == MFN_OK; i++) {
    oas[i]->mask = (uint32_t)(caps[i] - 1);
    st = dalloc((void **)&oas[i]->keys, caps[i] * 8, 0xff);
    if (st == MFN_OK) st = dalloc((void **)&oas[i]->v0, caps[i] * 4, 0);
    if (st == MFN_OK) st = dalloc((void **)&oas[i]->v1, caps[i] * 4, 0);
    if (st == MFN_OK) st = dalloc((void **)&oas[i]->special, 8, 0);
  }
  if (st == MFN_OK) st = dalloc((void **)&J.stats, sizeof(JointStats), 0);
  if (st != MFN_OK) { mfn_joint_free(j); return st; }
  *out = j;
  return MFN_OK;
}

// stored values of R above rmax plus stored values of A above amax (streams both tables once)
uint64_t mfn_joint_count_saturated(mfn_ctx *ctx, const mfn_table *R, const mfn_table *A, uint32_t rmax, uint32_t amax) {
  unsigned long long *d_b, nbig = 0;
  if (cudaMalloc(&d_b, 8) != cudaSuccess) return R->n_kmers / 64;
  cudaMemsetAsync(d_b, 0, 8, ctx->stream);
  const int grid = ctx->sm_count * 8;
  k_count_above<<<grid, 256, 0, ctx->stream>>>(R->dev, rmax, d_b);
  k_count_above<<<grid, 256, 0, ctx->stream>>>(A->dev, amax, d_b);
  ctx->launches += 2;
  cudaMemcpyAsync(&nbig, d_b, 8, cudaMemcpyDeviceToHost, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(d_b);
  return nbig;
}

// After the fill kernels: fetch the statistics and check that nothing was lost.
int mfn_joint_finish(mfn_ctx *ctx, mfn_joint *j, bool exact_count) {
  if (cudaMemcpyAsync(&j->stats, j->dev.stats, sizeof(JointStats), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
      cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
    mfn_set_error("joint table build: %s", cudaGetErrorString(cudaGetLastError()));
    return MFN_ERR_CUDA;
  }
  if (j->stats.errors) { mfn_set_error("joint table side tables overflowed (%llu)", j->stats.errors); return MFN_ERR_CAPACITY; }
  const uint64_t held = j->stats.n_home + j->stats.n_spill + j->stats.n_stash;
  if (exact_count ? held != j->n_kmers : held > j->n_kmers) {
    mfn_set_error("joint table holds %llu k-mers, expected %llu", (unsigned long long)held, (unsigned long long)j->n_kmers);
    return MFN_ERR_CUDA;
  }
  j->n_kmers = held;
  return MFN_OK;
}

int mfn_joint_build(mfn_ctx *ctx, const mfn_table *R, const mfn_table *A, mfn_joint **out) {
  if (R->dev.k != A->dev.k) { mfn_set_error("mer size mismatch: reads k=%u, assembly k=%u", R->dev.k, A->dev.k); return MFN_ERR_K_MISMATCH; }
  if (R->dev.probe_both || A->dev.probe_both) { mfn_set_error("joint table needs canonical databases"); return MFN_ERR_UNSUPPORTED; }
  MFN_CUDA(cudaSetDevice(ctx->device));
  const int grid = ctx->sm_count * 8;

  unsigned long long *d_n, absent = 0;
  MFN_CUDA(cudaMalloc(&d_n, 8));
  MFN_CUDA(cudaMemsetAsync(d_n, 0, 8, ctx->stream));
  k_count_absent<<<grid, 256, 0, ctx->stream>>>(A->dev, R->dev, d_n);
  ctx->launches++;
  MFN_CUDA(cudaMemcpyAsync(&absent, d_n, 8, cudaMemcpyDeviceToHost, ctx->stream));
  MFN_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(d_n);

  mfn_joint *j = nullptr;
  MFN_TRY(mfn_joint_alloc(ctx, R, A, R->n_kmers + absent, false,
                      [&](uint32_t rmax, uint32_t amax) { return mfn_joint_count_saturated(ctx, R, A, rmax, amax); }, &j));
  const JointDev &J = j->dev;
  if (J.w64) {
    k_joint_fill<true><<<grid, 256, 0, ctx->stream>>>(J, R->dev, 0);
    k_joint_fill<true><<<grid, 256, 0, ctx->stream>>>(J, A->dev, 1);
  } else {
    k_joint_fill<false><<<grid, 256, 0, ctx->stream>>>(J, R->dev, 0);
    k_joint_fill<false><<<grid, 256, 0, ctx->stream>>>(J, A->dev, 1);
  }
  ctx->launches += 2;
  int st = mfn_joint_finish(ctx, j, true);
  if (st != MFN_OK) { mfn_joint_free(j); return st; }
  *out = j;
  return MFN_OK;
}

// ---- C ABI ------------------------------------------------------------------------------------
extern "C" int mfn_tables_fuse(mfn_ctx *ctx, mfn_table *reads, const mfn_table *asm_) {
  if (!ctx || !reads || !asm_) return MFN_ERR_INVALID;
  MFN_REJECT_WIDE(reads, "mfn_tables_fuse"); MFN_REJECT_WIDE(asm_, "mfn_tables_fuse");
  if (reads->joint && reads->joint->asm_ == asm_) return MFN_OK;
  mfn_joint_free(reads->joint);
  reads->joint = nullptr;
  return mfn_joint_build(ctx, reads, asm_, &reads->joint);
}

extern "C" int mfn_tables_unfuse(mfn_table *reads) {
  if (!reads) return MFN_ERR_INVALID;
  mfn_joint_free(reads->joint);
  reads->joint = nullptr;
  return MFN_OK;
}

extern "C" int mfn_tables_restrict(mfn_ctx *ctx, mfn_table *reads, const mfn_table *asm_, const mfn_seq *seq,
                                   uint32_t shard, uint32_t nshards) {
  if (!ctx || !reads || !asm_ || !seq || nshards == 0 || shard >= nshards) return MFN_ERR_INVALID;
  MFN_REJECT_WIDE(reads, "mfn_tables_restrict"); MFN_REJECT_WIDE(asm_, "mfn_tables_restrict");
  mfn_joint_free(reads->hot);
  reads->hot = nullptr;
  const uint64_t tb = seq->ntiles * shard / nshards, te = seq->ntiles * (shard + 1) / nshards;
  mfn_table *keys = nullptr;
  MFN_TRY(mfn_table_count_seq_range(ctx, reads->dev.k, seq, tb * MFN_TILE, te * MFN_TILE, &keys));
  int st = mfn_joint_build_restricted(ctx, reads, asm_, keys, &reads->hot);
  mfn_table_free(keys);
  return st;
}

extern "C" int mfn_tables_unrestrict(mfn_table *reads) {
  if (!reads) return MFN_ERR_INVALID;
  mfn_joint_free(reads->hot);
  reads->hot = nullptr;
  return MFN_OK;
}

static int joint_info_of(const mfn_joint *j, mfn_joint_info_t *info);

extern "C" int mfn_restricted_info(const mfn_table *reads, mfn_joint_info_t *info) {
  if (!reads || !info) return MFN_ERR_INVALID;
  return joint_info_of(reads->hot, info);
}

extern "C" int mfn_joint_info(const mfn_table *reads, mfn_joint_info_t *info) {
  if (!reads || !info) return MFN_ERR_INVALID;
  return joint_info_of(reads->joint, info);
}

static int joint_info_of(const mfn_joint *j, mfn_joint_info_t *info) {
  memset(info, 0, sizeof *info);
  if (!j) return MFN_OK;
  info->fused = 1;
  info->slot_bits = j->dev.w64 ? 64 : 32;
  info->tag_bits = j->tag_bits; info->read_bits = j->dev.rbits; info->asm_bits = j->dev.abits;
  info->n_kmers = j->n_kmers; info->n_home = j->stats.n_home; info->n_spill = j->stats.n_spill;
  info->n_stash = j->stats.n_stash; info->n_bigvalue = j->stats.n_big;
  info->lines = j->dev.nlines; info->device_bytes = j->device_bytes;
  return MFN_OK;
}

extern "C" int mfn_joint_lookup(mfn_ctx *ctx, const mfn_table *reads, const uint64_t *kmers, uint64_t n,
                                uint32_t *v_read, uint32_t *v_asm) {
  if (!ctx || !reads || !reads->joint || (n && (!kmers || !v_read || !v_asm))) return MFN_ERR_INVALID;
  if (n == 0) return MFN_OK;
  MFN_CUDA(cudaSetDevice(ctx->device));
  unsigned long long *dk; uint32_t *dr, *da;
  MFN_CUDA(cudaMalloc(&dk, n * 8)); MFN_CUDA(cudaMalloc(&dr, n * 4)); MFN_CUDA(cudaMalloc(&da, n * 4));
  cudaMemcpyAsync(dk, kmers, n * 8, cudaMemcpyHostToDevice, ctx->stream);
  int grid = (int)std::min<uint64_t>(ceil_div_u64(n, 256), (uint64_t)ctx->sm_count * 8);
  k_joint_lookup<<<grid, 256, 0, ctx->stream>>>(reads->joint->dev, dk, n, dr, da);
  ctx->launches++;
  cudaMemcpyAsync(v_read, dr, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
  cudaMemcpyAsync(v_asm, da, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  cudaFree(dk); cudaFree(dr); cudaFree(da);
  if (e != cudaSuccess) { mfn_set_error("joint lookup: %s", cudaGetErrorString(e)); return MFN_ERR_CUDA; }
  return MFN_OK;
}
