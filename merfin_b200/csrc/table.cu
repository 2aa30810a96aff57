// Table construction, batch lookup and export kernels.
#include "table.cuh"

#include <cub/cub.cuh>
#include "table_host.h"

#include <algorithm>
#include <cmath>
#include <type_traits>

// ---------------------------------------------------------------------------
// side tables
// ---------------------------------------------------------------------------
__device__ __forceinline__ void mfn_oa_add(const OaTable &t, uint64_t key, uint32_t v, TableStats *st) {
  if (key == MFN_EMPTY64) { atomicAdd(t.special, v); return; }
  uint32_t i = mfn_oa_hash(key) & t.mask;
  for (uint32_t n = 0; n <= t.mask; n++) {
    unsigned long long cur = t.keys[i];
    if (cur == MFN_EMPTY64) {
      cur = atomicCAS(&t.keys[i], MFN_EMPTY64, (unsigned long long)key);
      if (cur == MFN_EMPTY64) cur = key;
    }
    if (cur == key) { atomicAdd(&t.vals[i], v); return; }
    i = (i + 1) & t.mask;
  }
  atomicAdd(&st->errors, 1ull);                 // side table full
}

// ---------------------------------------------------------------------------
// find-or-insert.  ADD: value accumulates (k-mer counting); otherwise first writer wins.
// ---------------------------------------------------------------------------
struct InsertCounts { uint32_t l1, l2, stash, big, dup; };

template <bool W64, bool ADD>
__device__ __forceinline__ void mfn_insert(const DevTable &t, uint64_t c, uint32_t v, InsertCounts &cnt) {
  typedef typename std::conditional<W64, unsigned long long, unsigned int>::type slot_t;
  constexpr int S = W64 ? 4 : 8;
  const slot_t vmask = (slot_t)t.vmax;
#pragma unroll 1
  for (int lvl = 0; lvl < 2; lvl++) {
    uint64_t bucket, tag;
    mfn_split(mfn_mix(c, t.lv[lvl].seed, t.kmask, t.k), t.lv[lvl].D, t.lv[lvl].magic, bucket, tag);
    slot_t *slots = (slot_t *)t.lv[lvl].slots + bucket * S;
    const slot_t want = (slot_t)tag << t.vbits;
#pragma unroll 1
    for (int i = 0; i < S; i++) {
      slot_t cur = *(volatile slot_t *)(slots + i);
      if (cur == 0) {
        slot_t nv = want | (slot_t)(v < t.vmax ? v : t.vmax);
        slot_t old = atomicCAS(slots + i, (slot_t)0, nv);
        if (old == 0) {
          if (lvl == 0) cnt.l1++; else cnt.l2++;
          if (v > t.vmax) { mfn_oa_add(t.big, c, v - t.vmax, t.stats); cnt.big++; }
          return;
        }
        cur = old;
      }
      if ((cur & ~vmask) == want) {               // the k-mer is already here
        if (!ADD) { cnt.dup++; return; }
        while (true) {
          uint32_t val = (uint32_t)(cur & vmask);
          if (val == t.vmax) { mfn_oa_add(t.big, c, v, t.stats); return; }
          uint64_t sum = (uint64_t)val + v;
          uint32_t nvv = sum < t.vmax ? (uint32_t)sum : t.vmax;
          slot_t old = atomicCAS(slots + i, cur, (slot_t)((cur & ~vmask) | nvv));
          if (old == cur) {
            if (sum > nvv) { mfn_oa_add(t.big, c, (uint32_t)(sum - nvv), t.stats); cnt.big++; }
            return;
          }
          cur = old;
        }
      }
    }
  }
  // both buckets full: level 3
  if (!ADD && mfn_oa_get(t.stash, c) != 0) { cnt.dup++; return; }
  mfn_oa_add(t.stash, c, v, t.stats);
  cnt.stash++;                                    // (ADD: counts increments, fixed up at finalize)
}

__device__ __forceinline__ void mfn_flush_counts(const DevTable &t, InsertCounts &cnt) {
  for (int o = 16; o > 0; o >>= 1) {
    cnt.l1 += __shfl_down_sync(0xffffffffu, cnt.l1, o);
    cnt.l2 += __shfl_down_sync(0xffffffffu, cnt.l2, o);
    cnt.stash += __shfl_down_sync(0xffffffffu, cnt.stash, o);
    cnt.big += __shfl_down_sync(0xffffffffu, cnt.big, o);
    cnt.dup += __shfl_down_sync(0xffffffffu, cnt.dup, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (cnt.l1) atomicAdd(&t.stats->n_l1, (unsigned long long)cnt.l1);
    if (cnt.l2) atomicAdd(&t.stats->n_l2, (unsigned long long)cnt.l2);
    if (cnt.stash) atomicAdd(&t.stats->n_stash, (unsigned long long)cnt.stash);
    if (cnt.big) atomicAdd(&t.stats->n_big, (unsigned long long)cnt.big);
    if (cnt.dup) atomicAdd(&t.stats->n_dup, (unsigned long long)cnt.dup);
  }
}

// (k-mer, value) arrays -> table.  Values are filtered/offset as merylExactLookup does
// (kmers-exact.C:578-580, 613): keep minV <= v <= maxV, store v - (minV - 1).
template <bool W64, bool ADD>
__global__ void __launch_bounds__(256) k_insert_pairs(DevTable t, const unsigned long long *kmers,
                                                      const uint32_t *values, uint64_t n,
                                                      uint32_t minV, uint32_t maxV, uint32_t voff) {
  InsertCounts cnt = {0, 0, 0, 0, 0};
  bool noncanon = false;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t v = values ? values[i] : 1u;
    if (v < minV || v > maxV) continue;
    uint64_t c = kmers[i] & t.kmask;
    if (c > mfn_revcomp(c, t.k)) noncanon = true;
    mfn_insert<W64, ADD>(t, c, v - voff, cnt);
  }
  if (noncanon) t.stats->n_noncanon = 1;
  mfn_flush_counts(t, cnt);
}

__global__ void __launch_bounds__(256) k_lookup(DevTable t, const unsigned long long *kmers, uint64_t n, uint32_t *out) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t c = kmers[i];
    out[i] = (c & ~t.kmask) ? 0u : mfn_value(t, c);
  }
}

__global__ void __launch_bounds__(256) k_export(DevTable t, unsigned long long *kmers, uint32_t *values,
                                                unsigned long long cap, unsigned long long *count) {
  mfn_foreach_entry(t, [&](uint64_t c, uint32_t v) {
    unsigned long long i = atomicAdd(count, 1ull);
    if (i < cap) { kmers[i] = c; values[i] = v; }
  });
}

__global__ void __launch_bounds__(256) k_count_entries(DevTable t, unsigned long long *count) {
  unsigned long long n = 0;
  mfn_foreach_entry(t, [&](uint64_t, uint32_t) { n++; });
  for (int o = 16; o > 0; o >>= 1) n += __shfl_down_sync(0xffffffffu, n, o);
  if ((threadIdx.x & 31) == 0 && n) atomicAdd(count, n);
}

// ---------------------------------------------------------------------------
// k-mer counting of a packed sequence set (replaces "meryl count" for -seqmers)
// and the synthetic read table.
// ---------------------------------------------------------------------------

// forward k-mer (first base in the high bits) of the window starting at global base g
__device__ __forceinline__ bool mfn_window(const unsigned long long *seq2, const unsigned long long *ok,
                                           uint64_t g, uint32_t k, uint64_t kmask, uint64_t &f, uint64_t &r) {
  uint64_t w = g >> 6, o = g & 63;
  uint64_t m = ok[w] >> o;
  if (o) m |= ok[w + 1] << (64 - o);
  uint64_t kb = (1ull << k) - 1;                   // k <= 32
  if ((m & kb) != kb) return false;
  uint64_t ws = g >> 5, os = (g & 31) * 2;
  uint64_t x = seq2[ws] >> os;
  if (os) x |= seq2[ws + 1] << (64 - os);
  x &= kmask;                                      // base g in the low bits: reverse of f
  r = x ^ (0xaaaaaaaaaaaaaaaaull & kmask);         // complement only: this is the reverse complement
  f = mfn_revcomp(r, k);
  return true;
}

template <bool W64>
__global__ void __launch_bounds__(256) k_count_seq(DevTable t, const unsigned long long *seq2,
                                                   const unsigned long long *ok, uint64_t g_begin, uint64_t G) {
  InsertCounts cnt = {0, 0, 0, 0, 0};
  // Adjacent windows of homopolymers and short tandem repeats are the same k-mer: the lanes of a
  // warp that hold equal k-mers elect one to insert their sum, so a megabase satellite costs
  // 1/32 of the increments (and of the contention on its slot).  The loop bound is warp-uniform.
  const uint32_t lane = threadIdx.x & 31;
  for (uint64_t g0 = g_begin + blockIdx.x * (uint64_t)blockDim.x + (threadIdx.x & ~31u); g0 < G; g0 += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t g = g0 + lane;
    uint64_t f = 0, r = 0;
    const bool valid = g < G && mfn_window(seq2, ok, g, t.k, t.kmask, f, r);
    const uint64_t c = f < r ? f : r;
    const uint32_t have = __ballot_sync(0xffffffffu, valid);
    if (!valid) continue;
    const uint32_t peers = __match_any_sync(have, c);
    if ((uint32_t)(__ffs(peers) - 1) == lane) mfn_insert<W64, true>(t, c, (uint32_t)__popc(peers), cnt);
  }
  mfn_flush_counts(t, cnt);
}

#include "synth.cuh"

template <bool W64>
__global__ void __launch_bounds__(256) k_synth_from_asm(DevTable reads, DevTable asmt, double cov, uint64_t seed) {
  InsertCounts cnt = {0, 0, 0, 0, 0};
  mfn_foreach_entry(asmt, [&](uint64_t c, uint32_t a) {
    // 0.5 % of k-mers are "collapsed" in the assembly (reads see twice the copies) and 0.5 %
    // "falsely duplicated" (half), so K* bins away from 0 are populated
    uint64_t sel = mfn_rng(c + 0x5851f42d4c957f2dull * (seed + 1)) % 1000;
    double lambda = cov * (double)a * (sel < 5 ? 2.0 : (sel < 10 ? 0.5 : 1.0));
    uint32_t v = mfn_poisson(lambda, c, seed);
    if (v) mfn_insert<W64, false>(reads, c, v, cnt);
  });
  mfn_flush_counts(reads, cnt);
}

template <bool W64>
__global__ void __launch_bounds__(256) k_copy_table(DevTable dst, DevTable src) {
  InsertCounts cnt = {0, 0, 0, 0, 0};
  mfn_foreach_entry(src, [&](uint64_t c, uint32_t v) { mfn_insert<W64, false>(dst, c, v, cnt); });
  mfn_flush_counts(dst, cnt);
}

template <bool W64>
__global__ void __launch_bounds__(256) k_synth_errors(DevTable reads, uint64_t n_error, uint64_t seed) {
  InsertCounts cnt = {0, 0, 0, 0, 0};
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_error; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t x = mfn_rng(i ^ (seed * 0x100000001b3ull)) & reads.kmask;
    uint64_t rc = mfn_revcomp(x, reads.k);
    uint64_t c = x < rc ? x : rc;
    uint32_t v = 1u + (uint32_t)((mfn_rng(x + seed) >> 40) & 1u) * ((mfn_rng(x) >> 41) & 1u);   // 1 (75 %) or 2
    mfn_insert<W64, false>(reads, c, v, cnt);      // an existing genomic k-mer keeps its value
  }
  mfn_flush_counts(reads, cnt);
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------

static uint32_t bits_of(uint64_t x) { uint32_t n = 0; while (x) { n++; x >>= 1; } return n; }
static uint64_t next_pow2(uint64_t x) { uint64_t p = 1; while (p < x) p <<= 1; return p; }

// expected fraction of entries that do not fit when Poisson(lambda) keys share S slots
static double overflow_fraction(double lambda, int S) {
  double p = std::exp(-lambda), e = 0.0;
  for (int x = 1; x < 200; x++) {
    p *= lambda / x;
    if (x > S) e += (x - S) * p;
  }
  return e / lambda;
}

static void level_geometry(uint32_t k, uint64_t buckets, TableLevel &lv, uint32_t &tagbits) {
  unsigned __int128 space = (unsigned __int128)1 << (2 * k);
  unsigned __int128 D = (space + buckets - 1) / buckets;
  if (D < 1) D = 1;
  if (D >> 64) D = ((unsigned __int128)1 << 64) - 1;
  lv.buckets = buckets;
  lv.D = (uint64_t)D;
  lv.magic = (lv.D == 1) ? ~0ull : (uint64_t)((((unsigned __int128)1) << 64) / lv.D);
  tagbits = std::max(1u, bits_of(lv.D - 1));
}

int mfn_table_alloc(mfn_ctx *ctx, uint32_t k, uint64_t n_expected, mfn_table **out,
                    const std::function<uint64_t(uint32_t)> &count_big) {
  if (k < 1 || k > 32) { mfn_set_error("k=%u: the bucketed tables hold 64-bit k-mers (1 <= k <= 32); k-mers of 33..64 bases come from mfn_table_load / mfn_table_count_seq", k); return MFN_ERR_UNSUPPORTED; }
  MFN_CUDA(cudaSetDevice(ctx->device));
  mfn_table *t = new mfn_table();
  t->ctx = ctx;
  t->device = ctx->device;
  t->device = ctx->device;
  DevTable &d = t->dev;
  d.k = k;
  d.kmask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
  const char *env = getenv("MFN_TABLE_LOAD");
  double load_scale = env ? atof(env) : 1.0;
  if (!(load_scale > 0.1 && load_scale < 2.0)) load_scale = 1.0;

  uint64_t n = std::max<uint64_t>(n_expected, 1024);
  uint32_t t1 = 0, t2 = 0;
  for (int w64 = 0; w64 < 2; w64++) {
    int S = w64 ? 4 : 8;
    double l1 = (w64 ? 2.0 : 5.0) * load_scale, l2 = (w64 ? 1.2 : 2.5);
    uint64_t b1 = std::max<uint64_t>(64, (uint64_t)std::ceil(n / l1));
    uint64_t n2 = (uint64_t)(overflow_fraction(l1, S) * 1.3 * n) + 2048;
    uint64_t b2 = std::max<uint64_t>(64, (uint64_t)std::ceil(n2 / l2));
    level_geometry(k, b1, d.lv[0], t1);
    level_geometry(k, b2, d.lv[1], t2);
    d.w64 = w64;
    uint32_t tmax = std::max(t1, t2);
    if (!w64 && tmax > 22) continue;                  // fewer than 10 value bits: use 64-bit slots
    d.vbits = w64 ? std::min(32u, 64u - tmax) : 32u - tmax;
    break;
  }
  d.vmax = d.vbits >= 32 ? 0xffffffffu : ((1u << d.vbits) - 1u);
  t->tag_bits = std::max(t1, t2);
  d.lv[0].seed = 0x9e3779b97f4a7c15ull & d.kmask;
  d.lv[1].seed = 0x3c6ef372fe94f82bull & d.kmask;

  auto dalloc = [&](void **p, size_t bytes, int fill) -> int {
    MFN_CUDA(cudaMalloc(p, bytes));
    t->allocs.push_back(*p);
    t->device_bytes += bytes;
    MFN_CUDA(cudaMemsetAsync(*p, fill, bytes, ctx->stream));
    return MFN_OK;
  };
  int st = MFN_OK;
  for (int l = 0; l < 2 && st == MFN_OK; l++) st = dalloc(&d.lv[l].slots, d.lv[l].buckets * 32, 0);
  OaTable *oas[2] = {&d.stash, &d.big};
  // excess table for saturated values: sized exactly when the caller can count them
  const uint64_t nbig = count_big ? count_big(d.vmax) : n / 64;
  uint64_t caps[2] = {next_pow2(std::max<uint64_t>(1 << 12, n / 1024)), next_pow2(2 * nbig + 4096)};
  for (int i = 0; i < 2 && st == MFN_OK; i++) {
    oas[i]->mask = (uint32_t)(caps[i] - 1);
    st = dalloc((void **)&oas[i]->keys, caps[i] * 8, 0xff);
    if (st == MFN_OK) st = dalloc((void **)&oas[i]->vals, caps[i] * 4, 0);
    if (st == MFN_OK) st = dalloc((void **)&oas[i]->special, 4, 0);
  }
  if (st == MFN_OK) st = dalloc((void **)&d.stats, sizeof(TableStats), 0);
  if (st != MFN_OK) { mfn_table_free(t); return st; }
  *out = t;
  return MFN_OK;
}

static int grid_for(mfn_ctx *ctx, uint64_t n, int per_sm = 8) {
  uint64_t blocks = ceil_div_u64(std::max<uint64_t>(n, 1), 256);
  return (int)std::min<uint64_t>(blocks, (uint64_t)ctx->sm_count * per_sm);
}

int mfn_table_insert_device(mfn_table *t, const unsigned long long *d_kmers, const uint32_t *d_values,
                            uint64_t n, uint32_t minV, uint32_t maxV, uint32_t voff, bool add) {
  if (n == 0) return MFN_OK;
  mfn_ctx *ctx = t->ctx;
  int grid = grid_for(ctx, n);
  if (t->dev.w64) {
    if (add) k_insert_pairs<true, true><<<grid, 256, 0, ctx->stream>>>(t->dev, d_kmers, d_values, n, minV, maxV, voff);
    else     k_insert_pairs<true, false><<<grid, 256, 0, ctx->stream>>>(t->dev, d_kmers, d_values, n, minV, maxV, voff);
  } else {
    if (add) k_insert_pairs<false, true><<<grid, 256, 0, ctx->stream>>>(t->dev, d_kmers, d_values, n, minV, maxV, voff);
    else     k_insert_pairs<false, false><<<grid, 256, 0, ctx->stream>>>(t->dev, d_kmers, d_values, n, minV, maxV, voff);
  }
  ctx->launches++;
  MFN_CUDA(cudaGetLastError());
  return MFN_OK;
}

int mfn_table_finalize(mfn_table *t, bool counted) {
  mfn_ctx *ctx = t->ctx;
  if (counted) {            // counting mode bumps n_stash per increment: recount the entries
    unsigned long long *d_n;
    MFN_CUDA(cudaMalloc(&d_n, 8));
    MFN_CUDA(cudaMemsetAsync(d_n, 0, 8, ctx->stream));
    k_count_entries<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(t->dev, d_n);
    ctx->launches++;
    unsigned long long n = 0;
    MFN_CUDA(cudaMemcpyAsync(&n, d_n, 8, cudaMemcpyDeviceToHost, ctx->stream));
    MFN_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_n);
    MFN_CUDA(cudaMemcpy(&t->host_stats, t->dev.stats, sizeof(TableStats), cudaMemcpyDeviceToHost));
    t->host_stats.n_stash = n - t->host_stats.n_l1 - t->host_stats.n_l2;
    t->n_kmers = n;
  } else {
    MFN_CUDA(cudaMemcpyAsync(&t->host_stats, t->dev.stats, sizeof(TableStats), cudaMemcpyDeviceToHost, ctx->stream));
    MFN_CUDA(cudaStreamSynchronize(ctx->stream));
    t->n_kmers = t->host_stats.n_l1 + t->host_stats.n_l2 + t->host_stats.n_stash;
  }
  if (t->host_stats.errors) {
    mfn_set_error("table side tables overflowed (%llu inserts lost); raise capacity", t->host_stats.errors);
    return MFN_ERR_CAPACITY;
  }
  t->dev.probe_both = t->host_stats.n_noncanon ? 1 : 0;
  return MFN_OK;
}

extern "C" void mfn_table_free(mfn_table *t) {
  if (!t) return;
  cudaSetDevice(t->device);
  mfn_joint_free(t->joint);
  mfn_joint_free(t->hot);
  for (void *p : t->allocs) cudaFree(p);
  mfn_wide_free(t->wide);
  delete t;
}

extern "C" int mfn_table_info(const mfn_table *t, mfn_table_info_t *info) {
  if (!t || !info) return MFN_ERR_INVALID;
  memset(info, 0, sizeof *info);
  if (t->wide) {
    mfn_wide_info(t, info);
    info->min_value = t->min_value; info->max_value = t->max_value; info->value_offset = t->value_offset;
    info->n_kmers = t->n_kmers;
    info->device_bytes = t->device_bytes;
    return MFN_OK;
  }
  info->k = t->dev.k;
  info->canonical = t->dev.probe_both ? 0 : 1;
  info->slot_bits = t->dev.w64 ? 64 : 32;
  info->tag_bits = t->tag_bits;
  info->value_bits = t->dev.vbits;
  info->min_value = t->min_value; info->max_value = t->max_value; info->value_offset = t->value_offset;
  info->n_kmers = t->n_kmers;
  info->n_level1 = t->host_stats.n_l1; info->n_level2 = t->host_stats.n_l2;
  info->n_stash = t->host_stats.n_stash; info->n_bigvalue = t->host_stats.n_big;
  info->buckets_level1 = t->dev.lv[0].buckets; info->buckets_level2 = t->dev.lv[1].buckets;
  info->device_bytes = t->device_bytes;
  return MFN_OK;
}

extern "C" int mfn_table_from_kmers(mfn_ctx *ctx, uint32_t k, const uint64_t *kmers, const uint32_t *values,
                                    uint64_t n, uint32_t minV, uint32_t maxV, mfn_table **out) {
  if (!ctx || !out || (n && !kmers)) return MFN_ERR_INVALID;
  if (minV == 0) minV = 1;                                        // kmers-exact.C:43
  if (maxV == 0xffffffffu) {                                      // kmers-exact.C:44: largest value present
    uint32_t m = 0;
    for (uint64_t i = 0; i < n; i++) m = std::max(m, values ? values[i] : 1u);
    maxV = m;
  }
  uint64_t keep = 0;
  for (uint64_t i = 0; i < n; i++) { uint32_t v = values ? values[i] : 1u; keep += (v >= minV && v <= maxV); }
  mfn_table *t = nullptr;
  MFN_TRY(mfn_table_alloc(ctx, k, keep, &t, [&](uint32_t vmax) {
    uint64_t c = 0;
    for (uint64_t i = 0; i < n; i++) { uint32_t v = values ? values[i] : 1u; c += (v >= minV && v <= maxV && v - (minV - 1) > vmax); }
    return c;
  }));
  t->min_value = minV; t->max_value = maxV; t->value_offset = minV - 1;
  const uint64_t CH = 1ull << 22;
  unsigned long long *dk = nullptr; uint32_t *dv = nullptr;
  int st = MFN_OK;
  if (n) {
    if (cudaMalloc(&dk, std::min(n, CH) * 8) != cudaSuccess || cudaMalloc(&dv, std::min(n, CH) * 4) != cudaSuccess) {
      cudaFree(dk); mfn_table_free(t); mfn_set_error("cudaMalloc staging"); return MFN_ERR_NOMEM;
    }
  }
  for (uint64_t o = 0; o < n && st == MFN_OK; o += CH) {
    uint64_t m = std::min(CH, n - o);
    cudaMemcpyAsync(dk, kmers + o, m * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (values) cudaMemcpyAsync(dv, values + o, m * 4, cudaMemcpyHostToDevice, ctx->stream);
    st = mfn_table_insert_device(t, dk, values ? dv : nullptr, m, minV, maxV, minV - 1, false);
    cudaStreamSynchronize(ctx->stream);
  }
  if (st == MFN_OK) st = mfn_table_finalize(t, false);
  cudaFree(dk); cudaFree(dv);
  if (st != MFN_OK) { mfn_table_free(t); return st; }
  *out = t;
  return MFN_OK;
}

extern "C" int mfn_table_lookup(mfn_ctx *ctx, const mfn_table *t, const uint64_t *kmers, uint64_t n, uint32_t *out) {
  if (!ctx || !t || (n && (!kmers || !out))) return MFN_ERR_INVALID;
  MFN_REJECT_WIDE(t, "mfn_table_lookup (use mfn_table_lookup_wide)");
  MFN_CUDA(cudaSetDevice(ctx->device));
  const uint64_t CH = 1ull << 22;
  unsigned long long *dk; uint32_t *dv;
  if (n == 0) return MFN_OK;
  MFN_CUDA(cudaMalloc(&dk, std::min(n, CH) * 8));
  MFN_CUDA(cudaMalloc(&dv, std::min(n, CH) * 4));
  for (uint64_t o = 0; o < n; o += CH) {
    uint64_t m = std::min(CH, n - o);
    cudaMemcpyAsync(dk, kmers + o, m * 8, cudaMemcpyHostToDevice, ctx->stream);
    k_lookup<<<grid_for(ctx, m), 256, 0, ctx->stream>>>(t->dev, dk, m, dv);
    ctx->launches++;
    cudaMemcpyAsync(out + o, dv, m * 4, cudaMemcpyDeviceToHost, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
  }
  cudaError_t e = cudaGetLastError();
  cudaFree(dk); cudaFree(dv);
  if (e != cudaSuccess) { mfn_set_error("lookup: %s", cudaGetErrorString(e)); return MFN_ERR_CUDA; }
  return MFN_OK;
}

extern "C" int mfn_table_export(mfn_ctx *ctx, const mfn_table *t, uint64_t *kmers, uint32_t *values,
                                uint64_t cap, uint64_t *n_out) {
  if (!ctx || !t || !n_out) return MFN_ERR_INVALID;
  MFN_REJECT_WIDE(t, "mfn_table_export");
  MFN_CUDA(cudaSetDevice(ctx->device));
  unsigned long long *dk = nullptr, *dn = nullptr; uint32_t *dv = nullptr;
  uint64_t c = std::max<uint64_t>(cap, 1);
  MFN_CUDA(cudaMalloc(&dk, c * 8)); MFN_CUDA(cudaMalloc(&dv, c * 4)); MFN_CUDA(cudaMalloc(&dn, 8));
  cudaMemsetAsync(dn, 0, 8, ctx->stream);
  k_export<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(t->dev, dk, dv, cap, dn);
  ctx->launches++;
  unsigned long long n = 0;
  cudaMemcpyAsync(&n, dn, 8, cudaMemcpyDeviceToHost, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  int st = MFN_OK;
  if (n > cap) st = MFN_ERR_CAPACITY;
  else if (n) {
    // ascending by k-mer (the order `meryl print` lists a database in): device radix sort, then
    // straight into the caller's arrays.  Infrastructure for tests / generators, not the hot path.
    unsigned long long *dk2 = nullptr; uint32_t *dv2 = nullptr; void *tmp = nullptr; size_t tmp_bytes = 0;
    cudaError_t e = cudaMalloc(&dk2, n * 8);
    if (e == cudaSuccess) e = cudaMalloc(&dv2, n * 4);
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dk2, dv, dv2, (int64_t)n, 0, 64, ctx->stream);
    if (e == cudaSuccess) e = cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1);
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, dk, dk2, dv, dv2, (int64_t)n, 0, 64, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(kmers, dk2, n * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(values, dv2, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(dk2); cudaFree(dv2); cudaFree(tmp);
    if (e != cudaSuccess) { mfn_set_error("export: %s", cudaGetErrorString(e)); st = MFN_ERR_CUDA; }
  }
  *n_out = n;
  cudaFree(dk); cudaFree(dv); cudaFree(dn);
  if (cudaGetLastError() != cudaSuccess) return MFN_ERR_CUDA;
  return st;
}

extern "C" int mfn_table_count_seq(mfn_ctx *ctx, uint32_t k, const mfn_seq *seq, mfn_table **out) {
  if (!ctx || !seq || !out) return MFN_ERR_INVALID;
  if (k > 32) return mfn_wide_count_seq(ctx, k, seq, out);                  // 128-bit k-mers: sort-based count (wide.cu)
  return mfn_table_count_seq_range(ctx, k, seq, 0, seq->G, out);
}

// The same over the windows that START in global positions [g_begin, g_end) (a shard's tile range).
int mfn_table_count_seq_range(mfn_ctx *ctx, uint32_t k, const mfn_seq *seq, uint64_t g_begin, uint64_t g_end, mfn_table **out) {
  if (!ctx || !seq || !out) return MFN_ERR_INVALID;
  g_end = std::min<uint64_t>(g_end, seq->G);
  g_begin = std::min(g_begin, g_end);
  const uint64_t span = g_end - g_begin;
  mfn_table *t = nullptr;
  // distinct k-mers <= windows; size for that upper bound
  MFN_TRY(mfn_table_alloc(ctx, k, span, &t));
  t->min_value = 1; t->max_value = 0xffffffffu; t->value_offset = 0;
  int grid = ctx->sm_count * 8;
  if (t->dev.w64) k_count_seq<true><<<grid, 256, 0, ctx->stream>>>(t->dev, seq->d_seq2, seq->d_ok, g_begin, g_end);
  else            k_count_seq<false><<<grid, 256, 0, ctx->stream>>>(t->dev, seq->d_seq2, seq->d_ok, g_begin, g_end);
  ctx->launches++;
  int st = (cudaGetLastError() == cudaSuccess) ? mfn_table_finalize(t, true) : MFN_ERR_CUDA;
  if (st != MFN_OK) { mfn_table_free(t); return st; }
  // The table was sized for one k-mer per base.  When far fewer are distinct (diploid or
  // repetitive assemblies), re-insert into a right-sized table: less HBM, fewer TLB misses.
  if (t->n_kmers * 10 < span * 8 && span > (1u << 20)) {
    mfn_table *c = nullptr;
    st = mfn_table_alloc(ctx, k, t->n_kmers, &c);
    if (st == MFN_OK) {
      c->min_value = 1; c->max_value = 0xffffffffu; c->value_offset = 0;
      if (c->dev.w64) k_copy_table<true><<<grid, 256, 0, ctx->stream>>>(c->dev, t->dev);
      else            k_copy_table<false><<<grid, 256, 0, ctx->stream>>>(c->dev, t->dev);
      ctx->launches++;
      st = (cudaGetLastError() == cudaSuccess) ? mfn_table_finalize(c, false) : MFN_ERR_CUDA;
    }
    if (st == MFN_OK && c->n_kmers != t->n_kmers) { mfn_set_error("table compaction lost entries"); st = MFN_ERR_CUDA; }
    if (st != MFN_OK) { mfn_table_free(c); mfn_table_free(t); return st; }
    mfn_table_free(t);
    t = c;
  }
  *out = t;
  return MFN_OK;
}

extern "C" int mfn_table_synth_reads(mfn_ctx *ctx, const mfn_seq *seq, const mfn_table *asm_, double cov,
                                     uint64_t n_error, uint64_t seed, mfn_table **out) {
  if (!ctx || !asm_ || !out) return MFN_ERR_INVALID;
  (void)seq;
  mfn_table *t = nullptr;
  MFN_TRY(mfn_table_alloc(ctx, asm_->dev.k, asm_->n_kmers + n_error, &t));
  t->min_value = 1; t->max_value = 0xffffffffu; t->value_offset = 0;
  int grid = ctx->sm_count * 8;
  if (t->dev.w64) {
    k_synth_from_asm<true><<<grid, 256, 0, ctx->stream>>>(t->dev, asm_->dev, cov, seed);
    if (n_error) k_synth_errors<true><<<grid, 256, 0, ctx->stream>>>(t->dev, n_error, seed);
  } else {
    k_synth_from_asm<false><<<grid, 256, 0, ctx->stream>>>(t->dev, asm_->dev, cov, seed);
    if (n_error) k_synth_errors<false><<<grid, 256, 0, ctx->stream>>>(t->dev, n_error, seed);
  }
  ctx->launches += n_error ? 2 : 1;
  int st = (cudaGetLastError() == cudaSuccess) ? mfn_table_finalize(t, false) : MFN_ERR_CUDA;
  if (st != MFN_OK) { mfn_table_free(t); return st; }
  *out = t;
  return MFN_OK;
}
