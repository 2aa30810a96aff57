// k-mers of 33..64 bases (the reference's kmdata is 128 bits wide, kmers-tiny.H:29): a second, plain path
// beside the bucketed tables and the range-partitioned scan, which are built around 64-bit k-mers.
//
//   table    the database's k-mers as they are stored -- ascending -- in two arrays (high / low 64 bits) with their
//            values, plus an index over the top bits of the k-mer that narrows a lookup to a few entries; a lookup
//            is a binary search of that range (merylExactLookup::value, kmers-lookup.H:334-380);
//   scan     the tile staging of k_scan (TMA bulk copies, double-buffered), 8 windows per thread cut out of a
//            192-bit register window; every window looks up its forward AND its reverse-complement k-mer in both
//            tables and adds the two values, which is literally what merfinGlobal::getK does (merfin-globals.C:
//            100-111): a canonical database holds one of the two, a palindrome counts twice (Q2);
//   join     -completeness: one lookup of every read k-mer in the assembly table (merfin-completeness.C:48-144);
//   count    the assembly's own k-mers when no -seqmers is given: canonical k-mers of all windows, two stable
//            radix sorts (low half, then high half), run lengths.
//
// The integer results go into the same accumulator / the same per-position arrays as for k <= 32, so everything
// after the probes (K* replay, reports, variant selection) is shared.
#include "scan_kernel.cuh"
#include "wide.cuh"

#include <cub/cub.cuh>

#include <thread>

struct mfn_wide { DevWide dev{}; };

// ---- table build --------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_wide_check(DevWide T, unsigned long long *bad) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x + 1; i < T.n; i += (uint64_t)gridDim.x * blockDim.x) {
    const bool less = T.hi[i - 1] < T.hi[i] || (T.hi[i - 1] == T.hi[i] && T.lo[i - 1] < T.lo[i]);
    if (!less) atomicAdd(bad, 1ull);
  }
}

__global__ void __launch_bounds__(256) k_wide_index(DevWide T, unsigned long long *index) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i <= T.n; i += (uint64_t)gridDim.x * blockDim.x)
    wide_index_fill(T, index, i);
}

__global__ void __launch_bounds__(256) k_wide_lookup(DevWide T, const unsigned long long *hi, const unsigned long long *lo, uint64_t n, uint32_t *out) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t h = hi[i], l = lo[i];
    const bool fits = T.k == 64 || (h >> (2 * T.k - 64)) == 0;
    out[i] = fits ? wide_value(T, h, l) : 0u;
  }
}

static int wide_grid(mfn_ctx *ctx, uint64_t n) {
  return (int)std::min<uint64_t>(ceil_div_u64(std::max<uint64_t>(n, 1), 256), (uint64_t)ctx->sm_count * 8);
}

// The table over n sorted k-mers already in device arrays (ownership of the arrays passes to the table).
static int wide_finish(mfn_ctx *ctx, uint32_t k, unsigned long long *dhi, unsigned long long *dlo, uint32_t *dval, uint64_t n,
                       uint64_t cap, mfn_table **out) {
  mfn_table *t = new mfn_table();
  t->ctx = ctx; t->device = ctx->device;
  t->dev.k = k;
  t->allocs.push_back(dhi); t->allocs.push_back(dlo); t->allocs.push_back(dval);
  t->wide = new mfn_wide();
  DevWide &W = t->wide->dev;
  W.hi = dhi; W.lo = dlo; W.val = dval; W.n = n; W.k = k;
  const uint32_t ib = wide_index_bits(n);
  W.ibits = ib;
  unsigned long long *index = nullptr, *bad = nullptr;
  const uint64_t index_bytes = ((1ull << ib) + 1) * 8;
  cudaError_t e = cudaMalloc(&index, index_bytes);
  if (e == cudaSuccess) { t->allocs.push_back(index); e = cudaMalloc(&bad, 8); }
  if (e == cudaSuccess) e = cudaMemsetAsync(bad, 0, 8, ctx->stream);
  unsigned long long nbad = 0;
  if (e == cudaSuccess) {
    k_wide_check<<<wide_grid(ctx, n), 256, 0, ctx->stream>>>(W, bad);
    k_wide_index<<<wide_grid(ctx, n + 1), 256, 0, ctx->stream>>>(W, index);
    ctx->launches += 2;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(&nbad, bad, 8, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(bad);
  if (e != cudaSuccess) { mfn_set_error("k=%u table: %s", k, cudaGetErrorString(e)); mfn_table_free(t); return e == cudaErrorMemoryAllocation ? MFN_ERR_NOMEM : MFN_ERR_CUDA; }
  if (nbad) {
    mfn_set_error("the k-mers of the database are not in ascending order (%llu places): a multiset or damaged database; not supported", nbad);
    mfn_table_free(t);
    return MFN_ERR_UNSUPPORTED;
  }
  W.index = index;
  t->n_kmers = n;
  t->device_bytes = cap * 20 + index_bytes;
  *out = t;
  return MFN_OK;
}

// mfn_table_load for k > 32: decode on host threads (a group of data files at a time), keep the values inside
// [minV, maxV] minus the offset (kmers-exact.C:578-613), append in file order = ascending k-mer order.
int mfn_wide_load(mfn_ctx *ctx, MerylIndex &idx, uint32_t minV, uint32_t maxV, mfn_table **out) {
  MFN_CUDA(cudaSetDevice(ctx->device));
  if (idx.version != 3) {
    // older master indexes carry a different statistics layout: take the value histogram from the data
    std::vector<uint64_t> cnt;
    for (uint32_t f = 0; f < (1u << idx.numFilesBits); f++)
      MFN_TRY(meryl_decode_file_wide(idx, f, [&](const uint64_t *, const uint64_t *, const uint32_t *v, uint64_t n) {
        for (uint64_t i = 0; i < n; i++) { if (v[i] >= cnt.size()) cnt.resize((size_t)v[i] + 1, 0); cnt[v[i]]++; }
      }));
    idx.histValue.clear(); idx.histCount.clear();
    for (size_t v = 0; v < cnt.size(); v++) if (cnt[v]) { idx.histValue.push_back(v); idx.histCount.push_back(cnt[v]); }
  }
  if (minV == 0) minV = 1;                                                 // kmers-exact.C:43
  if (maxV == 0xffffffffu) maxV = idx.histValue.empty() ? 0 : (uint32_t)idx.histValue.back();   // :44, kmers-histogram.H:73
  uint64_t keep = 0;                                                       // kmers-exact.C:64-72
  for (size_t i = 0; i < idx.histValue.size(); i++)
    if (idx.histValue[i] >= minV && idx.histValue[i] <= maxV) keep += idx.histCount[i];
  unsigned long long *dhi = nullptr, *dlo = nullptr; uint32_t *dval = nullptr;
  const uint64_t cap = std::max<uint64_t>(keep, 1);
  if (cudaMalloc(&dhi, cap * 8) != cudaSuccess || cudaMalloc(&dlo, cap * 8) != cudaSuccess || cudaMalloc(&dval, cap * 4) != cudaSuccess) {
    cudaFree(dhi); cudaFree(dlo); cudaFree(dval); cudaGetLastError();
    mfn_set_error("device memory for %llu k-mers of %u bases", (unsigned long long)keep, idx.k);
    return MFN_ERR_NOMEM;
  }
  const uint32_t nfiles = 1u << idx.numFilesBits;
  const unsigned nthreads = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  const uint32_t voff = minV - 1;
  uint64_t fill = 0;
  int status = MFN_OK;
  for (uint32_t f0 = 0; f0 < nfiles && status == MFN_OK; f0 += nthreads) {
    const uint32_t f1 = std::min(nfiles, f0 + nthreads);
    std::vector<std::vector<uint64_t>> H(f1 - f0), L(f1 - f0);
    std::vector<std::vector<uint32_t>> V(f1 - f0);
    std::vector<int> st(f1 - f0, MFN_OK);
    std::vector<std::string> errs(f1 - f0);
    std::vector<std::thread> th;
    for (uint32_t f = f0; f < f1; f++)
      th.emplace_back([&, f]() {
        auto &h = H[f - f0]; auto &l = L[f - f0]; auto &v = V[f - f0];
        st[f - f0] = meryl_decode_file_wide(idx, f, [&](const uint64_t *hh, const uint64_t *ll, const uint32_t *vv, uint64_t n) {
          for (uint64_t i = 0; i < n; i++)
            if (vv[i] >= minV && vv[i] <= maxV) { h.push_back(hh[i]); l.push_back(ll[i]); v.push_back(vv[i] - voff); }
        });
        if (st[f - f0] != MFN_OK) errs[f - f0] = mfn_last_error();
      });
    for (auto &t : th) t.join();
    for (uint32_t i = 0; i < f1 - f0 && status == MFN_OK; i++) {
      if (st[i] != MFN_OK) { status = st[i]; mfn_set_error("%s", errs[i].c_str()); break; }
      const uint64_t m = H[i].size();
      if (m == 0) continue;
      if (fill + m > cap) { mfn_set_error("'%s': more k-mers in the data files than its value histogram announces", idx.dir.c_str()); status = MFN_ERR_FORMAT; break; }
      cudaError_t e = cudaMemcpy(dhi + fill, H[i].data(), m * 8, cudaMemcpyHostToDevice);
      if (e == cudaSuccess) e = cudaMemcpy(dlo + fill, L[i].data(), m * 8, cudaMemcpyHostToDevice);
      if (e == cudaSuccess) e = cudaMemcpy(dval + fill, V[i].data(), m * 4, cudaMemcpyHostToDevice);
      if (e != cudaSuccess) { mfn_set_error("k-mer upload: %s", cudaGetErrorString(e)); status = MFN_ERR_CUDA; break; }
      fill += m;
    }
  }
  if (status != MFN_OK) { cudaFree(dhi); cudaFree(dlo); cudaFree(dval); return status; }
  mfn_table *t = nullptr;
  MFN_TRY(wide_finish(ctx, idx.k, dhi, dlo, dval, fill, cap, &t));     // (frees the arrays on failure)
  t->min_value = minV; t->max_value = maxV; t->value_offset = voff;
  *out = t;
  return MFN_OK;
}

void mfn_wide_free(mfn_wide *w) { delete w; }

void mfn_wide_info(const mfn_table *t, mfn_table_info_t *info) {
  info->k = t->dev.k;
  info->canonical = 0;                         // both orientations of every window are looked up, whatever the database holds
  info->slot_bits = 128;
  info->tag_bits = 2 * t->dev.k;
  info->value_bits = 32;
  info->n_level1 = t->n_kmers;
  info->buckets_level1 = (1ull << t->wide->dev.ibits);
}

extern "C" int mfn_table_lookup_wide(mfn_ctx *ctx, const mfn_table *t, const uint64_t *kmers_lo, const uint64_t *kmers_hi, uint64_t n, uint32_t *out) {
  if (!ctx || !t || (n && (!kmers_lo || !kmers_hi || !out))) return MFN_ERR_INVALID;
  if (!t->wide) { mfn_set_error("mfn_table_lookup_wide: a table of k <= 32 (use mfn_table_lookup)"); return MFN_ERR_INVALID; }
  if (n == 0) return MFN_OK;
  MFN_CUDA(cudaSetDevice(ctx->device));
  unsigned long long *dh = nullptr, *dl = nullptr; uint32_t *dv = nullptr;
  cudaError_t e = cudaMalloc(&dh, n * 8);
  if (e == cudaSuccess) e = cudaMalloc(&dl, n * 8);
  if (e == cudaSuccess) e = cudaMalloc(&dv, n * 4);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dh, kmers_hi, n * 8, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dl, kmers_lo, n * 8, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) {
    k_wide_lookup<<<wide_grid(ctx, n), 256, 0, ctx->stream>>>(t->wide->dev, dh, dl, n, dv);
    ctx->launches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, dv, n * 4, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cudaFree(dh); cudaFree(dl); cudaFree(dv);
  if (e != cudaSuccess) { mfn_set_error("lookup: %s", cudaGetErrorString(e)); return MFN_ERR_CUDA; }
  return MFN_OK;
}

// ---- scan -----------------------------------------------------------------------------------------
struct WideScanParams {
  DevWide R, A;
  const unsigned long long *seq2, *ok, *rec_start;
  uint32_t nseq;
  uint64_t tile_begin, tile_end;
  unsigned long long *tile_counter;        // zeroed before launch
  // -hist
  unsigned long long *acc;
  PairTable sparse;
  const unsigned long long *zmap;          // bit v set: readK(v) == 0 (v < zlen); NULL -> only v == 0
  uint32_t zlen;
  // -dump
  uint32_t *out_vr, *out_va;
  uint8_t *out_valid;
};

template <bool DUMP>
__global__ void __launch_bounds__(SCAN_THREADS, 1) k_scan_wide(const __grid_constant__ WideScanParams P) {
  __shared__ __align__(16) unsigned long long s_seq[2][SEQ_WORDS];
  __shared__ __align__(16) unsigned long long s_ok[2][OK_WORDS];
  __shared__ __align__(8) unsigned long long s_bar[2];
  __shared__ unsigned long long s_tile[2];
  __shared__ uint32_t s_hist[DUMP ? 1 : SH_R * SH_A];

  const uint32_t tid = threadIdx.x;
  const uint32_t k = P.R.k;                                       // 33 .. 64
  const uint64_t hmask = wide_hmask(k), kb = wide_kbits(k);

  if (!DUMP) for (uint32_t i = tid; i < SH_R * SH_A; i += SCAN_THREADS) s_hist[i] = 0;
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto fetch = [&](int buf) {                // thread 0: claim the next tile and start its TMA
    unsigned long long t = P.tile_begin + atomicAdd(P.tile_counter, 1ull);
    s_tile[buf] = t;
    if (t < P.tile_end) {
      mbar_expect_tx(&s_bar[buf], (SEQ_WORDS + OK_WORDS) * 8);
      tma_load_1d(s_seq[buf], P.seq2 + t * (MFN_TILE / 32), SEQ_WORDS * 8, &s_bar[buf]);
      tma_load_1d(s_ok[buf], P.ok + t * (MFN_TILE / 64), OK_WORDS * 8, &s_bar[buf]);
    }
  };
  if (tid == 0) fetch(0);

  unsigned long long my_valid = 0;           // windows this thread counted (for acc[0])
  uint32_t cur_rec = 0xffffffffu, n_ok0 = 0, n_miss0 = 0;
  auto flush_rec = [&]() {
    if (DUMP) return;
    uint32_t a = n_ok0, b = n_miss0;
    for (int o = 16; o > 0; o >>= 1) { a += __shfl_down_sync(0xffffffffu, a, o); b += __shfl_down_sync(0xffffffffu, b, o); }
    if ((tid & 31) == 0 && cur_rec < P.nseq) {
      if (a) atomicAdd(&P.acc[MFN_ACC_HEADER + MFN_DENSE_R * MFN_DENSE_A + 2 * cur_rec], (unsigned long long)a);
      if (b) atomicAdd(&P.acc[MFN_ACC_HEADER + MFN_DENSE_R * MFN_DENSE_A + 2 * cur_rec + 1], (unsigned long long)b);
    }
    n_ok0 = 0; n_miss0 = 0;
  };

  for (uint32_t it = 0;; it++) {
    const int buf = it & 1;
    __syncthreads();                         // s_tile[buf] visible; everyone is done with buf^1
    const uint64_t tile = s_tile[buf];
    if (tile >= P.tile_end) break;
    if (tid == 0) fetch(buf ^ 1);
    mbar_wait(&s_bar[buf], (it >> 1) & 1);

    // this thread's 8 windows: bases j0 .. j0 + 7 + 63 of the tile = a 192-bit slice (the tile carries a 64-base halo),
    // validity bits j0 .. j0 + 70
    const uint32_t j0 = tid * WPT;
    const uint64_t g0 = tile * MFN_TILE + j0;
    const WideSlice S = wide_slice(s_seq[buf], s_ok[buf], j0);

    uint32_t r0;
    {
      const uint64_t gt = tile * MFN_TILE;
      uint32_t a = 0, b = P.nseq;
      while (a < b) { uint32_t m = (a + b + 1) >> 1; if (P.rec_start[m] <= gt) a = m; else b = m - 1; }
      r0 = a;
    }
    if (r0 != cur_rec) { flush_rec(); cur_rec = r0; }        // uniform across the CTA
    uint32_t rec = r0;
    uint64_t next_start = (rec < P.nseq) ? P.rec_start[rec + 1] : ~0ull;

#pragma unroll 1
    for (int uu = 0; uu < WPT; uu++) {
      uint64_t xh, xl;
      const bool valid = wide_window(S, uu, hmask, kb, xh, xl);
      uint32_t vr = 0, va = 0;
      if (valid) wide_pair_of(P.R, P.A, k, hmask, xh, xl, vr, va);
      const uint64_t g = g0 + uu;
      while (g >= next_start) { rec++; next_start = (rec < P.nseq) ? P.rec_start[rec + 1] : ~0ull; }
      if (DUMP) {
        // positions are reported in the caller's coordinates: one separator per earlier record
        if (rec < P.nseq && g + 1 < next_start) {
          const uint64_t s = g - rec;
          P.out_vr[s] = vr; P.out_va[s] = va; P.out_valid[s] = valid ? 1 : 0;
        }
      } else if (valid) {
        bool miss = (vr == 0);
        if (vr != 0 && vr < P.zlen) miss = (__ldg(&P.zmap[vr >> 6]) >> (vr & 63)) & 1;
        if (rec == r0) { n_ok0++; n_miss0 += miss; }
        else {
          atomicAdd(&P.acc[MFN_ACC_HEADER + MFN_DENSE_R * MFN_DENSE_A + 2 * rec], 1ull);
          if (miss) atomicAdd(&P.acc[MFN_ACC_HEADER + MFN_DENSE_R * MFN_DENSE_A + 2 * rec + 1], 1ull);
        }
        my_valid++;
        if (vr < SH_R && va < SH_A) atomicAdd(&s_hist[va * SH_R + vr], 1u);
        else if (vr < MFN_DENSE_R && va < MFN_DENSE_A) atomicAdd(&P.acc[MFN_ACC_HEADER + va * MFN_DENSE_R + vr], 1ull);
        else sparse_add(P.sparse, P.acc, vr, va, 1ull);
      }
    }
  }

  flush_rec();
  if (!DUMP) {
    __syncthreads();
    for (uint32_t i = tid; i < SH_R * SH_A; i += SCAN_THREADS) {
      uint32_t c = s_hist[i];
      if (c) atomicAdd(&P.acc[MFN_ACC_HEADER + (i / SH_R) * MFN_DENSE_R + (i % SH_R)], (unsigned long long)c);
    }
    for (int o = 16; o > 0; o >>= 1) my_valid += __shfl_down_sync(0xffffffffu, my_valid, o);
    if ((tid & 31) == 0 && my_valid) atomicAdd(&P.acc[0], my_valid);
  }
}

template <bool DUMP>
static int wide_launch(mfn_ctx *ctx, WideScanParams &P) {
  if (P.tile_begin >= P.tile_end) return MFN_OK;
  unsigned long long *counter;
  MFN_CUDA(cudaMallocAsync(&counter, 8, ctx->stream));
  MFN_CUDA(cudaMemsetAsync(counter, 0, 8, ctx->stream));
  P.tile_counter = counter;
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_scan_wide<DUMP>, SCAN_THREADS, 0);
  if (per_sm < 1) per_sm = 1;
  const int grid = (int)std::min<uint64_t>((uint64_t)ctx->sm_count * per_sm, P.tile_end - P.tile_begin);
  k_scan_wide<DUMP><<<grid, SCAN_THREADS, 0, ctx->stream>>>(P);
  ctx->launches++;
  MFN_CUDA(cudaGetLastError());
  MFN_CUDA(cudaFreeAsync(counter, ctx->stream));
  return MFN_OK;
}

static int wide_pair(const mfn_table *R, const mfn_table *A) {
  if (!R->wide || !A->wide || R->dev.k != A->dev.k) { mfn_set_error("mer size mismatch: reads k=%u, assembly k=%u", R->dev.k, A->dev.k); return MFN_ERR_K_MISMATCH; }
  return MFN_OK;
}

int mfn_wide_hist_tiles(mfn_ctx *ctx, const mfn_table *reads, const mfn_table *asm_, const mfn_seq *seq, const mfn_kmodel *km,
                        uint64_t tile_begin, uint64_t tile_end, mfn_accum *acc) {
  MFN_TRY(wide_pair(reads, asm_));
  WideScanParams P{};
  P.R = reads->wide->dev; P.A = asm_->wide->dev;
  P.seq2 = seq->d_seq2; P.ok = seq->d_ok; P.rec_start = seq->d_rec_start; P.nseq = seq->nseq;
  P.tile_begin = tile_begin; P.tile_end = std::min(tile_end, seq->ntiles);
  P.acc = acc->d_buf; P.sparse = acc->sparse;
  unsigned long long *zmap = nullptr;
  uint32_t zhi = 0;
  MFN_TRY(mfn_build_zmap(ctx, km, &zmap, &P.zlen, &zhi));
  P.zmap = zmap;
  const int st = wide_launch<false>(ctx, P);
  if (zmap) cudaFreeAsync(zmap, ctx->stream);
  return st;
}

int mfn_wide_dump(mfn_ctx *ctx, const mfn_table *reads, const mfn_table *asm_, const mfn_seq *seq, uint32_t *d_v_read,
                  uint32_t *d_v_asm, uint8_t *d_valid) {
  MFN_TRY(wide_pair(reads, asm_));
  WideScanParams P{};
  P.R = reads->wide->dev; P.A = asm_->wide->dev;
  P.seq2 = seq->d_seq2; P.ok = seq->d_ok; P.rec_start = seq->d_rec_start; P.nseq = seq->nseq;
  P.tile_begin = 0; P.tile_end = seq->ntiles;
  P.out_vr = d_v_read; P.out_va = d_v_asm; P.out_valid = d_valid;
  if (seq->total_bases) {
    MFN_CUDA(cudaMemsetAsync(d_v_read, 0, seq->total_bases * 4, ctx->stream));
    MFN_CUDA(cudaMemsetAsync(d_v_asm, 0, seq->total_bases * 4, ctx->stream));
    MFN_CUDA(cudaMemsetAsync(d_valid, 0, seq->total_bases, ctx->stream));
  }
  return wide_launch<true>(ctx, P);
}

// ---- -completeness ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_join_wide(DevWide A, const unsigned long long *hi, const unsigned long long *lo, const uint32_t *values,
                                                   uint64_t n, unsigned long long *acc, PairTable sparse) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t vr = values[i];
    const uint32_t va = wide_value(A, hi[i], lo[i]);
    if (vr < MFN_DENSE_R && va < MFN_DENSE_A) atomicAdd(&acc[MFN_ACC_HEADER + va * MFN_DENSE_R + vr], 1ull);
    else sparse_add(sparse, acc, vr, va, 1ull);
  }
}

// data files [file_begin, file_end) of the read database joined with the raw assembly table
int mfn_wide_completeness(mfn_ctx *ctx, const MerylIndex &idx, const mfn_table *asm_raw, uint32_t file_begin, uint32_t file_end, mfn_accum *acc) {
  if (!asm_raw->wide || idx.k != asm_raw->dev.k) { mfn_set_error("mer size mismatch: reads k=%u, assembly k=%u", idx.k, asm_raw->dev.k); return MFN_ERR_K_MISMATCH; }
  MFN_CUDA(cudaSetDevice(ctx->device));
  const uint32_t nfiles = std::min(file_end, 1u << idx.numFilesBits);
  std::vector<uint64_t> H, L;
  std::vector<uint32_t> V;
  for (uint32_t f = file_begin; f < nfiles; f++) {
    H.clear(); L.clear(); V.clear();
    MFN_TRY(meryl_decode_file_wide(idx, f, [&](const uint64_t *hh, const uint64_t *ll, const uint32_t *vv, uint64_t n) {
      H.insert(H.end(), hh, hh + n); L.insert(L.end(), ll, ll + n); V.insert(V.end(), vv, vv + n);
    }));
    const uint64_t n = H.size();
    if (n == 0) continue;
    unsigned long long *dh = nullptr, *dl = nullptr; uint32_t *dv = nullptr;
    cudaError_t e = cudaMalloc(&dh, n * 8);
    if (e == cudaSuccess) e = cudaMalloc(&dl, n * 8);
    if (e == cudaSuccess) e = cudaMalloc(&dv, n * 4);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dh, H.data(), n * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dl, L.data(), n * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dv, V.data(), n * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
      k_join_wide<<<wide_grid(ctx, n), 256, 0, ctx->stream>>>(asm_raw->wide->dev, dh, dl, dv, n, acc->d_buf, acc->sparse);
      ctx->launches++;
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(dh); cudaFree(dl); cudaFree(dv);
    if (e != cudaSuccess) { mfn_set_error("-completeness (k=%u): %s", idx.k, cudaGetErrorString(e)); return MFN_ERR_CUDA; }
  }
  return MFN_OK;
}

// every k-mer of a resident read table against the assembly table (mfn_completeness_resident)
int mfn_wide_completeness_resident(mfn_ctx *ctx, const mfn_table *reads_raw, const mfn_table *asm_raw, mfn_accum *acc) {
  MFN_TRY(wide_pair(reads_raw, asm_raw));
  const DevWide &R = reads_raw->wide->dev;
  if (R.n == 0) return MFN_OK;
  k_join_wide<<<wide_grid(ctx, R.n), 256, 0, ctx->stream>>>(asm_raw->wide->dev, R.hi, R.lo, R.val, R.n, acc->d_buf, acc->sparse);
  ctx->launches++;
  MFN_CUDA(cudaGetLastError());
  return MFN_OK;
}

// ---- counting the k-mers of a sequence set (no -seqmers) ----------------------------------------
// canonical k-mer of every valid window that starts in [g_begin, g_end), in no particular order
__global__ void __launch_bounds__(256) k_wide_windows(const unsigned long long *seq2, const unsigned long long *ok, uint64_t g_begin, uint64_t g_end,
                                                      uint32_t k, unsigned long long *out_hi, unsigned long long *out_lo, unsigned long long *count) {
  const uint64_t hmask = wide_hmask(k), kb = wide_kbits(k);
  for (uint64_t g = g_begin + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; g < g_end; g += (uint64_t)gridDim.x * blockDim.x) {
    // validity bits g .. g + 63 and bases g .. g + 63 (three packed words cover 64 bases from any offset)
    const uint64_t wo = g >> 6, so = g & 63;
    const uint64_t vb = so ? ((ok[wo] >> so) | (ok[wo + 1] << (64 - so))) : ok[wo];
    if ((vb & kb) != kb) continue;
    const uint64_t w = g >> 5, sh = (g & 31) * 2;
    const uint64_t a = seq2[w], b = seq2[w + 1], c = seq2[w + 2];
    const uint64_t xl = sh ? ((a >> sh) | (b << (64 - sh))) : a;
    const uint64_t xh = (sh ? ((b >> sh) | (c << (64 - sh))) : b) & hmask;
    uint64_t fh, fl, rh, rl;
    wide_kmers(k, hmask, xh, xl, fh, fl, rh, rl);
    const bool fwd = fh < rh || (fh == rh && fl <= rl);
    const unsigned long long at = atomicAdd(count, 1ull);
    out_hi[at] = fwd ? fh : rh;
    out_lo[at] = fwd ? fl : rl;
  }
}

// over sorted k-mers: head[i] = 1 where a run of equal k-mers starts
__global__ void __launch_bounds__(256) k_wide_heads(const unsigned long long *hi, const unsigned long long *lo, uint64_t n, unsigned long long *head) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    head[i] = (i == 0 || hi[i] != hi[i - 1] || lo[i] != lo[i - 1]) ? 1ull : 0ull;
}

// rank[i] = inclusive sum of head: the run of entry i is distinct k-mer rank[i] - 1
__global__ void __launch_bounds__(256) k_wide_runs(const unsigned long long *hi, const unsigned long long *lo, const unsigned long long *rank, uint64_t n,
                                                   unsigned long long *out_hi, unsigned long long *out_lo, uint32_t *out_val) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t r = rank[i] - 1;
    const bool first = i == 0 || rank[i - 1] != rank[i];
    if (first) { out_hi[r] = hi[i]; out_lo[r] = lo[i]; }
    atomicAdd(&out_val[r], 1u);
  }
}

int mfn_wide_count_seq(mfn_ctx *ctx, uint32_t k, const mfn_seq *seq, mfn_table **out) {
  if (k < 33 || k > 64) { mfn_set_error("k=%u: k-mers hold at most 64 bases", k); return MFN_ERR_UNSUPPORTED; }
  MFN_CUDA(cudaSetDevice(ctx->device));
  const uint64_t G = seq->G;
  const uint64_t cap = std::max<uint64_t>(G, 1);
  unsigned long long *h0 = nullptr, *l0 = nullptr, *h1 = nullptr, *l1 = nullptr, *cnt = nullptr, *rank = nullptr;
  void *tmp = nullptr;
  unsigned long long *dhi = nullptr, *dlo = nullptr; uint32_t *dval = nullptr;
  auto cleanup = [&]() { cudaFree(h0); cudaFree(l0); cudaFree(h1); cudaFree(l1); cudaFree(cnt); cudaFree(rank); cudaFree(tmp); };
  cudaError_t e = cudaMalloc(&h0, cap * 8);
  if (e == cudaSuccess) e = cudaMalloc(&l0, cap * 8);
  if (e == cudaSuccess) e = cudaMalloc(&h1, cap * 8);
  if (e == cudaSuccess) e = cudaMalloc(&l1, cap * 8);
  if (e == cudaSuccess) e = cudaMalloc(&cnt, 8);
  if (e == cudaSuccess) e = cudaMemsetAsync(cnt, 0, 8, ctx->stream);
  unsigned long long n = 0;
  if (e == cudaSuccess && G) {
    k_wide_windows<<<wide_grid(ctx, G), 256, 0, ctx->stream>>>(seq->d_seq2, seq->d_ok, 0, G, k, h0, l0, cnt);
    ctx->launches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(&n, cnt, 8, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  unsigned long long ndist = 0;
  if (e == cudaSuccess && n) {
    // two stable passes: by the low half, then by the high half
    size_t tb1 = 0, tb2 = 0, tb3 = 0;
    e = cub::DeviceRadixSort::SortPairs(nullptr, tb1, l0, l1, h0, h1, (int64_t)n, 0, 64, ctx->stream);
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(nullptr, tb2, h1, h0, l1, l0, (int64_t)n, 0, (int)(2 * k - 64), ctx->stream);
    if (e == cudaSuccess) e = cudaMalloc(&rank, n * 8);
    if (e == cudaSuccess) e = cub::DeviceScan::InclusiveSum(nullptr, tb3, rank, rank, (int64_t)n, ctx->stream);
    const size_t tb = std::max(std::max(tb1, tb2), std::max<size_t>(tb3, 16));
    if (e == cudaSuccess) e = cudaMalloc(&tmp, tb);
    size_t t1 = tb, t2 = tb, t3 = tb;
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, t1, l0, l1, h0, h1, (int64_t)n, 0, 64, ctx->stream);
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, t2, h1, h0, l1, l0, (int64_t)n, 0, (int)(2 * k - 64), ctx->stream);
    // sorted k-mers are now in (h0, l0)
    if (e == cudaSuccess) {
      k_wide_heads<<<wide_grid(ctx, n), 256, 0, ctx->stream>>>(h0, l0, n, rank);
      ctx->launches += 3;
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cub::DeviceScan::InclusiveSum(tmp, t3, rank, rank, (int64_t)n, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&ndist, rank + (n - 1), 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  }
  const uint64_t dcap = std::max<uint64_t>(ndist, 1);
  if (e == cudaSuccess) e = cudaMalloc(&dhi, dcap * 8);
  if (e == cudaSuccess) e = cudaMalloc(&dlo, dcap * 8);
  if (e == cudaSuccess) e = cudaMalloc(&dval, dcap * 4);
  if (e == cudaSuccess) e = cudaMemsetAsync(dval, 0, dcap * 4, ctx->stream);
  if (e == cudaSuccess && n) {
    k_wide_runs<<<wide_grid(ctx, n), 256, 0, ctx->stream>>>(h0, l0, rank, n, dhi, dlo, dval);
    ctx->launches++;
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  cleanup();
  if (e != cudaSuccess) {
    cudaFree(dhi); cudaFree(dlo); cudaFree(dval);
    mfn_set_error("counting %u-mers: %s", k, cudaGetErrorString(e));
    return e == cudaErrorMemoryAllocation ? MFN_ERR_NOMEM : MFN_ERR_CUDA;
  }
  mfn_table *t = nullptr;
  MFN_TRY(wide_finish(ctx, k, dhi, dlo, dval, ndist, dcap, &t));
  t->min_value = 1; t->max_value = 0xffffffffu; t->value_offset = 0;
  *out = t;
  return MFN_OK;
}
