// Sequence sets: ASCII records -> 2-bit packed words + validity bitmap, resident in HBM.
//
// Global coordinate g runs over the concatenation of all records with ONE separator
// position after each record (never a valid base), so windows cannot span records and a
// single linear scan covers the whole set (kmers-iterator.H:115-155 semantics: a window is
// valid iff its k bases are all in ACGTacgt).
//   seq2[g/32] bits 2*(g%32)..+1 : (ascii >> 1) & 3  -> A=0 C=1 T=2 G=3 (kmers-tiny.H:69-83)
//   ok[g/64]   bit  g%64         : 1 iff the byte is one of ACGTacgt
#include "table_host.h"

#include <algorithm>

__device__ __forceinline__ bool is_acgt(unsigned c) {
  c |= 0x20u;
  return c == 'a' || c == 'c' || c == 'g' || c == 't';
}

// One thread packs the 64 global positions of one ok-word (two seq2 words) from the
// source bytes [s0, s1) staged at 'ascii'.  Positions whose source byte is outside the
// staged range are left untouched (atomicOr into zero-initialised arrays).
__global__ void __launch_bounds__(256) k_pack(const unsigned char *__restrict__ ascii, uint64_t s0, uint64_t s1,
                                              const unsigned long long *__restrict__ rec_start, uint32_t nseq,
                                              uint64_t w_begin, uint64_t w_end,
                                              unsigned long long *seq2, unsigned long long *ok) {
  uint64_t wi = w_begin + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (wi >= w_end) return;
  uint64_t g0 = wi * 64;
  // record containing g0: last r with rec_start[r] <= g0
  uint32_t lo = 0, hi = nseq;                    // rec_start has nseq+1 entries
  while (lo < hi) {
    uint32_t mid = (lo + hi + 1) >> 1;
    if (rec_start[mid] <= g0) lo = mid; else hi = mid - 1;
  }
  uint32_t r = lo;
  unsigned long long a = 0, b = 0, m = 0;
  uint64_t next = (r < nseq) ? rec_start[r + 1] : ~0ull;
  for (int i = 0; i < 64; i++) {
    uint64_t g = g0 + i;
    while (g >= next) { r++; next = (r < nseq) ? rec_start[r + 1] : ~0ull; }
    if (r >= nseq) break;                        // padding after the last record
    if (g == next - 1) continue;                 // the separator slot
    uint64_t s = g - r;                          // source byte (one separator per earlier record)
    if (s < s0 || s >= s1) continue;
    unsigned c = ascii[s - s0];
    if (!is_acgt(c)) continue;
    unsigned long long code = (c >> 1) & 3u;
    if (i < 32) a |= code << (2 * i); else b |= code << (2 * (i - 32));
    m |= 1ull << i;
  }
  if (a) atomicOr(&seq2[2 * wi], a);
  if (b) atomicOr(&seq2[2 * wi + 1], b);
  if (m) atomicOr(&ok[wi], m);
}

int mfn_seq_alloc(mfn_ctx *ctx, const uint64_t *seq_off, uint32_t nseq, mfn_seq **out, bool pooled) {
  MFN_CUDA(cudaSetDevice(ctx->device));
  mfn_seq *s = new mfn_seq();
  s->ctx = ctx;
  s->pooled = pooled;
  s->device = ctx->device;
  s->nseq = nseq;
  s->total_bases = nseq ? seq_off[nseq] - seq_off[0] : 0;
  s->rec_start.resize((size_t)nseq + 1);
  for (uint32_t i = 0; i <= nseq; i++) s->rec_start[i] = (seq_off[i] - seq_off[0]) + i;
  s->G = s->total_bases + nseq;
  s->ntiles = std::max<uint64_t>(1, ceil_div_u64(s->G, MFN_TILE));
  uint64_t gpad = (s->ntiles + 1) * MFN_TILE + 256;      // one spare tile + halo, all invalid
  size_t b2 = (gpad / 32 + 8) * 8, bo = (gpad / 64 + 8) * 8, br = ((size_t)nseq + 1) * 8;
  // pooled: a short-lived set (one mfn_dump / mfn_score_paths call) takes its arrays from the stream-ordered pool,
  // on the stream that fills them; no cudaMalloc / cudaFree (and the device-wide sync of the latter) per call
  cudaError_t e = pooled ? cudaMallocAsync(&s->d_seq2, b2, ctx->pack_stream) : cudaMalloc(&s->d_seq2, b2);
  if (e == cudaSuccess) e = pooled ? cudaMallocAsync(&s->d_ok, bo, ctx->pack_stream) : cudaMalloc(&s->d_ok, bo);
  if (e == cudaSuccess) e = pooled ? cudaMallocAsync(&s->d_rec_start, br, ctx->pack_stream) : cudaMalloc(&s->d_rec_start, br);
  if (e != cudaSuccess) { mfn_set_error("cudaMalloc(sequence): %s", cudaGetErrorString(e)); mfn_seq_free(s); return MFN_ERR_NOMEM; }
  s->device_bytes = b2 + bo + br;
  // the packed arrays are only ever written on the pack stream (see mfn_seq_pack_range)
  MFN_CUDA(cudaMemsetAsync(s->d_seq2, 0, b2, ctx->pack_stream));
  MFN_CUDA(cudaMemsetAsync(s->d_ok, 0, bo, ctx->pack_stream));
  MFN_CUDA(cudaMemcpyAsync(s->d_rec_start, s->rec_start.data(), br, cudaMemcpyHostToDevice, ctx->pack_stream));
  MFN_CUDA(cudaEventRecord(ctx->ev[MFN_STAGE_BUFS + 1], ctx->pack_stream));
  MFN_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev[MFN_STAGE_BUFS + 1], 0));
  *out = s;
  return MFN_OK;
}

static int ensure_staging(mfn_ctx *ctx, size_t bytes) {
  if (ctx->ascii_cap >= bytes) return MFN_OK;
  for (int i = 0; i < MFN_STAGE_BUFS; i++) { cudaFree(ctx->d_ascii[i]); ctx->d_ascii[i] = nullptr; }
  ctx->ascii_cap = 0;
  for (int i = 0; i < MFN_STAGE_BUFS; i++) MFN_CUDA(cudaMalloc(&ctx->d_ascii[i], bytes));
  ctx->ascii_cap = bytes;
  return MFN_OK;
}

// Upload and pack the source bytes that back global positions [g_begin, g_end).
// Three streams: H2D of 64 MB chunks into a ring of staging buffers (copy stream), k_pack of each
// chunk (pack stream), and the caller's compute stream, which is made to wait for the packed
// range at the end.  Neither uploads nor packing wait for scan kernels queued on the compute
// stream, so a caller that alternates pack_range / scan overlaps the next upload with the scan.
int mfn_seq_pack_range(mfn_seq *s, const char *bases, uint64_t g_begin, uint64_t g_end, cudaEvent_t done) {
  mfn_ctx *ctx = s->ctx;
  if (s->nseq == 0 || g_begin >= g_end) return MFN_OK;
  g_end = std::min(g_end, s->G);
  // source range: g - r for the record r containing g
  auto rec_of = [&](uint64_t g) { return (uint32_t)(std::upper_bound(s->rec_start.begin(), s->rec_start.end(), g) - s->rec_start.begin() - 1); };
  uint32_t ra = std::min(rec_of(g_begin), s->nseq - 1), rb = std::min(rec_of(g_end - 1), s->nseq - 1);
  uint64_t sa = std::min(g_begin - ra, s->total_bases), sb = std::min(g_end - rb, s->total_bases);
  if (sa >= sb) return MFN_OK;
  const size_t CH = 64ull << 20;
  MFN_TRY(ensure_staging(ctx, CH));
  cudaEvent_t ev_copy = ctx->ev[MFN_STAGE_BUFS], ev_pack = ctx->ev[MFN_STAGE_BUFS + 1];
  for (uint64_t o = sa; o < sb; o += CH) {
    const int buf = ctx->stage_next;
    ctx->stage_next = (ctx->stage_next + 1) % MFN_STAGE_BUFS;
    uint64_t e = std::min<uint64_t>(sb, o + CH);
    // the pack kernel that last read this buffer must be done before we overwrite it
    MFN_CUDA(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev[buf], 0));
    MFN_CUDA(cudaMemcpyAsync(ctx->d_ascii[buf], bases + o, e - o, cudaMemcpyHostToDevice, ctx->copy_stream));
    ctx->h2d_bytes += e - o;
    MFN_CUDA(cudaEventRecord(ev_copy, ctx->copy_stream));
    MFN_CUDA(cudaStreamWaitEvent(ctx->pack_stream, ev_copy, 0));
    // global words touched by source bytes [o, e)
    // g of a source byte x: x + r(x) where r is the record holding source byte x
    auto g_of_src = [&](uint64_t x) {
      // records are contiguous in the source: record r covers [rec_start[r]-r, rec_start[r+1]-r-1)
      uint32_t lo = 0, hi = s->nseq - 1;
      while (lo < hi) {
        uint32_t mid = (lo + hi + 1) >> 1;
        if (s->rec_start[mid] - mid <= x) lo = mid; else hi = mid - 1;
      }
      return x + lo;
    };
    uint64_t w0 = g_of_src(o) / 64, w1 = g_of_src(e - 1) / 64 + 1;
    uint64_t nw = w1 - w0;
    k_pack<<<(unsigned)ceil_div_u64(nw, 256), 256, 0, ctx->pack_stream>>>(
        (const unsigned char *)ctx->d_ascii[buf], o, e, s->d_rec_start, s->nseq, w0, w1, s->d_seq2, s->d_ok);
    ctx->launches++;
    MFN_CUDA(cudaEventRecord(ctx->ev[buf], ctx->pack_stream));
  }
  MFN_CUDA(cudaGetLastError());
  if (done) { MFN_CUDA(cudaEventRecord(done, ctx->pack_stream)); return MFN_OK; }
  MFN_CUDA(cudaEventRecord(ev_pack, ctx->pack_stream));
  MFN_CUDA(cudaStreamWaitEvent(ctx->stream, ev_pack, 0));
  return MFN_OK;
}

extern "C" int mfn_seq_create(mfn_ctx *ctx, const char *bases, const uint64_t *seq_off, uint32_t nseq, mfn_seq **out) {
  return mfn_seq_create_in(ctx, bases, seq_off, nseq, out, false);
}

int mfn_seq_create_in(mfn_ctx *ctx, const char *bases, const uint64_t *seq_off, uint32_t nseq, mfn_seq **out, bool pooled) {
  if (!ctx || !out || !seq_off || (nseq && !bases)) return MFN_ERR_INVALID;
  for (uint32_t i = 0; i < nseq; i++)
    if (seq_off[i + 1] < seq_off[i]) { mfn_set_error("seq_off must be non-decreasing"); return MFN_ERR_INVALID; }
  mfn_seq *s = nullptr;
  MFN_TRY(mfn_seq_alloc(ctx, seq_off, nseq, &s, pooled));
  int st = mfn_seq_pack_range(s, bases + seq_off[0], 0, s->G);
  if (st == MFN_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) { mfn_set_error("pack failed: %s", cudaGetErrorString(cudaGetLastError())); st = MFN_ERR_CUDA; }
  if (st != MFN_OK) { mfn_seq_free(s); return st; }
  *out = s;
  return MFN_OK;
}

extern "C" int mfn_seq_info(const mfn_seq *s, uint64_t *total_bases, uint32_t *nseq, uint64_t *device_bytes) {
  if (!s) return MFN_ERR_INVALID;
  if (total_bases) *total_bases = s->total_bases;
  if (nseq) *nseq = s->nseq;
  if (device_bytes) *device_bytes = s->device_bytes;
  return MFN_OK;
}

extern "C" void mfn_seq_free(mfn_seq *s) {
  if (!s) return;
  cudaSetDevice(s->device);
  if (s->pooled) {
    // (freed in stream order behind the scans the caller queued on the compute stream)
    if (s->d_seq2) cudaFreeAsync(s->d_seq2, s->ctx->stream);
    if (s->d_ok) cudaFreeAsync(s->d_ok, s->ctx->stream);
    if (s->d_rec_start) cudaFreeAsync(s->d_rec_start, s->ctx->stream);
  } else {
    cudaFree(s->d_seq2); cudaFree(s->d_ok); cudaFree(s->d_rec_start);
  }
  delete s;
}
