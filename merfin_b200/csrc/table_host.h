// Host-side entry points shared between translation units of libmerfin_b200.
#pragma once

#include "common.h"

#include <functional>

// count_big(vmax): how many stored values exceed vmax (sizes the excess table); null -> a guess
int mfn_table_alloc(mfn_ctx *ctx, uint32_t k, uint64_t n_expected, mfn_table **out,
                    const std::function<uint64_t(uint32_t)> &count_big = nullptr);
int mfn_table_insert_device(mfn_table *t, const unsigned long long *d_kmers, const uint32_t *d_values,
                            uint64_t n, uint32_t minV, uint32_t maxV, uint32_t voff, bool add);
int mfn_table_finalize(mfn_table *t, bool counted);

// ---- meryl on-disk database (host) ------------------------------------------
struct MerylIndex {
  uint32_t version = 0, prefixSize = 0, suffixSize = 0, numFilesBits = 0, numBlocksBits = 0, flags = 0;
  uint32_t k = 0;
  uint64_t numUnique = 0, numDistinct = 0, numTotal = 0;
  std::vector<uint64_t> histValue, histCount;     // ascending by value (v03)
  std::string dir;
};

int meryl_open(const char *dir, MerylIndex &idx);
std::string meryl_file_name(const MerylIndex &idx, uint32_t file, bool index);
// Decode data file 'file' block by block; sink(kmers, values, n) is called once per block.
// k-mers are 64-bit (k <= 32).
int meryl_decode_file(const MerylIndex &idx, uint32_t file,
                      const std::function<void(const uint64_t *, const uint32_t *, uint64_t)> &sink);
// Any k <= 64: sink(hi, lo, values, n) with the two halves of every 128-bit k-mer.
int meryl_decode_file_wide(const MerylIndex &idx, uint32_t file,
                           const std::function<void(const uint64_t *, const uint64_t *, const uint32_t *, uint64_t)> &sink);
// Stream a whole database to the device in chunks; fn(d_kmers, d_values, n) is called per chunk
// on ctx->stream with device arrays that stay valid until the next call.
int meryl_stream_to_device(mfn_ctx *ctx, const MerylIndex &idx,
                           const std::function<int(const unsigned long long *, const uint32_t *, uint64_t)> &fn,
                           uint32_t file_begin = 0, uint32_t file_end = 0xffffffffu);

// The same with the decode on the device (meryl_gpu.cu): the file images are uploaded as they are
// on disk and one thread per meryl block decodes its chain of codes.
int meryl_stream_to_device_gpu(mfn_ctx *ctx, const MerylIndex &idx,
                               const std::function<int(const unsigned long long *, const uint32_t *, uint64_t)> &fn,
                               uint32_t file_begin = 0, uint32_t file_end = 0xffffffffu);
// the decoder in use: the device one unless MFN_HOST_DECODE=1
inline int meryl_stream(mfn_ctx *ctx, const MerylIndex &idx,
                        const std::function<int(const unsigned long long *, const uint32_t *, uint64_t)> &fn,
                        uint32_t file_begin = 0, uint32_t file_end = 0xffffffffu) {
  const char *env = getenv("MFN_HOST_DECODE");
  return (env && atoi(env)) ? meryl_stream_to_device(ctx, idx, fn, file_begin, file_end)
                            : meryl_stream_to_device_gpu(ctx, idx, fn, file_begin, file_end);
}

// ---- k > 32 (wide.cu) -----------------------------------------------------------------
int mfn_wide_load(mfn_ctx *ctx, MerylIndex &idx, uint32_t minV, uint32_t maxV, mfn_table **out);
void mfn_wide_free(mfn_wide *w);
void mfn_wide_info(const mfn_table *t, mfn_table_info_t *info);
int mfn_wide_hist_tiles(mfn_ctx *ctx, const mfn_table *reads, const mfn_table *asm_, const mfn_seq *seq, const mfn_kmodel *km,
                        uint64_t tile_begin, uint64_t tile_end, mfn_accum *acc);
int mfn_wide_dump(mfn_ctx *ctx, const mfn_table *reads, const mfn_table *asm_, const mfn_seq *seq, uint32_t *d_v_read,
                  uint32_t *d_v_asm, uint8_t *d_valid);
int mfn_wide_completeness(mfn_ctx *ctx, const MerylIndex &idx, const mfn_table *asm_raw, uint32_t file_begin, uint32_t file_end, mfn_accum *acc);
int mfn_wide_completeness_resident(mfn_ctx *ctx, const mfn_table *reads_raw, const mfn_table *asm_raw, mfn_accum *acc);
int mfn_wide_count_seq(mfn_ctx *ctx, uint32_t k, const mfn_seq *seq, mfn_table **out);

// ---- sequences ------------------------------------------------------------------
int mfn_seq_alloc(mfn_ctx *ctx, const uint64_t *seq_off, uint32_t nseq, mfn_seq **out, bool pooled = false);
// mfn_seq_create; pooled: device arrays from the stream-ordered pool (a set that lives for one call; freed while its ctx is alive)
int mfn_seq_create_in(mfn_ctx *ctx, const char *bases, const uint64_t *seq_off, uint32_t nseq, mfn_seq **out, bool pooled);
// done == nullptr: the compute stream is made to wait for the packed range; otherwise `done` is
// recorded on the pack stream instead and the caller orders its own streams
int mfn_seq_pack_range(mfn_seq *s, const char *bases, uint64_t g_begin, uint64_t g_end, cudaEvent_t done = nullptr);
// hybrid ingest (ingest.cu)
struct mfn_ingest;
int mfn_ingest_begin(mfn_ctx *ctx, mfn_seq *seq, const char *bases, uint64_t tile_begin, uint64_t tile_end,
                     uint64_t piece_tiles, uint32_t *npieces);
int mfn_ingest_feed(mfn_ctx *ctx, uint32_t upto);
int mfn_ingest_wait(mfn_ctx *ctx, uint32_t piece);
void mfn_ingest_end(mfn_ctx *ctx);
void mfn_ingest_free(mfn_ingest *I);
// -hist scan of tiles [tile_begin, tile_end) of a resident sequence set
int mfn_hist_tiles(mfn_ctx *ctx, const mfn_table *reads, const mfn_table *asm_, const mfn_seq *seq,
                   const mfn_kmodel *km, uint64_t tile_begin, uint64_t tile_end, mfn_accum *acc);

// ---- fused (read, assembly) table ----------------------------------------------------------
#include "joint.cuh"
struct mfn_joint {
  JointDev dev{};
  JointStats stats{};
  const mfn_table *reads = nullptr, *asm_ = nullptr;
  uint64_t n_kmers = 0, device_bytes = 0;
  uint32_t tag_bits = 0;
  bool restricted = false;          // holds only the k-mers of one sequence shard: misses fall back to the per-database tables
  std::vector<void *> allocs;
};
int mfn_joint_build(mfn_ctx *ctx, const mfn_table *R, const mfn_table *A, mfn_joint **out);
int mfn_joint_build_restricted(mfn_ctx *ctx, const mfn_table *R, const mfn_table *A, const mfn_table *keys, mfn_joint **out);
int mfn_table_count_seq_range(mfn_ctx *ctx, uint32_t k, const mfn_seq *seq, uint64_t g_begin, uint64_t g_end, mfn_table **out);
void mfn_joint_free(mfn_joint *j);

// ---- range-partitioned -hist scan (sweep.cu) --------------------------------------------------
bool mfn_sweep_applicable(const mfn_table *R, const mfn_table *A, uint64_t ntiles, uint32_t nseq);
int mfn_sweep_hist(mfn_ctx *ctx, const mfn_table *R, const mfn_table *A, const mfn_seq *seq, const mfn_kmodel *km,
                   uint64_t tile_begin, uint64_t tile_end, mfn_accum *acc, const unsigned long long *zmap, uint32_t zlen,
                   uint32_t zhi);
cudaEvent_t mfn_ctx_mark(mfn_ctx *c);
struct mfn_sweep;
int mfn_sweep_begin(mfn_ctx *ctx, const mfn_table *R, const mfn_table *A, const mfn_seq *seq, uint64_t windows, mfn_accum *acc,
                    const unsigned long long *zmap, uint32_t zlen, uint32_t zhi, mfn_sweep **out);
int mfn_sweep_scatter(mfn_ctx *ctx, mfn_sweep *w, uint64_t tile_begin, uint64_t tile_end);
int mfn_sweep_finish(mfn_ctx *ctx, mfn_sweep *w, bool side_stream = false);
void mfn_sweep_join(mfn_ctx *ctx);
// -dump through the range-partitioned scan (sweep_dump.cu / sweep.cu)
bool mfn_sweep_dump_applicable(const mfn_table *R, const mfn_table *A, uint64_t ntiles, bool force);
int mfn_sweep_dump_piece(mfn_ctx *ctx, const mfn_table *R, const mfn_table *A, const mfn_seq *seq, uint64_t tile_begin, uint64_t tile_end,
                         uint32_t *d_vr, uint32_t *d_va, uint8_t *d_valid, int *ok);
int mfn_sweep_dump(mfn_ctx *ctx, const mfn_table *R, const mfn_table *A, const mfn_seq *seq, uint32_t *d_vr, uint32_t *d_va,
                   uint8_t *d_valid, uint64_t *done_tiles);
int mfn_build_zmap(mfn_ctx *ctx, const mfn_kmodel *km, unsigned long long **d_zmap, uint32_t *zlen, uint32_t *zhi);
