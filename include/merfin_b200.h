/*
 * merfin_b200.h -- C ABI of the B200-native per-base k-mer multiplicity path.
 *
 * This is the drop-in boundary for merfin's -hist / -dump / -completeness inner
 * loops and the -polish/-filter path scoring.  Plain C: opaque handles, plain
 * pointers and sizes.  No C++ or torch types cross it.  Every entry point
 * returns MFN_OK (0) or a negative mfn_status; nothing here calls exit().
 *
 * One mfn_ctx drives ONE CUDA device (one process per GPU).  Multi-GPU runs
 * shard the tiles of a sequence set across ranks (shard, nshards arguments),
 * each rank accumulates into an mfn_accum whose device buffer is a flat
 * uint64 array that the caller sum-all-reduces (NCCL) before mfn_accum_fetch.
 *
 * There is NO CPU fallback: without a CUDA device mfn_ctx_create fails with
 * MFN_ERR_NO_DEVICE and every other call needs a ctx.
 *
 * Citations (file:line) are relative to the reference tree
 * (arangrhie/merfin @ 04c4332e); each entry point names the reference
 * interface it replaces.
 */
#ifndef MERFIN_B200_H
#define MERFIN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  MFN_OK                 =  0,
  MFN_ERR_NO_DEVICE      = -1,   /* no usable CUDA device (there is no CPU path)        */
  MFN_ERR_CUDA           = -2,   /* a CUDA call failed; see mfn_last_error()            */
  MFN_ERR_IO             = -3,   /* file missing / unreadable                           */
  MFN_ERR_FORMAT         = -4,   /* not a meryl database / corrupt block                */
  MFN_ERR_UNSUPPORTED    = -5,   /* k > 64, multiset database, a 64-bit-only call on a k > 32 table */
  MFN_ERR_INVALID        = -6,   /* bad argument                                        */
  MFN_ERR_CAPACITY       = -7,   /* caller buffer or internal side table too small      */
  MFN_ERR_K_MISMATCH     = -8,   /* tables built for different k (kmers-reader.C:171)   */
  MFN_ERR_NOMEM          = -9
} mfn_status;

typedef struct mfn_ctx   mfn_ctx;     /* one device, its streams and staging buffers      */
typedef struct mfn_table mfn_table;   /* one k-mer -> value table resident in HBM         */
typedef struct mfn_seq   mfn_seq;     /* a set of sequences, 2-bit packed, resident       */
typedef struct mfn_accum mfn_accum;   /* joint (v_read, v_asm) histogram + record counters */

const char *mfn_strerror(int status);
const char *mfn_last_error(void);          /* thread-local detail of the last failure */
const char *mfn_version(void);

/* ---- context ------------------------------------------------------------ */

/* A ctx = one device + its own streams, events and staging buffers.  Several contexts may exist on one
 * device; tables, sequence sets and VCFs are plain device / host objects that any ctx OF THE SAME DEVICE
 * may read (scans never write them), so a host that wants the device parts of concurrent mfn_dump /
 * mfn_score_paths / mfn_variants_process calls to overlap gives each group of threads its own ctx over
 * one set of tables (bench.py --mode polish does: "scorer lanes"). */
int  mfn_ctx_create(mfn_ctx **ctx, int device);
void mfn_ctx_destroy(mfn_ctx *ctx);
int  mfn_ctx_device(const mfn_ctx *ctx);
/* Run all subsequent work of this ctx on an existing CUDA stream (cudaStream_t
 * passed as void*; NULL = the ctx's own stream).  Lets a host framework that
 * owns the device (e.g. torch) order our kernels with its collectives. */
int  mfn_ctx_set_stream(mfn_ctx *ctx, void *cuda_stream);
int  mfn_ctx_sync(mfn_ctx *ctx);
/* Pinned host memory for callers that want full-rate H2D/D2H of their buffers. */
int  mfn_host_alloc(void **p, size_t bytes);
void mfn_host_free(void *p);

/* ---- tables: replaces merylExactLookup::load / ::value -------------------
 * kmers-lookup.H:118-137, kmers-exact.C:34-73 (filter, offset), 648-671.
 *
 * Mer sizes: 1 <= k <= 64, like the reference's 128-bit kmdata (kmers-tiny.H:29).  Up to k = 32 a
 * k-mer is one 64-bit word and the tables are the bucketed ones all the fast paths are built on.
 * For k = 33..64 mfn_table_load / mfn_table_count_seq build a table of sorted 128-bit k-mers
 * (csrc/wide.cu); mfn_hist*, mfn_dump*, mfn_score_paths, mfn_variants_process and mfn_completeness*
 * take such tables and give the reference's numbers through a plain scan that looks every window up
 * in both orientations; the entry points that exist only for 64-bit k-mers (mfn_table_from_kmers,
 * mfn_table_lookup, mfn_table_export, mfn_tables_fuse / _restrict*, mfn_joint_lookup, mfn_meryl_write,
 * mfn_table_synth_reads) return MFN_ERR_UNSUPPORTED for them, and mfn_table_info reports
 * canonical = 0 and slot_bits = 128.                                                             */

/* Load an on-disk meryl database (formats v01-v03; kmers-reader.C:127-206,
 * kmers-files.C:24-128).  minV/maxV follow merfin's -min/-max: 0 -> 1,
 * UINT32_MAX -> largest value in the DB; entries outside are dropped and the
 * stored value is v-(minV-1), which is what value() returns (Q1). */
int  mfn_table_load(mfn_ctx *ctx, const char *meryl_dir, uint32_t minV, uint32_t maxV,
                    mfn_table **out);
/* Build from host arrays of (k-mer, value); k-mers use meryl's 2-bit code
 * A=0 C=1 T=2 G=3, first base in the high bits (kmers-tiny.H:69-83). */
int  mfn_table_from_kmers(mfn_ctx *ctx, uint32_t k, const uint64_t *kmers, const uint32_t *values,
                          uint64_t n, uint32_t minV, uint32_t maxV, mfn_table **out);
/* Count the canonical k-mers of a resident sequence set straight into a table:
 * replaces the system("meryl count ...") merfin runs when -seqmers is absent
 * (merfin-globals.C:168-189). */
int  mfn_table_count_seq(mfn_ctx *ctx, uint32_t k, const mfn_seq *seq, mfn_table **out);
/* Synthetic read database for benchmarks: for every distinct canonical k-mer of
 * seq with assembly count a (taken from asm_), value = Poisson(cov_per_copy * a)
 * drawn from a hash of the k-mer (zeros are dropped -> "missing"), plus n_error
 * pseudo-random error k-mers with value 1..2.  Deterministic in seed. */
int  mfn_table_synth_reads(mfn_ctx *ctx, const mfn_seq *seq, const mfn_table *asm_,
                           double cov_per_copy, uint64_t n_error, uint64_t seed, mfn_table **out);

typedef struct {
  uint32_t k;
  uint32_t canonical;        /* 1: every stored k-mer <= its reverse complement   */
  uint32_t slot_bits;        /* 32 or 64                                          */
  uint32_t tag_bits, value_bits;
  uint32_t min_value, max_value, value_offset;
  uint64_t n_kmers;          /* entries stored                                    */
  uint64_t n_level1, n_level2, n_stash, n_bigvalue;
  uint64_t buckets_level1, buckets_level2;
  uint64_t device_bytes;
} mfn_table_info_t;

int  mfn_table_info(const mfn_table *t, mfn_table_info_t *info);
void mfn_table_free(mfn_table *t);
/* Batch merylExactLookup::value(kmer) (kmers-lookup.H:334-380): out[i] = stored
 * value of kmers[i] exactly as given (no canonicalisation), 0 if absent. */
int  mfn_table_lookup(mfn_ctx *ctx, const mfn_table *t, const uint64_t *kmers, uint64_t n,
                      uint32_t *out);
/* The same for a table of k = 33..64: k-mer i is kmers_hi[i] << 64 | kmers_lo[i]. */
int  mfn_table_lookup_wide(mfn_ctx *ctx, const mfn_table *t, const uint64_t *kmers_lo, const uint64_t *kmers_hi,
                           uint64_t n, uint32_t *out);
/* Export every stored (k-mer, stored value), ascending k-mer order. */
int  mfn_table_export(mfn_ctx *ctx, const mfn_table *t, uint64_t *kmers, uint32_t *values,
                      uint64_t cap, uint64_t *n);

/* ---- fused (read, assembly) table ------------------------------------------------------
 * merfin always looks a k-mer up in BOTH databases (merfin-globals.C:107-108).  Fusing the
 * two tables into one whose 128-byte lines hold tag + read value + assembly value makes that
 * ONE memory access per window.  The fused table is attached to `reads`; every scan entry
 * point that is handed (reads, asm_) uses it automatically.  Needs canonical databases. */
int  mfn_tables_fuse(mfn_ctx *ctx, mfn_table *reads, const mfn_table *asm_);
int  mfn_tables_unfuse(mfn_table *reads);

typedef struct {
  uint32_t fused;            /* 0: no fused table attached                          */
  uint32_t slot_bits, tag_bits, read_bits, asm_bits;
  uint64_t n_kmers;          /* size of the union                                   */
  uint64_t n_home, n_spill, n_stash, n_bigvalue;
  uint64_t lines, device_bytes;
} mfn_joint_info_t;

int  mfn_joint_info(const mfn_table *reads, mfn_joint_info_t *info);
/* The fused table RESTRICTED to the k-mers of one shard of a resident sequence set (shard, nshards
 * as in mfn_hist_resident / mfn_hist_host): a -hist scan only ever asks for k-mers of the sequences
 * it scans, so this is the table the range-partitioned scan streams -- 1.5 G instead of 3.5 G entries
 * at the 3 Gbp benchmark, 1/N of that again when N ranks each restrict to their own shard.  It plays
 * the part of the -min/-max filter of merylExactLookup::load (kmers-exact.C:34-73): entries no
 * lookup of this run can ask for are not loaded.  Results stay exact for ANY sequence: a window
 * whose k-mer is not in the restricted table is looked up in the per-database tables; only the
 * speed depends on scanning the sequences the table was built for.  Attached to `reads` next to
 * the full fused table (which mfn_dump / mfn_score_paths keep using); needs canonical databases. */
int  mfn_tables_restrict(mfn_ctx *ctx, mfn_table *reads, const mfn_table *asm_, const mfn_seq *seq,
                         uint32_t shard, uint32_t nshards);
int  mfn_tables_unrestrict(mfn_table *reads);
int  mfn_restricted_info(const mfn_table *reads, mfn_joint_info_t *info);
/* The same table built from a read database that is STREAMED through the device and never resident
 * (a database of twice the HBM, BASELINE configs[4]; kmers-exact.C:95-146 gives up when the table does
 * not fit -memory): its blocks are decoded chunk by chunk and semi-joined on the device against the
 * k-mers of the sequence shard; read-only k-mers (two thirds of a 30x database) are dropped on the
 * spot, so the table never holds more than one entry per distinct k-mer of the shard.  *reads_out is
 * a read-table handle WITHOUT a per-database table behind it: it serves -hist (mfn_hist_resident /
 * mfn_hist_host, always as the range-partitioned scan) of the sequences it was restricted to; windows of
 * other sequences would read as absent from the reads, and mfn_dump / mfn_score_paths refuse it.
 * min_value / max_value as in mfn_table_load.  mfn_tables_restrict_synth is the benchmark's stand-in
 * for a database of any size: the synthetic read database of mfn_table_synth_reads (same draws),
 * its n_error error k-mers generated in chunks of chunk_kmers and pushed through the same semi-join. */
typedef struct {
  uint64_t db_kmers;       /* k-mers the database holds (streamed) */
  uint64_t kept_kmers;     /* of those, k-mers of the shard that went into the table */
  uint64_t shard_kmers;    /* distinct k-mers of the sequence shard */
  uint64_t table_bytes;    /* device memory of the restricted table */
  double   seconds;        /* wall clock of the call */
} mfn_stream_stats;
int  mfn_tables_restrict_stream(mfn_ctx *ctx, const char *reads_dir, uint32_t min_value, uint32_t max_value,
                                const mfn_table *asm_, const mfn_seq *seq, uint32_t shard, uint32_t nshards,
                                mfn_table **reads_out, mfn_stream_stats *stats /* or NULL */);
int  mfn_tables_restrict_synth(mfn_ctx *ctx, const mfn_table *asm_, const mfn_seq *seq, uint32_t shard, uint32_t nshards,
                               double cov_per_copy, uint64_t n_error, uint64_t seed, uint64_t chunk_kmers,
                               mfn_table **reads_out, mfn_stream_stats *stats /* or NULL */);
/* (value_read(kmer), value_asm(kmer)) for k-mers exactly as given, through the fused table. */
int  mfn_joint_lookup(mfn_ctx *ctx, const mfn_table *reads, const uint64_t *kmers, uint64_t n,
                      uint32_t *v_read, uint32_t *v_asm);

/* ---- meryl database writer (generator side; kmers-writer.C:117-288) ------ */
/* Writes a v03 database the reference's merylFileReader loads.  k-mers must be
 * distinct; they are sorted here. */
int  mfn_meryl_write(const char *meryl_dir, uint32_t k, const uint64_t *kmers,
                     const uint32_t *values, uint64_t n);

/* ---- meryl database header (merylFileReader's constructor, kmers-reader.C:128-205, and the
 * statistics block of a v03 master index, kmers-histogram.C:170-200); host only, no GPU work. */
typedef struct {
  uint32_t  k, version;               /* mer size; master-index format 1..3          */
  uint32_t  prefix_bits, suffix_bits; /* split of the 2k bits                        */
  uint32_t  files, blocks_per_file;   /* 2^numFilesBits, 2^numBlocksBits             */
  uint32_t  multiset, hist_len;       /* flags bit 0; value-histogram entries (v03)  */
  uint64_t  n_unique, n_distinct, n_total;   /* v03 only, else 0                     */
} mfn_meryl_info_t;
int  mfn_meryl_info(const char *meryl_dir, mfn_meryl_info_t *info);
/* The stored value histogram (value, number of distinct k-mers with it), ascending value;
 * MFN_ERR_CAPACITY with *n = entries needed when cap is too small. */
int  mfn_meryl_histogram(const char *meryl_dir, uint64_t *values, uint64_t *counts, uint64_t cap, uint64_t *n);
/* Every (k-mer, value) of a database in stored (ascending k-mer) order, decoded on the host
 * (merylFileBlockReader::decodeBlock, kmers-files.C:82-128): k <= 64, a k-mer is kmers_hi << 64 | kmers_lo
 * (kmers_hi may be NULL for k <= 32).  All arrays NULL: *n = the number of k-mers.  MFN_ERR_CAPACITY
 * with *n = entries needed when cap is too small. */
int  mfn_meryl_read(const char *meryl_dir, uint64_t *kmers_lo, uint64_t *kmers_hi, uint32_t *values, uint64_t cap, uint64_t *n);

/* ---- sequences: replaces dnaSeq + kmerIterator::nextBase ------------------
 * kmers-iterator.H:93-160.  bases = concatenated ASCII records, record i is
 * bases[seq_off[i] .. seq_off[i+1]).  Any byte outside ACGTacgt breaks windows;
 * windows never span records. */
int  mfn_seq_create(mfn_ctx *ctx, const char *bases, const uint64_t *seq_off, uint32_t nseq,
                    mfn_seq **out);
int  mfn_seq_info(const mfn_seq *s, uint64_t *total_bases, uint32_t *nseq, uint64_t *device_bytes);
void mfn_seq_free(mfn_seq *s);

/* ---- K* model pieces the device needs ------------------------------------
 * merfin-globals.C:66-98: readK==0 <=> v_read==0, or v_read within the -prob
 * table and its readK column is 0.  Only that predicate is evaluated on the
 * device (for per-record kmissing); all floating point stays on the host. */
typedef struct {
  double          peak;
  uint32_t        nprob;
  const uint32_t *probK;
  const double   *probP;
} mfn_kmodel;

/* ---- accumulators --------------------------------------------------------- */

typedef struct { uint64_t kasm, kmissing; } mfn_seq_counts;
typedef struct { uint32_t v_read, v_asm; uint64_t count; } mfn_pair_count;

int  mfn_accum_create(mfn_ctx *ctx, uint32_t nseq, mfn_accum **out);
int  mfn_accum_reset(mfn_ctx *ctx, mfn_accum *a);
void mfn_accum_free(mfn_accum *a);
/* Flat device buffer (uint64 words) holding the dense joint histogram (v_read < 4096, v_asm < 64)
 * and the per-record counters: sum-all-reduce it across ranks, then fetch / read it on any rank. */
int  mfn_accum_device_buffer(const mfn_accum *a, void **dptr, uint64_t *nwords);
/* Single rank: everything the accumulator holds -- per-record counters, the dense tier and the pairs
 * outside it, which live in a rank-local side table.  After an all-reduce of the flat buffer the
 * header (word [1]) counts the side-table pairs of EVERY rank while this call only sees its own:
 * it then fails with MFN_ERR_INVALID instead of dropping the others (include_dense = 0 fetches
 * just this rank's side table). */
int  mfn_accum_fetch(mfn_ctx *ctx, const mfn_accum *a, int include_dense, int include_sparse,
                     mfn_seq_counts *per_seq /*[nseq] or NULL*/,
                     mfn_pair_count *pairs, uint64_t pairs_cap, uint64_t *npairs);
/* Multi-rank recipe: all-reduce the flat buffer; mfn_accum_read it (D2H of nwords words); when its
 * word [1] is not 0, every rank calls mfn_accum_sparse_fetch and the lists are gathered (all-gather of
 * the counts, then of the padded lists); mfn_accum_decode_gathered sums equal pairs of different ranks
 * and refuses a gather that does not hold exactly word [1] entries.  A real assembly (satellites, Alu:
 * v_asm >= 64) needs this; the synthetic benchmark never leaves the dense tier. */
int  mfn_accum_read(mfn_ctx *ctx, const mfn_accum *a, uint64_t *host_buf /*[nwords]*/);
int  mfn_accum_sparse_fetch(mfn_ctx *ctx, const mfn_accum *a, mfn_pair_count *pairs, uint64_t pairs_cap,
                            uint64_t *npairs /* this rank's count, also when pairs_cap is too small */);

/* One process driving several devices: dst += src over the flat buffer, copied device to device
 * (NVLink between peers) and added by a kernel on dst's device, on dst's stream; src's ctx must be
 * idle (mfn_ctx_sync).  The single-process counterpart of the all-reduce above; pairs outside the
 * dense tier are gathered as described there (mfn_accum_sparse_fetch per device). */
int  mfn_accum_add_from(mfn_ctx *dst_ctx, mfn_accum *dst, const mfn_accum *src);

/* Layout of the flat buffer and host-side decoding of a copy of it (no GPU needed): lets a
 * caller all-reduce with any transport and decode on any rank.  mfn_accum_decode refuses a buffer
 * whose header counts side-table pairs (MFN_ERR_INVALID): use mfn_accum_decode_gathered. */
int  mfn_accum_layout(uint32_t nseq, uint32_t *header_words, uint32_t *dense_r, uint32_t *dense_a,
                      uint64_t *nwords);
int  mfn_accum_decode(const uint64_t *host_buf, uint32_t nseq, mfn_seq_counts *per_seq,
                      mfn_pair_count *pairs, uint64_t pairs_cap, uint64_t *npairs);
int  mfn_accum_decode_gathered(const uint64_t *host_buf, uint32_t nseq,
                               const mfn_pair_count *sparse, uint64_t n_sparse /* all ranks' lists, concatenated */,
                               mfn_seq_counts *per_seq, mfn_pair_count *pairs, uint64_t pairs_cap, uint64_t *npairs);
/* Global base range [g_begin, g_end) (records concatenated with one separator slot after
 * each) whose windows shard `shard` of `nshards` scans. */
int  mfn_shard_tiles(uint64_t total_bases, uint32_t nseq, uint32_t shard, uint32_t nshards,
                     uint64_t *g_begin, uint64_t *g_end);

/* ---- the hot path ---------------------------------------------------------- */

/* -hist inner loop (merfin-histogram.C:54-91 + merfin-globals.C:101-110) over the
 * tiles of 'seq' owned by (shard, nshards), resident inputs, asynchronous on the
 * ctx stream.  Adds into 'acc'. */
int  mfn_hist_resident(mfn_ctx *ctx, const mfn_table *reads, const mfn_table *asm_,
                       const mfn_seq *seq, const mfn_kmodel *km /*NULL: v_read==0 only*/,
                       uint32_t shard, uint32_t nshards, mfn_accum *acc);

/* The same over HOST buffers for one shard: uploads (pinned memory gives full PCIe rate) and
 * packs only the bases behind the shard's tiles, overlapping H2D with the scan.  Adds into acc;
 * the caller all-reduces and fetches. */
int  mfn_hist_host(mfn_ctx *ctx, const mfn_table *reads, const mfn_table *asm_,
                   const char *bases, const uint64_t *seq_off, uint32_t nseq, const mfn_kmodel *km,
                   uint32_t shard, uint32_t nshards, mfn_accum *acc);

/* Whole -hist call with HOST buffers: H2D + pack + probe + fetch.  Replaces
 * loadSequence/processHistogram/outputHistogram (merfin.C:30-53,
 * merfin-histogram.C:35-136) for one batch of records. */
int  mfn_hist(mfn_ctx *ctx, const mfn_table *reads, const mfn_table *asm_,
              const char *bases, const uint64_t *seq_off, uint32_t nseq, const mfn_kmodel *km,
              mfn_seq_counts *per_seq /*[nseq]*/,
              mfn_pair_count *pairs, uint64_t pairs_cap, uint64_t *npairs);

/* -dump inner loop (merfin-dump.C:44-67): per window start position of every
 * record (index = seq_off[i] + p), the integer pair and validity.  Positions
 * with no window (last k-1 of a record) get valid=0. */
int  mfn_dump_resident(mfn_ctx *ctx, const mfn_table *reads, const mfn_table *asm_,
                       const mfn_seq *seq, uint32_t *d_v_read, uint32_t *d_v_asm, uint8_t *d_valid);
int  mfn_dump(mfn_ctx *ctx, const mfn_table *reads, const mfn_table *asm_,
              const char *bases, const uint64_t *seq_off, uint32_t nseq,
              uint32_t *v_read /*[total bases]*/, uint32_t *v_asm, uint8_t *valid);

/* Variant path scoring (varMer.C:66-90 probes): same as mfn_dump over a batch of
 * short path strings.  mfn_dump and mfn_score_paths may be called from several host threads on
 * one ctx (the device part of a call holds the ctx's lock); output buffers in pinned memory
 * (mfn_host_alloc) are filled at full PCIe speed. */
int  mfn_score_paths(mfn_ctx *ctx, const mfn_table *reads, const mfn_table *asm_,
                     const char *paths, const uint64_t *path_off, uint64_t npaths,
                     uint32_t *v_read, uint32_t *v_asm, uint8_t *valid);

/* ---- -filter / -polish / -better / -strict / -loose (merfin-variants.C:131-310) --------------
 * The reference's unit of work is processVariants(G, T, S) for one sequence record S: cluster the
 * record's variants, enumerate the allele combinations of every cluster, score every path with
 * varMer::score and let the mode's selector write VCF records.  The same here: mfn_vcf_load =
 * vcfFile::load (vcf.C:93-151), mfn_vcf_cluster = mergeChrPosGT (vcf.C:156-246),
 * mfn_variants_process = processVariants for one record (its probes are batched through
 * mfn_score_paths).  Records are independent: call mfn_variants_process from as many host threads
 * as there are cores.  shard / nshards split the clusters of the record into consecutive ranges
 * (rank r of N scores range r of every record, no collective; the texts of ranks 0..N-1 concatenate
 * to the unsharded result -- the job-array grain of scripts/parallel1/_submit_merfin.sh, made finer
 * than one contig). */
typedef struct mfn_vcf mfn_vcf;
typedef struct { uint64_t headers, records, contigs, contig_ids, excluded, clusters; } mfn_vcf_info_t;
enum { MFN_MODE_FILTER = 4, MFN_MODE_POLISH = 5, MFN_MODE_BETTER = 6, MFN_MODE_STRICT = 7, MFN_MODE_LOOSE = 8 };  /* merfin-globals.H:32-40 */
int  mfn_vcf_load(const char *path, mfn_vcf **out);
int  mfn_vcf_cluster(mfn_vcf *v, uint32_t k, uint32_t comb, int nosplit);      /* prints the reference's per-contig summary to stderr */
int  mfn_vcf_info(const mfn_vcf *v, mfn_vcf_info_t *info);
/* header lines, each followed by '\n'; *need = bytes wanted (call with buf = NULL to size) */
int  mfn_vcf_header(const mfn_vcf *v, char *buf, uint64_t cap, uint64_t *need);
uint64_t mfn_vcf_clusters_of(const mfn_vcf *v, const char *ident);
void mfn_vcf_free(mfn_vcf *v);
/* *result (and *debug when debug != NULL: the -debug records, var_mer_id = running id) are
 * malloc'ed texts the caller releases with mfn_text_free. */
int  mfn_variants_process(mfn_ctx *ctx, const mfn_table *reads, const mfn_table *asm_, mfn_vcf *v,
                          const char *ident, const char *seq, uint64_t seq_len, int mode, uint32_t comb,
                          const mfn_kmodel *km, uint32_t shard, uint32_t nshards,
                          char **result, uint64_t *result_len, char **debug, uint64_t *debug_len, uint64_t *var_mer_id);
void mfn_text_free(char *p);

/* -completeness (merfin-completeness.C:48-144): every k-mer of the read DB as
 * stored (one strand, raw values, no -min/-max, Q4) joined with the raw
 * assembly DB value; result = joint histogram of (raw v_read, raw v_asm). */
int  mfn_completeness(mfn_ctx *ctx, const char *reads_dir, const char *asm_dir,
                      mfn_pair_count *pairs, uint64_t pairs_cap, uint64_t *npairs);
/* The same, one shard at a time for multi-GPU runs: data files [64 s/n, 64 (s+1)/n) of the read
 * database (the grain of the reference's OpenMP loop, merfin-completeness.C:65) are joined with an
 * assembly table loaded RAW (mfn_table_load(..., 0, UINT32_MAX)) into acc; the caller
 * sum-all-reduces the accumulator buffers (mfn_accum_device_buffer) and fetches the pairs. */
int  mfn_completeness_shard(mfn_ctx *ctx, const char *reads_dir, const mfn_table *asm_raw,
                            uint32_t shard, uint32_t nshards, mfn_accum *acc);

/* ---- host replay of the K* arithmetic (bit-exact doubles) ------------------ */

typedef struct {
  uint64_t  kasm, kmissing;
  double    koverCpy;
  double    qv_missing, qv_star;      /* histoQV, merfin-histogram.C:22-31 */
  uint32_t  n_undr, n_over;           /* bins filled below                 */
} mfn_hist_summary;

/* From the joint histogram: undr/over K* bins (merfin-histogram.C:71-90).
 * undr/over are caller arrays of cap entries; returns MFN_ERR_CAPACITY with
 * n_undr/n_over set to what is needed. */
int  mfn_hist_report(const mfn_pair_count *pairs, uint64_t npairs, const mfn_kmodel *km,
                     uint32_t k, uint64_t *undr, uint64_t *over, uint32_t cap,
                     mfn_hist_summary *summary);
/* Text of the -hist output file (merfin-histogram.C:151-163) into buf. */
int  mfn_hist_format(const uint64_t *undr, uint32_t n_undr, const uint64_t *over, uint32_t n_over,
                     char *buf, uint64_t buflen, uint64_t *needed);
void mfn_getK(const mfn_kmodel *km, uint32_t v_read, uint32_t v_asm,
              double *readK, double *asmK, double *prob);          /* merfin-globals.C:66-98   */
double mfn_getKmetric(double readK, double asmK);                  /* merfin-globals.H:248-261 */
/* -dump text of one record (merfin-dump.C:87-93, outputDump): "%s\t%lu\t%.2f\t%.2f\t%.2f\n" for
 * every valid position whose (readK, asmK, k*) are not all zero, written to file descriptor fd
 * in position order; fd < 0 only counts (what -skipMissing does, merfin-dump.C:34-38).  The
 * arrays are the record's slices of mfn_dump's output.  Formatting runs on `threads` host threads
 * (0 = all cores).  counts receives the record's kasm / kmissing (merfin-dump.C:58-66). */
int  mfn_dump_write(int fd, const char *ident, uint64_t n, const uint32_t *v_read, const uint32_t *v_asm,
                    const uint8_t *valid, const mfn_kmodel *km, int threads, mfn_seq_counts *counts);
/* The -completeness join over two RESIDENT tables (both loaded with min_value 0: raw values, Q4): every k-mer
 * the read table holds against the assembly table, into the same accumulator layout as mfn_completeness_shard.
 * Runs the join without a database on disk (the benchmark's tables are synthesised on the device). */
int  mfn_completeness_resident(mfn_ctx *ctx, const mfn_table *reads_raw, const mfn_table *asm_raw, mfn_accum *acc);
/* -completeness sums from the joint histogram of raw values. */
int  mfn_completeness_report(const mfn_pair_count *pairs, uint64_t npairs, const mfn_kmodel *km,
                             double *total, double *undr);

/* ---- host-side packing (no device needed) -------------------------------------- */
/* Pack global positions [g_begin, g_end) (multiples of 64) of a sequence set into the device
 * layout: 2 bits per base, (ascii >> 1) & 3 (kmers-tiny.H:69-83), and one validity bit per base
 * (ACGTacgt, kmers-iterator.H:97-102).  rec_start[r] = seq_off[r] - seq_off[0] + r (one separator
 * slot after every record), nseq + 1 entries.  seq2_out / ok_out receive (g_end - g_begin) / 32
 * resp. / 64 words.  Used by mfn_hist_host's host packing lane; thread-safe on disjoint ranges. */
void mfn_pack_host_range(const char *bases, const uint64_t *rec_start, uint32_t nseq, uint64_t g_begin,
                         uint64_t g_end, uint64_t *seq2_out, uint64_t *ok_out);
const char *mfn_pack_host_isa(void);       /* "avx2" or "scalar" ("avx512" on request: MFN_PACK_ISA) */

/* ---- choice of the -hist scan ------------------------------------------------------------
 * -hist over a big batch runs as the range-partitioned scan whenever a fused or
 * restricted table exists; its bins are sized for k-mers that spread over the table (15 % slack
 * and a side list of 1/16 of the windows).  A batch dominated by a handful of k-mers can overflow
 * that: the loss is recorded in the accumulator and mfn_accum_fetch / mfn_accum_decode then fail
 * with MFN_ERR_CAPACITY instead of returning a short count.  mfn_hist retries by itself; callers of
 * mfn_hist_host / mfn_hist_resident (which return before the device is done) recover with
 *     mfn_ctx_set_sweep(ctx, 0); mfn_accum_reset(ctx, acc); <the same call again>; mfn_ctx_set_sweep(ctx, -1);
 * mode: -1 automatic (default), 0 never (always the direct scan), 1 whenever a table allows it.
 * The MFN_SWEEP environment variable, when set, overrides the automatic choice the same way.
 * The ctx is NOT thread-safe (one host thread drives a ctx at a time) except where a function says
 * otherwise (mfn_dump, mfn_score_paths, mfn_variants_process). */
int  mfn_ctx_set_sweep(mfn_ctx *ctx, int mode);

/* ---- timing hooks for benchmarks -------------------------------------------- */
/* Number of kernels this library launched on the ctx since creation. */
uint64_t mfn_ctx_launch_count(const mfn_ctx *ctx);
/* Bytes of sequence data the host-buffer calls copied host -> device on this ctx since creation
 * (raw ASCII of the upload lane + packed words of the host packing lane). */
uint64_t mfn_ctx_h2d_bytes(const mfn_ctx *ctx);
/* Per-pass device times of the range-partitioned -hist scan.  After mfn_ctx_time_passes(ctx, 1)
 * every scan brackets its passes with CUDA events on the launch stream; mfn_ctx_pass_times
 * synchronises and returns the summed milliseconds of pass 1 (scatter) and pass 2 (probe) and
 * the number of scans measured since the last call. */
int  mfn_ctx_time_passes(mfn_ctx *ctx, int enable);
int  mfn_ctx_pass_times(mfn_ctx *ctx, double *scatter_ms, double *probe_ms, uint32_t *nscans);

#ifdef __cplusplus
}
#endif
#endif /* MERFIN_B200_H */
