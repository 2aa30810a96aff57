// Internal declarations shared by the host and device parts of libmerfin_b200.
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/merfin_b200.h"

void mfn_set_error(const char *fmt, ...);

#define MFN_CUDA(call)                                                                       \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess) {                                                                 \
      mfn_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));   \
      return MFN_ERR_CUDA;                                                                   \
    }                                                                                        \
  } while (0)

// entry points that only exist for 64-bit k-mers
#define MFN_REJECT_WIDE(t, what)                                                                        \
  do {                                                                                                  \
    if ((t) && (t)->wide) {                                                                             \
      mfn_set_error("%s: not available for k > 32 (tables of 128-bit k-mers)", what);                    \
      return MFN_ERR_UNSUPPORTED;                                                                       \
    }                                                                                                   \
  } while (0)

#define MFN_TRY(call)                 \
  do {                                \
    int s_ = (call);                  \
    if (s_ != MFN_OK) return s_;      \
  } while (0)

// ---------------------------------------------------------------------------
// Device-side table description (passed to kernels by value).
//
// Layout in HBM (DESIGN.md "Table layout"): two levels of 32-byte buckets, one
// bucket = one DRAM sector.  A k-mer c is mapped by a bijective mix of its 2k
// bits to h; bucket = h / D, tag = h % D, so (bucket, tag) identifies c exactly
// and only tag + value are stored: slot = tag << vbits | value, value != 0,
// slot == 0 means empty.  A bucket whose slots are all occupied may have spilled
// entries to the next level; level 3 is a small open-addressing side table with
// full keys.  Values >= vmax saturate in the slot and keep their excess in a
// second side table.
// ---------------------------------------------------------------------------

struct OaTable {          // open addressing, 64-bit keys, EMPTY = ~0
  unsigned long long *keys;
  uint32_t           *vals;
  uint32_t            mask;      // capacity - 1
  uint32_t           *special;   // value for key == ~0 (only reachable when k == 32)
};

struct TableLevel {
  void    *slots;         // buckets * 32 bytes
  uint64_t buckets;
  uint64_t D;             // divisor: bucket = h / D
  uint64_t magic;         // floor(2^64 / D) for the division
  uint64_t seed;
};

struct TableStats {       // updated by the build kernels
  unsigned long long n_l1, n_l2, n_stash, n_big, n_dup, n_noncanon, errors;
};

struct DevTable {
  TableLevel lv[2];
  OaTable    stash, big;
  uint64_t   kmask;       // low 2k bits
  uint32_t   k;
  uint32_t   w64;         // slots are 64-bit (4 per bucket) instead of 32-bit (8 per bucket)
  uint32_t   vbits;       // value field width
  uint32_t   vmax;        // saturated value (== value field mask, capped to 32 bits)
  uint32_t   probe_both;  // table is not canonical: look up f and r (Q2), else min(f, r)
  TableStats *stats;
};

#define MFN_STAGE_BUFS 4              // ring of 64 MB ASCII staging buffers for host-buffer calls

struct mfn_ctx {
  int          device = 0;
  cudaStream_t stream = nullptr;       // stream in use
  cudaStream_t own_stream = nullptr;
  cudaStream_t copy_stream = nullptr;    // H2D of the ASCII staging buffers
  cudaStream_t pack_stream = nullptr;    // k_pack: the only writer of a sequence set's packed arrays
  cudaStream_t probe_stream = nullptr;   // probe pass of a finished sub-sweep next to the scatter passes of the following one
  cudaEvent_t  ev_probe[2] = {};
  int          side_probe_ctas = 1;      // CTAs per SM of a probe pass on the second stream
  cudaEvent_t  ev[MFN_STAGE_BUFS + 3] = {};   // [b]: pack kernel that last read staging buffer b; then copy / pack / order
  int          sm_count = 0;
  std::atomic<uint64_t> launches{0};
  std::atomic<uint64_t> h2d_bytes{0};    // bytes of sequence data copied host -> device by host-buffer calls
  // staging for host-buffer calls
  char  *d_ascii[MFN_STAGE_BUFS] = {};
  size_t ascii_cap = 0;
  int    stage_next = 0;
  void  *h_pinned = nullptr;
  size_t pinned_cap = 0;
  // optional per-pass timing of the sweep (mfn_ctx_time_passes)
  int  sweep_mode = -1;                  // mfn_ctx_set_sweep: -1 automatic, 0 never, 1 whenever possible
  bool time_passes = false;
  std::vector<cudaEvent_t> pass_ev;      // triples: before scatter, between, after probe
  size_t pass_ev_used = 0;
  struct mfn_ingest *ingest = nullptr;   // host packing lane of the host-buffer -hist path (ingest.cu)
  std::mutex dump_mu;                    // mfn_dump / mfn_score_paths: one device section at a time per ctx
  struct mfn_seq   *scratch_seq = nullptr;
  struct mfn_accum *scratch_acc = nullptr;
};

struct mfn_joint;
struct mfn_wide;                       // k > 32: sorted 128-bit k-mers (wide.cu)

struct mfn_table {
  mfn_ctx   *ctx = nullptr;
  int        device = 0;          // (kept apart from ctx: the object may be freed after its ctx)
  mfn_joint *joint = nullptr;     // fused (this, assembly) table built by mfn_tables_fuse
  mfn_joint *hot = nullptr;       // fused table restricted to the k-mers of one sequence shard (mfn_tables_restrict)
  mfn_wide  *wide = nullptr;      // k > 32: the table lives here and `dev` only carries k
  bool       streamed = false;    // no per-database table behind this handle: the database was streamed into `hot` (joint_stream.cu)
  DevTable   dev{};
  TableStats host_stats{};
  uint32_t   min_value = 0, max_value = 0, value_offset = 0;
  uint32_t   tag_bits = 0;
  uint64_t   device_bytes = 0;
  uint64_t   n_kmers = 0;
  std::vector<void *> allocs;
};

struct mfn_seq {
  mfn_ctx  *ctx = nullptr;
  int       device = 0;
  uint32_t  nseq = 0;
  uint64_t  total_bases = 0;      // caller's concatenated length
  uint64_t  G = 0;                // global coordinates incl. one separator per record
  uint64_t  ntiles = 0;
  unsigned long long *d_seq2 = nullptr;   // 2 bits per base, base g at bits 2*(g%32) of word g/32
  unsigned long long *d_ok = nullptr;     // 1 bit per base: set iff the byte is one of ACGTacgt
  unsigned long long *d_rec_start = nullptr;   // [nseq+1] global coordinate of each record start
  std::vector<uint64_t> rec_start;             // host copy
  uint64_t  device_bytes = 0;
  bool      pooled = false;       // device arrays come from the stream-ordered pool (mfn_seq_create_in)
};

// Dense tier of the joint histogram: v_read < MFN_DENSE_R and v_asm < MFN_DENSE_A.
#define MFN_DENSE_R 4096u
#define MFN_DENSE_A 64u
#define MFN_ACC_HEADER 8u           // words: [0] valid windows, [1] sparse pairs inserted, [2] errors
#define MFN_SPARSE_CAP (1u << 22)

struct PairTable {                  // (v_read << 32 | v_asm) -> count, EMPTY = ~0
  unsigned long long *keys;
  unsigned long long *counts;
  uint32_t            mask;
  unsigned long long *local_n;      // pairs inserted on THIS rank (header word [1] becomes the sum over ranks after an all-reduce)
};

struct mfn_accum {
  mfn_ctx  *ctx = nullptr;
  int       device = 0;
  uint32_t  nseq = 0;
  uint64_t  nwords = 0;             // header + dense + 2*nseq
  unsigned long long *d_buf = nullptr;
  unsigned long long *h_buf = nullptr;   // pinned mirror of d_buf for fetches
  PairTable sparse{};
};

// tile geometry of the scan kernels
#define MFN_TILE 2048               // window start positions per tile (multiple of 64)

static inline __host__ __device__ uint64_t ceil_div_u64(uint64_t a, uint64_t b) { return (a + b - 1) / b; }
