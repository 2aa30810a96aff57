// GPU decode of on-disk meryl data files (reference: merylFileBlockReader::
// loadBlock / decodeBlock, kmers-files.C:24-128, stuffedBits::getUnary / getBinary, bits.C:371-410,
// 523-572).
//
// A data file is a sequence of blocks.  Each block is one dumped bit container (bits.C:106-235):
// u64 maxBits, u32 nSub, u32 blocksMax, u64 bgn[nSub], u64 len[nSub], then the words of every
// sub-block; bits are MSB-first inside little-endian 64-bit words and a coded value never
// straddles sub-blocks.  The bit stream is: header (magic x2, block prefix, nKmers, kCode,
// unaryBits, binaryBits, k1, cCode, c1, c2), then per k-mer unary(delta of the high part) +
// binary(low binaryBits bits), then all values as 32- or 64-bit binary.
//
// The host only walks the container framing and the block headers (a few words per block) and
// uploads the file image as it is on disk (~4.5 bytes per k-mer instead of 12 decoded).  The
// codes themselves are a serial chain inside a block, but a database has 64 x 2^b blocks:
// one THREAD per block decodes its chain front to back and writes (k-mer, value) pairs at the
// block's output offset; tens of thousands of independent chains keep the device busy.
#include "table_host.h"

#include <sys/stat.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <mutex>
#include <thread>

namespace {

constexpr uint64_t MAGIC_DATA_1 = 0x7461446c7972656dull;    // "merylDat"
constexpr uint64_t MAGIC_DATA_2 = 0x0a3030656c694661ull;    // "aFile00\n"
constexpr int MAX_SUB = 64;                                  // sub-blocks per block we accept (the writer uses 8..32)

struct GBlock {
  uint64_t sub0;          // index of the block's first sub-block in the sub-block table
  uint32_t nsub;
  uint32_t binaryBits, valueBits;
  uint32_t start_sub;     // cursor after the header
  uint64_t start_pos;
  uint64_t n;             // k-mers in the block
  uint64_t out;           // output offset
  uint64_t base;          // blockPrefix << suffixSize
};
struct GSub { uint64_t word; uint64_t len; };    // first word in the image, bits

// device-side forward bit cursor over a block's sub-blocks (same rules as the host BitCursor)
struct DCursor {
  const unsigned long long *img;
  const GSub *subs;
  uint32_t blk;
  uint64_t pos, len;
  const unsigned long long *w;
  __device__ __forceinline__ void enter(uint32_t b) { blk = b; pos = 0; len = subs[b].len; w = img + subs[b].word; }
  __device__ __forceinline__ void fit(uint64_t width) { if (pos + width > len) enter(blk + 1); }
  __device__ __forceinline__ uint64_t binary(uint32_t width) {
    if (width == 0) return 0;
    fit(width);
    const uint64_t wi = pos >> 6, bo = pos & 63;
    uint64_t v;
    if (bo + width <= 64) v = (__ldg(&w[wi]) << bo) >> (64 - width);
    else {
      const uint32_t w2 = (uint32_t)(bo + width - 64);
      v = (((__ldg(&w[wi]) << bo) >> bo) << w2) | (__ldg(&w[wi + 1]) >> (64 - w2));
    }
    pos += width;
    return v;
  }
  __device__ __forceinline__ uint64_t unary() {
    fit(1);
    uint64_t wi = pos >> 6;
    const uint64_t bo = pos & 63;
    uint64_t cur = __ldg(&w[wi]) << bo;
    uint64_t n = 0;
    if (cur == 0) {
      n = 64 - bo;
      wi++;
      while ((cur = __ldg(&w[wi])) == 0) { n += 64; wi++; }
    }
    n += (uint64_t)__clzll((long long)cur);
    pos += n + 1;
    return n;
  }
};

__global__ void __launch_bounds__(128) k_meryl_decode(const unsigned long long *__restrict__ img, const GBlock *__restrict__ blocks,
                                                       const GSub *__restrict__ subs, uint32_t nblocks,
                                                       unsigned long long *__restrict__ kmers, uint32_t *__restrict__ values) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblocks) return;
  const GBlock B = blocks[b];
  DCursor c;
  c.img = img; c.subs = subs + B.sub0;
  c.enter(B.start_sub);
  c.pos = B.start_pos;
  uint64_t high = 0;
  for (uint64_t i = 0; i < B.n; i++) {
    high += c.unary();
    const uint64_t low = c.binary(B.binaryBits);
    kmers[B.out + i] = B.base | (B.binaryBits >= 64 ? low : ((high << B.binaryBits) | low));
  }
  for (uint64_t i = 0; i < B.n; i++) values[B.out + i] = (uint32_t)c.binary(B.valueBits);
}

// host: walk one file image, append its blocks / sub-blocks to the tables.  word0 = index of the
// image's first word in the device buffer.
struct HostCursor {
  const uint64_t *img; const GSub *subs; uint32_t blk = 0; uint64_t pos = 0;
  void fit(uint64_t width) { if (pos + width > subs[blk].len) { blk++; pos = 0; } }
  uint64_t binary(uint32_t width) {
    fit(width);
    const uint64_t *w = img + subs[blk].word;
    const uint64_t wi = pos >> 6, bo = pos & 63;
    uint64_t v;
    if (bo + width <= 64) v = (w[wi] << bo) >> (64 - width);
    else { const uint32_t w2 = (uint32_t)(bo + width - 64); v = (((w[wi] << bo) >> bo) << w2) | (w[wi + 1] >> (64 - w2)); }
    pos += width;
    return v;
  }
};

int index_image(const MerylIndex &idx, const std::string &path, const uint64_t *img, size_t nwords, uint64_t word0,
                std::vector<GBlock> &blocks, std::vector<GSub> &subs, uint64_t &nout) {
  size_t off = 0;                                   // in words
  while (off + 2 <= nwords) {
    uint32_t nsub;
    memcpy(&nsub, (const char *)(img + off) + 8, 4);
    if (nsub == 0) break;
    if (nsub > MAX_SUB) { mfn_set_error("'%s': block with %u sub-blocks", path.c_str(), nsub); return MFN_ERR_FORMAT; }
    if (off + 2 + 2 * (size_t)nsub > nwords) { mfn_set_error("'%s': truncated block", path.c_str()); return MFN_ERR_FORMAT; }
    const uint64_t *lens = img + off + 2 + nsub;
    size_t w = off + 2 + 2 * (size_t)nsub;
    GBlock B{};
    B.sub0 = subs.size(); B.nsub = nsub;
    uint64_t total = 0;
    for (uint32_t i = 0; i < nsub; i++) {
      const size_t nw = (lens[i] + 63) / 64;
      if (w + nw > nwords) { mfn_set_error("'%s': truncated block", path.c_str()); return MFN_ERR_FORMAT; }
      subs.push_back({w, lens[i]});               // relative to this image for the header walk below
      total += lens[i];
      w += nw;
    }
    off = w;
    if (total == 0) { subs.resize(B.sub0); break; }
    HostCursor c{img, subs.data() + B.sub0};
    if (c.binary(64) != MAGIC_DATA_1 || c.binary(64) != MAGIC_DATA_2) { mfn_set_error("'%s': block magic number mismatch", path.c_str()); return MFN_ERR_FORMAT; }
    const uint64_t blockPrefix = (uint32_t)c.binary(64);
    B.n = c.binary(64);
    const uint32_t kCode = (uint32_t)c.binary(8);
    (void)c.binary(32);
    B.binaryBits = (uint32_t)c.binary(32);
    (void)c.binary(64);
    const uint32_t cCode = (uint32_t)c.binary(8);
    (void)c.binary(64); (void)c.binary(64);
    if (kCode != 1 || (cCode != 1 && cCode != 2)) { mfn_set_error("'%s': unknown block coding %u/%u", path.c_str(), kCode, cCode); return MFN_ERR_FORMAT; }
    if (B.binaryBits > 64 || idx.suffixSize > 64) { mfn_set_error("'%s': k-mer suffix wider than 64 bits", path.c_str()); return MFN_ERR_UNSUPPORTED; }
    B.valueBits = cCode == 1 ? 32 : 64;
    B.start_sub = c.blk; B.start_pos = c.pos;
    B.base = idx.suffixSize >= 64 ? 0 : (blockPrefix << idx.suffixSize);
    B.out = nout;
    nout += B.n;
    for (uint32_t i = 0; i < nsub; i++) subs[B.sub0 + i].word += word0;     // now relative to the device buffer
    blocks.push_back(B);
  }
  return MFN_OK;
}

// file size in 64-bit words, plus two words of slack (the cursors may peek one word ahead)
bool file_words(const std::string &path, size_t &nwords) {
  struct stat st;
  if (stat(path.c_str(), &st) != 0) return false;
  nwords = ((size_t)st.st_size + 7) / 8 + 2;
  return true;
}

bool slurp_into(const std::string &path, uint64_t *dst, size_t nwords) {
  FILE *F = fopen(path.c_str(), "rb");
  if (!F) return false;
  memset(dst + (nwords >= 3 ? nwords - 3 : 0), 0, (nwords >= 3 ? 3 : nwords) * 8);
  const size_t want = (nwords - 2) * 8;
  size_t got = 0;
  while (got < want) {
    const size_t r = fread((char *)dst + got, 1, want - got, F);
    if (r == 0) break;
    got += r;
  }
  fclose(F);
  return got + 8 > want;                            // the size was rounded up to whole words
}

// One group of whole files staged in pinned memory.
struct Group {
  uint32_t f0 = 0, f1 = 0;
  std::vector<uint64_t> word0;                      // first word of each file in the staging buffer
  size_t words = 0;
  int status = MFN_OK;
  std::string err;
};

}  // namespace

// Same contract as meryl_stream_to_device (meryl_db.cpp), with the decode on the device.
// Files are read straight into one of two pinned staging buffers by a few reader threads while
// the device works on the other one: read, upload, decode and the consumer overlap.
int meryl_stream_to_device_gpu(mfn_ctx *ctx, const MerylIndex &idx,
                               const std::function<int(const unsigned long long *, const uint32_t *, uint64_t)> &fn,
                               uint32_t file_begin, uint32_t file_end) {
  if (idx.k > 32) { mfn_set_error("k=%u: this build supports k <= 32", idx.k); return MFN_ERR_UNSUPPORTED; }
  MFN_CUDA(cudaSetDevice(ctx->device));
  const uint32_t nfiles = std::min(file_end, 1u << idx.numFilesBits);      // data files [file_begin, nfiles)
  const unsigned nthreads = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  const bool dbg = getenv("MFN_LOAD_DEBUG") != nullptr;
  auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_begin = now();
  double t_wait = 0, t_index = 0, t_dev = 0;
  uint64_t n_total = 0, bytes_total = 0;

  // groups of whole files of at most ~1 GiB (a single bigger file is its own group)
  std::vector<size_t> fwords(nfiles, 0);
  for (uint32_t f = file_begin; f < nfiles; f++)
    if (!file_words(meryl_file_name(idx, f, false), fwords[f])) { mfn_set_error("cannot read '%s'", meryl_file_name(idx, f, false).c_str()); return MFN_ERR_IO; }
  const size_t GROUP_WORDS = (size_t)1 << 23;               // 64 MiB of file image per group (two pinned buffers of the group size: pinning costs ~0.4 s per GB, which a small database would not earn back)
  std::vector<Group> groups;
  for (uint32_t f = file_begin; f < nfiles;) {
    Group g; g.f0 = f;
    while (f < nfiles && (f == g.f0 || g.words + fwords[f] <= GROUP_WORDS)) { g.word0.push_back(g.words); g.words += fwords[f]; f++; }
    g.f1 = f;
    groups.push_back(std::move(g));
  }
  size_t max_words = 0;
  for (auto &g : groups) max_words = std::max(max_words, g.words);
  // two pinned staging buffers, kept in the ctx between calls (pinning memory costs ~0.4 s per GB)
  if (ctx->pinned_cap < 2 * max_words * 8) {
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    ctx->h_pinned = nullptr; ctx->pinned_cap = 0;
    if (cudaMallocHost(&ctx->h_pinned, 2 * max_words * 8) != cudaSuccess) { mfn_set_error("cudaMallocHost(meryl staging, %zu bytes)", 2 * max_words * 8); return MFN_ERR_NOMEM; }
    ctx->pinned_cap = 2 * max_words * 8;
  }
  uint64_t *h_stage[2] = {(uint64_t *)ctx->h_pinned, (uint64_t *)ctx->h_pinned + max_words};
  auto read_group = [&](Group &g, uint64_t *dst) {
    std::atomic<uint32_t> next{g.f0};
    std::vector<std::thread> th;
    std::mutex mu;
    for (unsigned t = 0; t < std::min<unsigned>(nthreads, g.f1 - g.f0); t++)
      th.emplace_back([&]() {
        for (uint32_t f; (f = next++) < g.f1;)
          if (!slurp_into(meryl_file_name(idx, f, false), dst + g.word0[f - g.f0], fwords[f])) {
            std::lock_guard<std::mutex> lk(mu);
            g.status = MFN_ERR_IO; g.err = "cannot read '" + meryl_file_name(idx, f, false) + "'";
          }
      });
    for (auto &t : th) t.join();
  };

  int status = MFN_OK;
  unsigned long long *d_img = nullptr, *d_k = nullptr; uint32_t *d_v = nullptr; GBlock *d_b = nullptr; GSub *d_s = nullptr;
  size_t cap_img = 0, cap_k = 0, cap_v = 0, cap_b = 0, cap_s = 0;
  auto grow = [&](void **p, size_t &cap, size_t need) -> int {
    if (need <= cap) return MFN_OK;
    if (*p) cudaFree(*p);
    *p = nullptr; cap = 0;
    need += need / 4;
    if (cudaMalloc(p, need) != cudaSuccess) { mfn_set_error("cudaMalloc(meryl decode, %zu bytes)", need); return MFN_ERR_NOMEM; }
    cap = need;
    return MFN_OK;
  };
  std::thread reader;
  if (!groups.empty()) reader = std::thread(read_group, std::ref(groups[0]), h_stage[0]);
  for (size_t gi = 0; gi < groups.size() && status == MFN_OK; gi++) {
    Group &g = groups[gi];
    uint64_t *img = h_stage[gi & 1];
    double t0 = now();
    reader.join();                                            // group gi is in its staging buffer
    t_wait += now() - t0;
    // the other buffer is free (its upload was synchronised last round): read ahead
    if (gi + 1 < groups.size()) reader = std::thread(read_group, std::ref(groups[gi + 1]), h_stage[(gi + 1) & 1]);
    if (g.status != MFN_OK) { status = g.status; mfn_set_error("%s", g.err.c_str()); break; }
    t0 = now();
    std::vector<GBlock> blocks; std::vector<GSub> subs;
    uint64_t nout = 0;
    for (uint32_t f = g.f0; f < g.f1 && status == MFN_OK; f++)
      status = index_image(idx, meryl_file_name(idx, f, false), img + g.word0[f - g.f0], fwords[f] - 2, g.word0[f - g.f0], blocks, subs, nout);
    t_index += now() - t0;
    if (status != MFN_OK || nout == 0) continue;
    t0 = now();
    n_total += nout; bytes_total += g.words * 8;
    status = grow((void **)&d_img, cap_img, g.words * 8);
    if (status == MFN_OK) status = grow((void **)&d_k, cap_k, nout * 8);
    if (status == MFN_OK) status = grow((void **)&d_v, cap_v, nout * 4);
    if (status == MFN_OK) status = grow((void **)&d_b, cap_b, blocks.size() * sizeof(GBlock));
    if (status == MFN_OK) status = grow((void **)&d_s, cap_s, subs.size() * sizeof(GSub));
    if (status != MFN_OK) break;
    MFN_CUDA(cudaMemcpyAsync(d_img, img, g.words * 8, cudaMemcpyHostToDevice, ctx->stream));
    MFN_CUDA(cudaMemcpyAsync(d_b, blocks.data(), blocks.size() * sizeof(GBlock), cudaMemcpyHostToDevice, ctx->stream));
    MFN_CUDA(cudaMemcpyAsync(d_s, subs.data(), subs.size() * sizeof(GSub), cudaMemcpyHostToDevice, ctx->stream));
    k_meryl_decode<<<(unsigned)ceil_div_u64(blocks.size(), 128), 128, 0, ctx->stream>>>(d_img, d_b, d_s, (uint32_t)blocks.size(), d_k, d_v);
    ctx->launches++;
    MFN_CUDA(cudaGetLastError());
    status = fn(d_k, d_v, nout);
    MFN_CUDA(cudaStreamSynchronize(ctx->stream));          // staging, descriptor vectors and device buffers are reused
    t_dev += now() - t0;
  }
  if (reader.joinable()) reader.join();
  cudaFree(d_img); cudaFree(d_k); cudaFree(d_v); cudaFree(d_b); cudaFree(d_s);
  if (dbg)
    fprintf(stderr, "[meryl load] %s: %llu k-mers, %.2f GB on disk, %zu groups; %.3f s total: waiting for reads %.3f s, index %.3f s, upload+decode+consume %.3f s\n",
            idx.dir.c_str(), (unsigned long long)n_total, bytes_total / 1e9, groups.size(), now() - t_begin, t_wait, t_index, t_dev);
  return status;
}
