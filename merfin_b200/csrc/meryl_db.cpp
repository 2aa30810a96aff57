// meryl on-disk database: reader (tables are loaded from it) and writer (synthetic
// generators emit it so the reference can load the same data).
//
// Format facts restated from the reference (file:line relative to the reference tree):
//   master index   kmers-writer.C:117-151, kmers-reader.C:60-177, stats kmers-histogram.C:69-102,180-207
//   data files     names kmers.C:29-55; block header + codes kmers-writer.C:181-288, kmers-files.C:24-128
//   bit container  bits.C:106-235 (framing), bits.H:665-734 (values never straddle sub-blocks),
//                  bits.C:371-410 (unary), bits.C:523-572 (binary); bits are MSB-first in u64 words
#include "table_host.h"

#include <algorithm>
#include <mutex>
#include <atomic>
#include <map>
#include <thread>

#include <sys/stat.h>

namespace {

constexpr uint64_t MAGIC_INDEX_1 = 0x646e496c7972656dull;   // "merylInd"
constexpr uint64_t MAGIC_INDEX_V1 = 0x31302e765f5f7865ull;  // "ex__v.01"
constexpr uint64_t MAGIC_INDEX_V2 = 0x32302e765f5f7865ull;
constexpr uint64_t MAGIC_INDEX_V3 = 0x33302e765f5f7865ull;
constexpr uint64_t MAGIC_DATA_1 = 0x7461446c7972656dull;    // "merylDat"
constexpr uint64_t MAGIC_DATA_2 = 0x0a3030656c694661ull;    // "aFile00\n"

// One dumped bit container parsed in place from a memory image of the file.
struct BitBlocks {
  std::vector<uint64_t> len;                 // bits per sub-block
  std::vector<const uint64_t *> dat;         // words of each sub-block (inside the file image)
  size_t bytes = 0;                          // bytes consumed from the image

  bool parse(const uint8_t *p, size_t avail) {
    if (avail < 16) return false;
    uint32_t nsub;
    memcpy(&nsub, p + 8, 4);
    size_t off = 16;
    if (avail < off + (size_t)nsub * 16) return false;
    len.resize(nsub); dat.resize(nsub);
    const uint8_t *lens = p + off + (size_t)nsub * 8;
    off += (size_t)nsub * 16;
    for (uint32_t i = 0; i < nsub; i++) {
      memcpy(&len[i], lens + (size_t)i * 8, 8);
      size_t nw = (len[i] + 63) / 64;
      if (avail < off + nw * 8) return false;
      dat[i] = (const uint64_t *)(p + off);
      off += nw * 8;
    }
    bytes = off;
    return true;
  }
  uint64_t total_bits() const { uint64_t t = 0; for (uint64_t l : len) { if (!l) break; t += l; } return t; }
};

// Forward bit cursor over the sub-blocks.
struct BitCursor {
  const BitBlocks &b;
  uint32_t blk = 0;
  uint64_t pos = 0;            // bit position in the current sub-block
  const uint64_t *w = nullptr;
  explicit BitCursor(const BitBlocks &bb) : b(bb) { w = b.dat.empty() ? nullptr : b.dat[0]; }

  inline void fit(uint64_t width) {          // a value that does not fit starts the next sub-block
    if (pos + width > b.len[blk]) { blk++; pos = 0; w = b.dat[blk]; }
  }
  inline uint64_t binary(uint32_t width) {
    if (width == 0) return 0;
    fit(width);
    uint64_t wi = pos >> 6, bo = pos & 63;   // bo bits of word wi are already consumed (from the top)
    uint64_t v;
    if (bo + width <= 64) {
      v = (w[wi] << bo) >> (64 - width);
    } else {
      uint32_t w2 = (uint32_t)(bo + width - 64);
      v = (((w[wi] << bo) >> bo) << w2) | (w[wi + 1] >> (64 - w2));
    }
    pos += width;
    return v;
  }
  inline uint64_t unary() {                   // number of 0 bits before the next 1 bit
    fit(1);
    uint64_t wi = pos >> 6, bo = pos & 63;
    uint64_t cur = w[wi] << bo;
    uint64_t n = 0;
    if (cur == 0) {
      n = 64 - bo;
      wi++;
      while ((cur = w[wi]) == 0) { n += 64; wi++; }
    }
    n += (uint64_t)__builtin_clzll(cur);
    pos += n + 1;
    return n;
  }
};

bool read_file(const std::string &path, std::vector<uint8_t> &buf) {
  FILE *F = fopen(path.c_str(), "rb");
  if (!F) return false;
  struct stat st;
  if (fstat(fileno(F), &st) != 0) { fclose(F); return false; }
  buf.resize((size_t)st.st_size + 16);          // slack so word reads at the end stay in bounds
  size_t got = st.st_size ? fread(buf.data(), 1, (size_t)st.st_size, F) : 0;
  fclose(F);
  if (got != (size_t)st.st_size) return false;
  memset(buf.data() + got, 0, 16);
  buf.resize(got + 16);
  return true;
}

}  // namespace

std::string meryl_file_name(const MerylIndex &idx, uint32_t file, bool index) {
  std::string s = idx.dir + "/0x";
  for (int b = (int)idx.numFilesBits - 1; b >= 0; b--) s += ((file >> b) & 1) ? '1' : '0';
  s += index ? ".merylIndex" : ".merylData";
  return s;
}

int meryl_open(const char *dir, MerylIndex &idx) {
  idx = MerylIndex();
  idx.dir = dir;
  std::vector<uint8_t> buf;
  std::string path = idx.dir + "/merylIndex";
  if (!read_file(path, buf)) {
    mfn_set_error("'%s' doesn't appear to be a meryl database; cannot read '%s'", dir, path.c_str());
    return MFN_ERR_IO;
  }
  BitBlocks bb;
  if (!bb.parse(buf.data(), buf.size() - 16) || bb.total_bits() < 64 * 2 + 32 * 4) {
    mfn_set_error("'%s': truncated master index", path.c_str());
    return MFN_ERR_FORMAT;
  }
  BitCursor c(bb);
  uint64_t m1 = c.binary(64), m2 = c.binary(64);
  if (m1 != MAGIC_INDEX_1 || (m2 != MAGIC_INDEX_V1 && m2 != MAGIC_INDEX_V2 && m2 != MAGIC_INDEX_V3)) {
    mfn_set_error("'%s' fails the meryl magic number check", path.c_str());
    return MFN_ERR_FORMAT;
  }
  idx.version = (m2 == MAGIC_INDEX_V1) ? 1 : (m2 == MAGIC_INDEX_V2) ? 2 : 3;
  idx.prefixSize = (uint32_t)c.binary(32);
  idx.suffixSize = (uint32_t)c.binary(32);
  idx.numFilesBits = (uint32_t)c.binary(32);
  idx.numBlocksBits = (uint32_t)c.binary(32);
  idx.flags = idx.version >= 2 ? (uint32_t)c.binary(32) : 0;
  idx.k = (idx.prefixSize + idx.suffixSize) / 2;
  if (idx.version == 3) {
    idx.numUnique = c.binary(64);
    idx.numDistinct = c.binary(64);
    idx.numTotal = c.binary(64);
    uint64_t hl = c.binary(64);
    if (hl > (1ull << 32)) { mfn_set_error("'%s': implausible histogram length", path.c_str()); return MFN_ERR_FORMAT; }
    idx.histValue.resize(hl); idx.histCount.resize(hl);
    for (uint64_t i = 0; i < hl; i++) { idx.histValue[i] = c.binary(64); idx.histCount[i] = c.binary(64); }
  }
  if (idx.numFilesBits > 12) { mfn_set_error("'%s': implausible file count", path.c_str()); return MFN_ERR_FORMAT; }
  return MFN_OK;
}

int meryl_decode_file(const MerylIndex &idx, uint32_t file,
                      const std::function<void(const uint64_t *, const uint32_t *, uint64_t)> &sink) {
  std::vector<uint8_t> buf;
  std::string path = meryl_file_name(idx, file, false);
  if (!read_file(path, buf)) { mfn_set_error("cannot read '%s'", path.c_str()); return MFN_ERR_IO; }
  size_t size = buf.size() - 16, off = 0;
  std::vector<uint64_t> kmers;
  std::vector<uint32_t> values;
  while (off < size) {
    BitBlocks bb;
    if (!bb.parse(buf.data() + off, size - off)) { mfn_set_error("'%s': truncated block", path.c_str()); return MFN_ERR_FORMAT; }
    off += bb.bytes;
    if (bb.total_bits() == 0) break;
    BitCursor c(bb);
    if (c.binary(64) != MAGIC_DATA_1 || c.binary(64) != MAGIC_DATA_2) {
      mfn_set_error("'%s': block magic number mismatch", path.c_str());
      return MFN_ERR_FORMAT;
    }
    uint64_t blockPrefix = (uint32_t)c.binary(64);
    uint64_t n = c.binary(64);
    uint32_t kCode = (uint32_t)c.binary(8);
    (void)c.binary(32);                                   // unaryBits (informational)
    uint32_t binaryBits = (uint32_t)c.binary(32);
    (void)c.binary(64);
    uint32_t cCode = (uint32_t)c.binary(8);
    (void)c.binary(64); (void)c.binary(64);
    if (kCode != 1 || (cCode != 1 && cCode != 2)) { mfn_set_error("'%s': unknown block coding %u/%u", path.c_str(), kCode, cCode); return MFN_ERR_FORMAT; }
    if (binaryBits > 64 || idx.suffixSize > 64) { mfn_set_error("'%s': k-mer suffix wider than 64 bits", path.c_str()); return MFN_ERR_UNSUPPORTED; }
    kmers.resize(n); values.resize(n);
    uint64_t high = 0;
    const uint64_t base = idx.suffixSize >= 64 ? 0 : (blockPrefix << idx.suffixSize);
    for (uint64_t i = 0; i < n; i++) {
      high += c.unary();
      uint64_t lowbits = c.binary(binaryBits);
      kmers[i] = base | (binaryBits >= 64 ? lowbits : ((high << binaryBits) | lowbits));
    }
    const uint32_t vw = cCode == 1 ? 32 : 64;
    for (uint64_t i = 0; i < n; i++) values[i] = (uint32_t)c.binary(vw);
    sink(kmers.data(), values.data(), n);
  }
  return MFN_OK;
}

// The same for any k <= 64: k-mers as (hi, lo) halves of the 128-bit kmdata (kmers-tiny.H:29).  The binary
// part of a suffix wider than 64 bits is stored as two values, its top binaryBits - 64 bits and then its low
// 64 bits (kmers-files.C:88-107); each is framed on its own (a value never straddles sub-blocks).
int meryl_decode_file_wide(const MerylIndex &idx, uint32_t file,
                           const std::function<void(const uint64_t *, const uint64_t *, const uint32_t *, uint64_t)> &sink) {
  typedef unsigned __int128 u128;
  std::vector<uint8_t> buf;
  std::string path = meryl_file_name(idx, file, false);
  if (!read_file(path, buf)) { mfn_set_error("cannot read '%s'", path.c_str()); return MFN_ERR_IO; }
  size_t size = buf.size() - 16, off = 0;
  std::vector<uint64_t> hi, lo;
  std::vector<uint32_t> values;
  if (idx.suffixSize >= 128 || idx.prefixSize + idx.suffixSize > 128) { mfn_set_error("'%s': k-mers wider than 128 bits", path.c_str()); return MFN_ERR_UNSUPPORTED; }
  while (off < size) {
    BitBlocks bb;
    if (!bb.parse(buf.data() + off, size - off)) { mfn_set_error("'%s': truncated block", path.c_str()); return MFN_ERR_FORMAT; }
    off += bb.bytes;
    if (bb.total_bits() == 0) break;
    BitCursor c(bb);
    if (c.binary(64) != MAGIC_DATA_1 || c.binary(64) != MAGIC_DATA_2) {
      mfn_set_error("'%s': block magic number mismatch", path.c_str());
      return MFN_ERR_FORMAT;
    }
    const uint64_t blockPrefix = (uint32_t)c.binary(64);
    const uint64_t n = c.binary(64);
    const uint32_t kCode = (uint32_t)c.binary(8);
    (void)c.binary(32);                                   // unaryBits (informational)
    const uint32_t binaryBits = (uint32_t)c.binary(32);
    (void)c.binary(64);
    const uint32_t cCode = (uint32_t)c.binary(8);
    (void)c.binary(64); (void)c.binary(64);
    if (kCode != 1 || (cCode != 1 && cCode != 2)) { mfn_set_error("'%s': unknown block coding %u/%u", path.c_str(), kCode, cCode); return MFN_ERR_FORMAT; }
    if (binaryBits > 128) { mfn_set_error("'%s': k-mer suffix wider than 128 bits", path.c_str()); return MFN_ERR_FORMAT; }
    hi.resize(n); lo.resize(n); values.resize(n);
    const uint32_t ls = binaryBits <= 64 ? 0 : binaryBits - 64, rs = binaryBits <= 64 ? binaryBits : 64;
    const u128 base = (u128)blockPrefix << idx.suffixSize;
    u128 high = 0;
    for (uint64_t i = 0; i < n; i++) {
      high += c.unary();
      u128 suf = high;
      if (ls) suf = (suf << ls) | c.binary(ls);
      if (rs == 64) suf = ((suf << 32) << 32) | c.binary(64);
      else if (rs) suf = (suf << rs) | c.binary(rs);
      const u128 km = base | suf;
      lo[i] = (uint64_t)km; hi[i] = (uint64_t)(km >> 64);
    }
    const uint32_t vw = cCode == 1 ? 32 : 64;
    for (uint64_t i = 0; i < n; i++) values[i] = (uint32_t)c.binary(vw);
    sink(hi.data(), lo.data(), values.data(), n);
  }
  return MFN_OK;
}

// Every (k-mer, value) of a database in stored order (ascending k-mer), decoded on the host.
extern "C" int mfn_meryl_read(const char *dir, uint64_t *kmers_lo, uint64_t *kmers_hi, uint32_t *values, uint64_t cap, uint64_t *n) {
  if (!dir || !n) return MFN_ERR_INVALID;
  MerylIndex idx;
  MFN_TRY(meryl_open(dir, idx));
  if (idx.k > 64) { mfn_set_error("'%s': k=%u; k-mers hold at most 64 bases", dir, idx.k); return MFN_ERR_UNSUPPORTED; }
  if (idx.k > 32 && kmers_lo && !kmers_hi) { mfn_set_error("'%s': k=%u needs kmers_hi", dir, idx.k); return MFN_ERR_INVALID; }
  uint64_t fill = 0;
  const uint32_t nfiles = 1u << idx.numFilesBits;
  for (uint32_t f = 0; f < nfiles; f++)
    MFN_TRY(meryl_decode_file_wide(idx, f, [&](const uint64_t *h, const uint64_t *l, const uint32_t *v, uint64_t m) {
      for (uint64_t i = 0; i < m; i++, fill++) {
        if (fill >= cap) continue;
        if (kmers_lo) kmers_lo[fill] = l[i];
        if (kmers_hi) kmers_hi[fill] = h[i];
        if (values) values[fill] = v[i];
      }
    }));
  *n = fill;
  return fill > cap && (kmers_lo || kmers_hi || values) ? MFN_ERR_CAPACITY : MFN_OK;
}

int meryl_stream_to_device(mfn_ctx *ctx, const MerylIndex &idx,
                           const std::function<int(const unsigned long long *, const uint32_t *, uint64_t)> &fn,
                           uint32_t file_begin, uint32_t file_end) {
  if (idx.k > 32) { mfn_set_error("k=%u: this build supports k <= 32", idx.k); return MFN_ERR_UNSUPPORTED; }
  const uint32_t nfiles = std::min(file_end, 1u << idx.numFilesBits);      // data files [file_begin, nfiles)
  const unsigned nthreads = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
  const uint64_t CH = 1ull << 22;                 // entries per device chunk
  unsigned long long *dk = nullptr; uint32_t *dv = nullptr;
  uint64_t *hk = nullptr; uint32_t *hv = nullptr;
  MFN_CUDA(cudaSetDevice(ctx->device));
  MFN_CUDA(cudaMalloc(&dk, CH * 8));
  MFN_CUDA(cudaMalloc(&dv, CH * 4));
  MFN_CUDA(cudaMallocHost(&hk, CH * 8));
  MFN_CUDA(cudaMallocHost(&hv, CH * 4));
  int status = MFN_OK;
  uint64_t fill = 0;
  auto flush = [&]() -> int {
    if (fill == 0) return MFN_OK;
    MFN_CUDA(cudaMemcpyAsync(dk, hk, fill * 8, cudaMemcpyHostToDevice, ctx->stream));
    MFN_CUDA(cudaMemcpyAsync(dv, hv, fill * 4, cudaMemcpyHostToDevice, ctx->stream));
    int s = fn(dk, dv, fill);
    MFN_CUDA(cudaStreamSynchronize(ctx->stream));
    fill = 0;
    return s;
  };
  // decode a group of files in parallel, then feed them to the device in file order
  for (uint32_t f0 = file_begin; f0 < nfiles && status == MFN_OK; f0 += nthreads) {
    uint32_t f1 = std::min(nfiles, f0 + nthreads);
    std::vector<std::vector<uint64_t>> K(f1 - f0);
    std::vector<std::vector<uint32_t>> V(f1 - f0);
    std::vector<int> st(f1 - f0, MFN_OK);
    std::vector<std::string> errs(f1 - f0);
    std::vector<std::thread> th;
    for (uint32_t f = f0; f < f1; f++)
      th.emplace_back([&, f]() {
        auto &k = K[f - f0]; auto &v = V[f - f0];
        st[f - f0] = meryl_decode_file(idx, f, [&](const uint64_t *kk, const uint32_t *vv, uint64_t n) {
          k.insert(k.end(), kk, kk + n); v.insert(v.end(), vv, vv + n);
        });
        if (st[f - f0] != MFN_OK) errs[f - f0] = mfn_last_error();
      });
    for (auto &t : th) t.join();
    for (uint32_t i = 0; i < f1 - f0 && status == MFN_OK; i++) {
      if (st[i] != MFN_OK) { status = st[i]; mfn_set_error("%s", errs[i].c_str()); break; }
      uint64_t n = K[i].size(), o = 0;
      while (o < n && status == MFN_OK) {
        uint64_t m = std::min(n - o, CH - fill);
        memcpy(hk + fill, K[i].data() + o, m * 8);
        memcpy(hv + fill, V[i].data() + o, m * 4);
        fill += m; o += m;
        if (fill == CH) status = flush();
      }
    }
  }
  if (status == MFN_OK) status = flush();
  cudaFree(dk); cudaFree(dv); cudaFreeHost(hk); cudaFreeHost(hv);
  return status;
}

// ---------------------------------------------------------------------------
// mfn_table_load
// ---------------------------------------------------------------------------
extern "C" int mfn_table_load(mfn_ctx *ctx, const char *dir, uint32_t minV, uint32_t maxV, mfn_table **out) {
  if (!ctx || !dir || !out) return MFN_ERR_INVALID;
  MerylIndex idx;
  MFN_TRY(meryl_open(dir, idx));
  if (idx.k > 64) { mfn_set_error("'%s': k=%u; k-mers hold at most 64 bases", dir, idx.k); return MFN_ERR_UNSUPPORTED; }
  if (idx.flags & 1) { mfn_set_error("'%s' is a multiset database; not supported", dir); return MFN_ERR_UNSUPPORTED; }
  if (idx.k > 32) return mfn_wide_load(ctx, idx, minV, maxV, out);          // 128-bit k-mers: wide.cu
  const uint32_t nfiles = 1u << idx.numFilesBits;
  if (idx.version != 3) {
    // older master indexes carry a different statistics layout: take the value histogram from the data
    std::vector<uint64_t> cnt;
    for (uint32_t f = 0; f < nfiles; f++)
      MFN_TRY(meryl_decode_file(idx, f, [&](const uint64_t *, const uint32_t *v, uint64_t n) {
        for (uint64_t i = 0; i < n; i++) { if (v[i] >= cnt.size()) cnt.resize((size_t)v[i] + 1, 0); cnt[v[i]]++; }
      }));
    for (size_t v = 0; v < cnt.size(); v++) if (cnt[v]) { idx.histValue.push_back(v); idx.histCount.push_back(cnt[v]); }
  }
  if (minV == 0) minV = 1;                                                 // kmers-exact.C:43
  if (maxV == 0xffffffffu) maxV = idx.histValue.empty() ? 0 : (uint32_t)idx.histValue.back();   // :44, kmers-histogram.H:73
  uint64_t keep = 0;                                                       // kmers-exact.C:64-72
  for (size_t i = 0; i < idx.histValue.size(); i++)
    if (idx.histValue[i] >= minV && idx.histValue[i] <= maxV) keep += idx.histCount[i];
  mfn_table *t = nullptr;
  MFN_TRY(mfn_table_alloc(ctx, idx.k, keep, &t, [&](uint32_t vmax) {       // from the value histogram
    uint64_t c = 0;
    for (size_t i = 0; i < idx.histValue.size(); i++)
      if (idx.histValue[i] >= minV && idx.histValue[i] <= maxV && idx.histValue[i] - (minV - 1) > vmax) c += idx.histCount[i];
    return c;
  }));
  t->min_value = minV; t->max_value = maxV; t->value_offset = minV - 1;
  int st = meryl_stream(ctx, idx, [&](const unsigned long long *dk, const uint32_t *dv, uint64_t n) {
    return mfn_table_insert_device(t, dk, dv, n, minV, maxV, minV - 1, false);
  });
  if (st == MFN_OK) st = mfn_table_finalize(t, false);
  if (st == MFN_OK && t->host_stats.n_dup) {
    mfn_set_error("'%s' holds duplicate k-mers (%llu); multiset databases are not supported", dir, t->host_stats.n_dup);
    st = MFN_ERR_UNSUPPORTED;
  }
  if (st != MFN_OK) { mfn_table_free(t); return st; }
  *out = t;
  return MFN_OK;
}

// ---------------------------------------------------------------------------
// writer
// ---------------------------------------------------------------------------
namespace {

struct BitWriter {
  uint64_t maxBits;
  std::vector<std::vector<uint64_t>> blocks;
  std::vector<uint64_t> lens;
  uint64_t pos = 0;
  explicit BitWriter(uint64_t mb) : maxBits((mb + 63) & ~63ull) { blocks.emplace_back(maxBits / 64, 0); lens.push_back(0); }
  void room(uint64_t width) {
    if (pos + width <= maxBits) return;
    lens.back() = pos;
    blocks.emplace_back(maxBits / 64, 0); lens.push_back(0); pos = 0;
  }
  void binary(uint32_t width, uint64_t v) {
    if (width == 0) return;
    room(width);
    if (width < 64) v &= (1ull << width) - 1;
    auto &b = blocks.back();
    uint64_t wi = pos >> 6, bo = pos & 63;
    if (bo + width <= 64) b[wi] |= v << (64 - bo - width);
    else { uint32_t w2 = (uint32_t)(bo + width - 64); b[wi] |= v >> w2; b[wi + 1] |= v << (64 - w2); }
    pos += width;
  }
  void unary(uint64_t v) {
    room(v + 1);
    pos += v;                                    // v zero bits (memory is already clear)
    blocks.back()[pos >> 6] |= 1ull << (63 - (pos & 63));
    pos += 1;
  }
  bool dump(FILE *F) {
    lens.back() = pos;
    uint32_t n = 0;
    while (n < lens.size() && lens[n] > 0) n++;
    uint32_t bmax = std::max<uint32_t>(32, (n + 31) / 32 * 32);
    bool ok = fwrite(&maxBits, 8, 1, F) == 1 && fwrite(&n, 4, 1, F) == 1 && fwrite(&bmax, 4, 1, F) == 1;
    uint64_t bgn = 0;
    for (uint32_t i = 0; i < n && ok; i++) { ok = fwrite(&bgn, 8, 1, F) == 1; bgn += lens[i]; }
    for (uint32_t i = 0; i < n && ok; i++) ok = fwrite(&lens[i], 8, 1, F) == 1;
    for (uint32_t i = 0; i < n && ok; i++) {
      size_t nw = (lens[i] + 63) / 64;
      ok = fwrite(blocks[i].data(), 8, nw, F) == nw;
    }
    return ok;
  }
};

}  // namespace

extern "C" int mfn_meryl_info(const char *dir, mfn_meryl_info_t *info) {
  if (!dir || !info) return MFN_ERR_INVALID;
  MerylIndex idx;
  MFN_TRY(meryl_open(dir, idx));
  memset(info, 0, sizeof(*info));
  info->k = idx.k; info->version = idx.version;
  info->prefix_bits = idx.prefixSize; info->suffix_bits = idx.suffixSize;
  info->files = 1u << idx.numFilesBits; info->blocks_per_file = 1u << idx.numBlocksBits;
  info->multiset = idx.flags & 1u; info->hist_len = (uint32_t)idx.histValue.size();
  info->n_unique = idx.numUnique; info->n_distinct = idx.numDistinct; info->n_total = idx.numTotal;
  return MFN_OK;
}

extern "C" int mfn_meryl_histogram(const char *dir, uint64_t *values, uint64_t *counts, uint64_t cap, uint64_t *n) {
  if (!dir || !n) return MFN_ERR_INVALID;
  MerylIndex idx;
  MFN_TRY(meryl_open(dir, idx));
  *n = idx.histValue.size();
  if (*n > cap || (*n && (!values || !counts))) return MFN_ERR_CAPACITY;
  for (uint64_t i = 0; i < *n; i++) { values[i] = idx.histValue[i]; counts[i] = idx.histCount[i]; }
  return MFN_OK;
}

extern "C" int mfn_meryl_write(const char *dir, uint32_t k, const uint64_t *kmers, const uint32_t *values, uint64_t n) {
  if (k > 32) { mfn_set_error("mfn_meryl_write: k=%u; k-mers are passed in one 64-bit word (k <= 32)", k); return MFN_ERR_UNSUPPORTED; }
  if (!dir || k < 4 || (n && (!kmers || !values))) { mfn_set_error("mfn_meryl_write: need 4 <= k <= 32"); return MFN_ERR_INVALID; }
  const uint32_t filesBits = 6;
  uint32_t blocksBits = 0;
  while (blocksBits < 16 && (n >> (filesBits + blocksBits)) > 8192) blocksBits++;
  if (2 * k < filesBits + blocksBits + 1) blocksBits = 2 * k - filesBits - 1;
  const uint32_t prefixSize = filesBits + blocksBits, suffixSize = 2 * k - prefixSize;

  // bucket the pairs by data file (top 6 bits), then every file is sorted and written on its own
  // thread: the 64 files are independent
  struct KV { uint64_t k; uint32_t v; };
  std::vector<uint64_t> fstart(65, 0);
  for (uint64_t i = 0; i < n; i++) fstart[(kmers[i] >> (2 * k - filesBits)) + 1]++;
  for (int f = 0; f < 64; f++) fstart[f + 1] += fstart[f];
  std::vector<KV> kv(n);
  {
    std::vector<uint64_t> fill(fstart.begin(), fstart.end() - 1);
    for (uint64_t i = 0; i < n; i++) kv[fill[kmers[i] >> (2 * k - filesBits)]++] = {kmers[i], values[i]};
  }
  mkdir(dir, 0777);
  MerylIndex idx; idx.dir = dir; idx.numFilesBits = filesBits;
  std::vector<int> fstatus(64, MFN_OK);
  std::vector<std::map<uint32_t, uint64_t>> fhist(64);
  auto write_file = [&](uint32_t f) {
    KV *seg = kv.data() + fstart[f];
    const uint64_t m = fstart[f + 1] - fstart[f];
    std::sort(seg, seg + m, [](const KV &a, const KV &b) { return a.k < b.k; });
    for (uint64_t i = 1; i < m; i++) if (seg[i].k == seg[i - 1].k) { fstatus[f] = MFN_ERR_INVALID; return; }
    {
      // value histogram of this file (small values dominate: count them in a flat array)
      std::vector<uint64_t> small(1 << 16, 0);
      auto &big = fhist[f];
      for (uint64_t i = 0; i < m; i++) { if (seg[i].v < small.size()) small[seg[i].v]++; else big[seg[i].v]++; }
      for (uint32_t v = 0; v < small.size(); v++) if (small[v]) big[v] += small[v];
    }
    FILE *D = fopen(meryl_file_name(idx, f, false).c_str(), "wb");
    FILE *I = fopen(meryl_file_name(idx, f, true).c_str(), "wb");
    if (!D || !I) { if (D) fclose(D); if (I) fclose(I); fstatus[f] = MFN_ERR_IO; return; }
    std::vector<uint64_t> index((size_t)3 << blocksBits);
    for (uint64_t b = 0; b < (1ull << blocksBits); b++) { index[3 * b] = 0; index[3 * b + 1] = ~0ull; index[3 * b + 2] = 0; }
    bool ok = true;
    uint64_t i = 0;
    while (ok && i < m) {
      uint64_t prefix = suffixSize >= 64 ? 0 : seg[i].k >> suffixSize;
      uint64_t j = i;
      while (j < m && (suffixSize >= 64 ? 0 : seg[j].k >> suffixSize) == prefix) j++;
      uint64_t nk = j - i;
      uint32_t unaryBits = 0; uint64_t unarySum = 1;
      while (unarySum < nk) { unaryBits++; unarySum <<= 1; }
      if (unaryBits > suffixSize) unaryBits = suffixSize;
      uint32_t binaryBits = suffixSize - unaryBits;
      uint64_t blockSize = 10 * 64 + 2 * unarySum + nk * binaryBits / 16 + nk * 32 / 16;
      blockSize = (blockSize & ~0x3ffull) + 1024;
      BitWriter w(blockSize);
      w.binary(64, MAGIC_DATA_1); w.binary(64, MAGIC_DATA_2);
      w.binary(64, prefix); w.binary(64, nk);
      w.binary(8, 1); w.binary(32, unaryBits); w.binary(32, binaryBits); w.binary(64, 0);
      w.binary(8, 1); w.binary(64, 0); w.binary(64, 0);
      const uint64_t smask = suffixSize >= 64 ? ~0ull : ((1ull << suffixSize) - 1);
      uint64_t last = 0;
      for (uint64_t x = i; x < j; x++) {
        uint64_t suf = seg[x].k & smask;
        uint64_t hi = binaryBits >= 64 ? 0 : suf >> binaryBits;
        w.unary(hi - last);
        w.binary(binaryBits, suf);
        last = hi;
      }
      for (uint64_t x = i; x < j; x++) w.binary(32, seg[x].v);
      uint64_t blk = prefix & ((1ull << blocksBits) - 1);
      index[3 * blk] = prefix; index[3 * blk + 1] = (uint64_t)ftell(D); index[3 * blk + 2] = nk;
      ok = w.dump(D);
      i = j;
    }
    ok = ok && fwrite(index.data(), 8, index.size(), I) == index.size();
    fclose(D); fclose(I);
    if (!ok) fstatus[f] = MFN_ERR_IO;
  };
  {
    const unsigned nthreads = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::atomic<uint32_t> next{0};
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nthreads; t++)
      th.emplace_back([&]() { for (uint32_t f; (f = next++) < 64;) write_file(f); });
    for (auto &t : th) t.join();
  }
  for (int f = 0; f < 64; f++) {
    if (fstatus[f] == MFN_ERR_INVALID) { mfn_set_error("mfn_meryl_write: duplicate k-mer"); return MFN_ERR_INVALID; }
    if (fstatus[f] != MFN_OK) { mfn_set_error("cannot write under '%s'", dir); return MFN_ERR_IO; }
  }

  // master index with the value histogram (the reference sizes its table from it)
  std::map<uint32_t, uint64_t> hist;
  for (auto &h : fhist) for (auto &e : h) hist[e.first] += e.second;
  uint64_t numUnique = 0, numTotal = 0;
  for (auto &e : hist) { numTotal += (uint64_t)e.first * e.second; if (e.first == 1) numUnique = e.second; }
  BitWriter m(32 * 1024);
  m.binary(64, MAGIC_INDEX_1); m.binary(64, MAGIC_INDEX_V3);
  m.binary(32, prefixSize); m.binary(32, suffixSize); m.binary(32, filesBits); m.binary(32, blocksBits); m.binary(32, 0);
  m.binary(64, numUnique); m.binary(64, n); m.binary(64, numTotal); m.binary(64, (uint64_t)hist.size());
  for (auto &e : hist) { m.binary(64, e.first); m.binary(64, e.second); }
  FILE *M = fopen((std::string(dir) + "/merylIndex").c_str(), "wb");
  if (!M) { mfn_set_error("cannot write '%s/merylIndex'", dir); return MFN_ERR_IO; }
  bool ok = m.dump(M);
  fclose(M);
  return ok ? MFN_OK : MFN_ERR_IO;
}
