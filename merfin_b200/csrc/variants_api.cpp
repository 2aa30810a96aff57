// C ABI over the variant front-end (host/variants.cpp): VCF handle, clustering, and
// processVariants for one sequence record with the GPU as the path scorer (mfn_score_paths).
// Plain C++ above the public ABI; no device code here.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/merfin_b200.h"
#include "host/variants.h"

void mfn_set_error(const char *fmt, ...);

struct mfn_vcf {
  mfnv::VcfFile file;
  bool clustered = false;
  uint32_t k = 0;
};

// ---- pinned result buffers, pooled across calls and threads -------------------------------------
namespace {
struct PinnedBlock { void *p = nullptr; size_t cap = 0; };
struct PinnedPool {
  std::mutex mu;
  std::vector<PinnedBlock> free_;
  size_t held = 0;                                       // bytes in the free list
  PinnedBlock get(size_t bytes) {
    {
      std::lock_guard<std::mutex> lk(mu);
      size_t best = free_.size();
      for (size_t i = 0; i < free_.size(); i++)
        if (free_[i].cap >= bytes && (best == free_.size() || free_[i].cap < free_[best].cap)) best = i;
      if (best < free_.size()) { PinnedBlock b = free_[best]; free_.erase(free_.begin() + best); held -= b.cap; return b; }
    }
    PinnedBlock b;
    const size_t cap = std::max<size_t>(bytes + bytes / 4, 1 << 20);
    if (mfn_host_alloc(&b.p, cap) != MFN_OK) return PinnedBlock{};
    b.cap = cap;
    return b;
  }
  void put(PinnedBlock b) {
    if (!b.p) return;
    std::lock_guard<std::mutex> lk(mu);
    // keep at most 64 blocks / 2 GB around
    if (free_.size() >= 64 || held + b.cap > (2ull << 30)) { mfn_host_free(b.p); return; }
    held += b.cap;
    free_.push_back(b);
  }
};
PinnedPool g_pool;
struct Held { PinnedBlock b; ~Held() { g_pool.put(b); } };
}  // namespace

extern "C" int mfn_vcf_load(const char *path, mfn_vcf **out) {
  if (!path || !out) return MFN_ERR_INVALID;
  std::unique_ptr<mfn_vcf> v(new mfn_vcf());
  std::string err;
  if (!v->file.load(path, err)) { mfn_set_error("%s", err.c_str()); return MFN_ERR_IO; }
  *out = v.release();
  return MFN_OK;
}

extern "C" int mfn_vcf_cluster(mfn_vcf *v, uint32_t k, uint32_t comb, int nosplit) {
  if (!v || k == 0) return MFN_ERR_INVALID;
  if (v->clustered) { mfn_set_error("the variants of this VCF are clustered already"); return MFN_ERR_INVALID; }
  v->file.mergeChrPosGT(k, comb, nosplit != 0);
  v->clustered = true;
  v->k = k;
  return MFN_OK;
}

extern "C" int mfn_vcf_info(const mfn_vcf *v, mfn_vcf_info_t *info) {
  if (!v || !info) return MFN_ERR_INVALID;
  info->headers = v->file.headers.size();
  info->records = v->file.num_records;
  info->contigs = v->file.byChr.size();
  info->contig_ids = (uint64_t)v->file.numChr;
  info->excluded = v->file.excluded;
  info->clusters = 0;
  for (const auto &kv : v->file.byChr) info->clusters += kv.second.size();
  return MFN_OK;
}

extern "C" int mfn_vcf_header(const mfn_vcf *v, char *buf, uint64_t cap, uint64_t *need) {
  if (!v || !need) return MFN_ERR_INVALID;
  uint64_t n = 0;
  for (const std::string &h : v->file.headers) n += h.size() + 1;
  *need = n;
  if (!buf) return MFN_OK;
  if (cap < n) return MFN_ERR_CAPACITY;
  char *p = buf;
  for (const std::string &h : v->file.headers) { memcpy(p, h.data(), h.size()); p += h.size(); *p++ = '\n'; }
  return MFN_OK;
}

extern "C" uint64_t mfn_vcf_clusters_of(const mfn_vcf *v, const char *ident) {
  if (!v || !ident) return 0;
  auto it = v->file.byChr.find(ident);
  return it == v->file.byChr.end() ? 0 : it->second.size();
}

extern "C" void mfn_vcf_free(mfn_vcf *v) { delete v; }

extern "C" void mfn_text_free(char *p) { free(p); }

static char *dup_text(const std::string &s) {
  char *p = (char *)malloc(s.size() + 1);
  if (!p) return nullptr;
  memcpy(p, s.data(), s.size());
  p[s.size()] = 0;
  return p;
}

extern "C" int mfn_variants_process(mfn_ctx *ctx, const mfn_table *reads, const mfn_table *asm_, mfn_vcf *v,
                                    const char *ident, const char *seq, uint64_t seq_len, int mode, uint32_t comb,
                                    const mfn_kmodel *km, uint32_t shard, uint32_t nshards,
                                    char **result, uint64_t *result_len, char **debug, uint64_t *debug_len, uint64_t *var_mer_id) {
  if (!ctx || !reads || !asm_ || !v || !ident || !seq || !km || !result || nshards == 0 || shard >= nshards) return MFN_ERR_INVALID;
  if (mode < MFN_MODE_FILTER || mode > MFN_MODE_LOOSE) { mfn_set_error("unknown variant mode %d", mode); return MFN_ERR_INVALID; }
  if (!v->clustered) { mfn_set_error("mfn_vcf_cluster has not been called"); return MFN_ERR_INVALID; }
  *result = nullptr;
  if (result_len) *result_len = 0;
  if (debug) *debug = nullptr;
  if (debug_len) *debug_len = 0;
  auto it = v->file.byChr.find(ident);
  if (it == v->file.byChr.end()) { *result = dup_text(std::string()); return *result ? MFN_OK : MFN_ERR_NOMEM; }

  mfnv::Options opt;
  opt.k = v->k; opt.comb = comb; opt.km = *km; opt.debug = debug != nullptr;
  opt.mode = (mfnv::Mode)mode;
  // the GPU as the scorer: results land in a pinned block that lives as long as the ScoreOut
  mfnv::PathScorer scorer = [&](const std::string &paths, const std::vector<uint64_t> &off, mfnv::ScoreOut &out) -> int {
    const size_t n = (paths.size() + 1 + 15) & ~(size_t)15;
    const size_t bytes = n * 9;
    Held *h = (Held *)out.hold.get();
    if (!h || h->b.cap < bytes) {
      out.hold.reset();                                  // hand the smaller block back first
      PinnedBlock b = g_pool.get(bytes);
      if (!b.p) return MFN_ERR_NOMEM;
      std::shared_ptr<Held> nh(new Held());
      nh->b = b;
      h = nh.get();
      out.hold = nh;
    }
    uint32_t *vr = (uint32_t *)h->b.p, *va = vr + n;
    uint8_t *ok = (uint8_t *)(va + n);
    int st = mfn_score_paths(ctx, reads, asm_, paths.data(), off.data(), off.size() - 1, vr, va, ok);
    out.v_read = vr; out.v_asm = va; out.valid = ok;
    return st;
  };
  const size_t nc = it->second.size();
  const size_t cb = nc * shard / nshards, ce = nc * (shard + 1) / nshards;
  size_t batch = 32768;                                  // clusters per device call (~3 MB of paths, ~30 MB of results)
  if (const char *env = getenv("MFN_VARIANT_BATCH")) batch = (size_t)std::max(1, atoi(env));
  std::string res, dbg;
  uint64_t id = var_mer_id ? *var_mer_id : 0;
  int st = mfnv::process_sequence(ident, seq, seq_len, it->second, opt, scorer, res, debug ? &dbg : nullptr, id, cb, ce, batch);
  if (st != 0) return st < 0 ? st : MFN_ERR_CUDA;
  if (var_mer_id) *var_mer_id = id;
  *result = dup_text(res);
  if (!*result) return MFN_ERR_NOMEM;
  if (result_len) *result_len = res.size();
  if (debug) {
    *debug = dup_text(dbg);
    if (!*debug) { free(*result); *result = nullptr; return MFN_ERR_NOMEM; }
    if (debug_len) *debug_len = dbg.size();
  }
  return MFN_OK;
}
