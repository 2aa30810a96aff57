// Variant front-end of the merfin-compatible CLI (-filter / -polish / -better / -strict / -loose):
// VCF loading, clustering of nearby variants, enumeration of allele combinations ("paths"),
// replay of varMer::score from integer (v_read, v_asm) pairs, and the path selectors.
//
// Host-only control logic written against the behaviour of the reference (file:line in the
// reference tree): vcfRecord.H:50-76, vcf.C:23-87 (genotype -> allele list), vcf.C:156-246
// (clustering), merfin-variants.C:22-126 (path enumeration), :131-310 (per-cluster driver),
// varMer.C:48-145 (score), :150-467 (selectors), :472-550 (output records), :553-659 (stats).
// The k-mer probes themselves are NOT done here: a PathScorer (the GPU library) returns the
// integer pairs for every window of every path.
#pragma once
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <string_view>
#include <vector>

#include "../../../include/merfin_b200.h"

namespace mfnv {

enum Mode { FILTER = 4, POLISH = 5, BETTER = 6, STRICT = 7, LOOSE = 8 };   // merfin-globals.H:32-40

// Inline storage for the first N elements (a cluster nearly always holds one or two variants, a genotype
// two or three alleles): no allocation per VCF record.  T is trivially copyable.
template <class T, unsigned N> class SmallVec {
  T inl_[N];
  std::vector<T> rest_;
  uint32_t n_ = 0;
 public:
  size_t size() const { return n_; }
  bool empty() const { return n_ == 0; }
  void push_back(const T &v) { if (n_ < N) inl_[n_] = v; else rest_.push_back(v); n_++; }
  const T &operator[](size_t i) const { return i < N ? inl_[i] : rest_[i - N]; }
  T &operator[](size_t i) { return i < N ? inl_[i] : rest_[i - N]; }
};

// One data line.  The text itself stays in the VcfFile's copy of the file; the ten columns are spans of it.
struct VcfRecord {
  enum { CHROM, POS, ID, REF, ALT, QUAL, FILTER, INFO, FORMAT, SAMPLE, NCOL };
  const char *line = nullptr;
  uint32_t col_off[NCOL] = {}, col_len[NCOL] = {};
  uint32_t pos = 0;
  double qual = 0.0;
  std::string_view col(int i) const { return std::string_view(line + col_off[i], col_len[i]); }
  std::string save() const;                 // vcfRecord.H:78-99: the first ten columns, QUAL as %.1f
};

struct GtAllele {                           // vcf.H:37-50
  const VcfRecord *record = nullptr;
  uint32_t pos = 0, refLen = 0;
  double qual = 0.0;
  SmallVec<std::string_view, 3> alleles;    // [0] is the reference allele; may be EMPTY (Q10)
};

struct PosGT {                              // vcf.H:59-90: a cluster of nearby variants
  const std::string *chr = nullptr;         // (the key of VcfFile::byChr)
  uint32_t rStart = 0, rEnd = 0;
  SmallVec<GtAllele *, 2> gts;
};

// The bytes of a file: a read-only mapping of a plain file (no copy, no page touched twice), or what the decompressor
// of a .gz / .bz2 / .xz file wrote.
struct FileText {
  const char *p = nullptr;
  size_t n = 0;
  FileText() = default;
  FileText(const FileText &) = delete;
  FileText &operator=(const FileText &) = delete;
  ~FileText();
  bool load(const std::string &path, std::string &err);
 private:
  std::vector<char> owned_;
  void *map_ = nullptr;
  size_t map_len_ = 0;
};

struct VcfFile {
  std::vector<std::string> headers;
  uint64_t num_records = 0;
  std::map<std::string, std::vector<PosGT *>> byChr;
  uint64_t excluded = 0;
  int numChr = 0;
  bool load(const std::string &path, std::string &err);              // vcf.C:93-151
  void mergeChrPosGT(uint32_t ksize, uint32_t comb, bool nosplit);   // vcf.C:156-246
 private:
  // the file's text and the objects of every data line, in the pieces the parsing threads made them in
  FileText text_;
  std::vector<std::vector<VcfRecord>> records_;
  std::vector<std::vector<GtAllele>> alleles_;
  std::vector<std::vector<PosGT>> posgts_;
};

// paths: concatenated path strings + offsets (as mfn_score_paths takes them); the scorer leaves one
// (v_read, v_asm, valid) per base position (window start) of every path in `out`: either in out's own
// vectors or in memory it owns itself (pinned buffers of the GPU scorer), kept alive through out.hold.
struct ScoreOut {
  const uint32_t *v_read = nullptr, *v_asm = nullptr;
  const uint8_t *valid = nullptr;
  std::vector<uint32_t> vr_store, va_store;
  std::vector<uint8_t> ok_store;
  std::shared_ptr<void> hold;
};
typedef std::function<int(const std::string &paths, const std::vector<uint64_t> &off, ScoreOut &out)> PathScorer;

struct Options {
  uint32_t k = 0, comb = 15;
  Mode mode = POLISH;
  mfn_kmodel km{};
  bool debug = false;
};

// All clusters of one sequence record: returns the VCF text merfin appends to its output for this
// record (processVariants' s->result) and, if opt.debug, the debug lines.  `varMerId` numbers
// the debug lines across records.
// Clusters [cluster_begin, cluster_end) only (a shard of the record; the outputs of consecutive ranges
// concatenate to the whole record's), scored in batches of batch_clusters (0: all at once).
int process_sequence(const std::string &ident, const char *seq, uint64_t seqLen, std::vector<PosGT *> &clusters,
                     const Options &opt, const PathScorer &scorer, std::string &result, std::string *debug,
                     uint64_t &varMerId, size_t cluster_begin = 0, size_t cluster_end = (size_t)-1, size_t batch_clusters = 0);

}  // namespace mfnv
