// Bulk FASTA/FASTQ ingest for the merfin-compatible front-end.
// Record semantics follow the reference loader (sequence.C:822-990, ident sequence.C:556-579):
//   * whitespace inside a sequence is skipped; a FASTA record body ends at '>' or '@' or EOF;
//   * ident = header up to the first whitespace;
//   * .gz / .bz2 / .xz inputs are read through the matching decompressor (files-compressed.C:95-141).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

struct SeqSet {
  std::vector<std::string> ident;
  std::vector<uint64_t> off;        // nseq + 1 offsets into bases
  std::vector<char> bases;          // all records concatenated
  size_t nseq() const { return ident.size(); }
  uint64_t length(size_t i) const { return off[i + 1] - off[i]; }
  const char *seq(size_t i) const { return bases.data() + off[i]; }
};

// Reads every record of `path`.  Returns false (and sets err) if the file cannot be read.
bool load_sequences(const std::string &path, SeqSet &out, std::string &err);
// Whole file (through a decompressor if the name says so) into memory.
bool slurp_file(const std::string &path, std::vector<char> &data, std::string &err);
