#include "fasta.h"

#include <cstdio>
#include <cstring>

static bool ends_with(const std::string &s, const char *suf) {
  size_t n = strlen(suf);
  return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

bool slurp_file(const std::string &path, std::vector<char> &data, std::string &err) {
  FILE *F = nullptr;
  bool piped = false;
  std::string cmd;
  if (ends_with(path, ".gz")) cmd = "gzip -dc '" + path + "'";
  else if (ends_with(path, ".bz2")) cmd = "bzip2 -dc '" + path + "'";
  else if (ends_with(path, ".xz")) cmd = "xz -dc '" + path + "'";
  if (!cmd.empty()) {
    FILE *probe = fopen(path.c_str(), "rb");
    if (!probe) { err = "cannot open '" + path + "'"; return false; }
    fclose(probe);
    F = popen(cmd.c_str(), "r");
    piped = true;
  } else {
    F = fopen(path.c_str(), "rb");
  }
  if (!F) { err = "cannot open '" + path + "'"; return false; }
  data.clear();
  std::vector<char> buf(1 << 24);
  size_t n;
  while ((n = fread(buf.data(), 1, buf.size(), F)) > 0) data.insert(data.end(), buf.begin(), buf.begin() + n);
  if (piped) pclose(F); else fclose(F);
  return true;
}

static inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; }

bool load_sequences(const std::string &path, SeqSet &out, std::string &err) {
  std::vector<char> data;
  if (!slurp_file(path, data, err)) return false;
  out.ident.clear(); out.off.assign(1, 0); out.bases.clear();
  out.bases.reserve(data.size());
  const size_t N = data.size();
  size_t p = 0;
  while (p < N) {
    while (p < N && is_space(data[p])) p++;
    if (p >= N) break;
    const char mark = data[p];
    if (mark != '>' && mark != '@') break;              // not at a record start: stop, like the reference
    size_t h0 = ++p;
    while (p < N && data[p] != '\n') p++;
    size_t h1 = p;
    while (h0 < h1 && is_space(data[h0])) h0++;          // ident = first word of the header
    size_t he = h0;
    while (he < h1 && !is_space(data[he]) && (unsigned char)data[he] > 32 && (unsigned char)data[he] < 127) he++;
    out.ident.emplace_back(data.data() + h0, he - h0);
    if (p < N) p++;
    if (mark == '>') {
      // body: everything up to the next '>' or '@' (anywhere), whitespace dropped
      while (p < N && data[p] != '>' && data[p] != '@') {
        const char *nl = (const char *)memchr(data.data() + p, '\n', N - p);
        size_t e = nl ? (size_t)(nl - data.data()) : N;
        // a marker inside the line ends the record there
        size_t stop = e;
        for (size_t q = p; q < e; q++) if (data[q] == '>' || data[q] == '@') { stop = q; break; }
        for (size_t q = p; q < stop; q++) if (!is_space(data[q])) out.bases.push_back(data[q]);
        p = stop;
        if (stop == e && p < N) p++;                     // consume the newline
      }
    } else {
      // FASTQ: one sequence line, '+' line, one quality line
      while (p < N && is_space(data[p])) p++;
      while (p < N && data[p] != '\n') { if (!is_space(data[p])) out.bases.push_back(data[p]); p++; }
      while (p < N && is_space(data[p])) p++;
      if (p < N && data[p] == '+') {
        while (p < N && data[p] != '\n') p++;
        while (p < N && is_space(data[p])) p++;
        while (p < N && data[p] != '\n') p++;
      }
    }
    out.off.push_back(out.bases.size());
  }
  return true;
}
