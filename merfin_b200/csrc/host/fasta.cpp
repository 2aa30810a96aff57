#include "fasta.h"

#include <immintrin.h>

#include <algorithm>
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

// Append the non-whitespace bytes of [p, p + n) to `bases`.  A sequence line is normally free of
// whitespace, so blocks of 32 bytes without any byte <= ' ' are copied as they are (AVX2 when the
// CPU has it); anything else goes through the per-byte filter.
__attribute__((target("avx2"))) static size_t clean_prefix_avx2(const char *p, size_t n) {
  size_t i = 0;
  const __m256i lim = _mm256_set1_epi8(0x21);
  for (; i + 32 <= n; i += 32) {
    const __m256i v = _mm256_loadu_si256((const __m256i *)(p + i));
    // bytes 0x00..0x20 (and none of 0x80..0xff, which compare as negative: checked below)
    const __m256i low = _mm256_cmpgt_epi8(lim, v);
    if (_mm256_movemask_epi8(low)) break;
  }
  return i;
}
static void append_bases(std::vector<char> &bases, const char *p, size_t n) {
  static const bool avx2 = (__builtin_cpu_init(), __builtin_cpu_supports("avx2"));
  size_t i = 0;
  while (i < n) {
    if (avx2 && n - i >= 32) {
      const size_t c = clean_prefix_avx2(p + i, n - i);
      if (c) { bases.insert(bases.end(), p + i, p + i + c); i += c; continue; }
    }
    const size_t e = std::min(n, i + 32);
    for (; i < e; i++) if (!is_space(p[i])) bases.push_back(p[i]);
  }
}

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
        {
          const char *m1 = (const char *)memchr(data.data() + p, '>', e - p);
          if (m1) stop = (size_t)(m1 - data.data());
          const char *m2 = (const char *)memchr(data.data() + p, '@', stop - p);
          if (m2) stop = (size_t)(m2 - data.data());
        }
        append_bases(out.bases, data.data() + p, stop - p);
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
