#include "fasta.h"

#include <immintrin.h>

#include <algorithm>
#include <cstdio>
#include <cstring>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

static bool ends_with(const std::string &s, const char *suf) {
  size_t n = strlen(suf);
  return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

bool slurp_file(const std::string &path, std::vector<char> &data, std::string &err) {
  FILE *F = nullptr;
  bool piped = false;
  std::string cmd;
  // the path goes through /bin/sh: single quotes, every embedded single quote written as '\''
  std::string q = "'";
  for (char c : path) { if (c == '\'') q += "'\\''"; else q += c; }
  q += "'";
  if (ends_with(path, ".gz")) cmd = "gzip -dc " + q;
  else if (ends_with(path, ".bz2")) cmd = "bzip2 -dc " + q;
  else if (ends_with(path, ".xz")) cmd = "xz -dc " + q;
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
  size_t n;
  if (!piped) {                                       // a plain file: one read of its size (more follows if it grew)
    struct stat st;
    if (fstat(fileno(F), &st) == 0 && st.st_size > 0) {
      data.resize((size_t)st.st_size);
      n = fread(data.data(), 1, data.size(), F);
      data.resize(n);
    }
  }
  std::vector<char> buf(1 << 22);
  while ((n = fread(buf.data(), 1, buf.size(), F)) > 0) data.insert(data.end(), buf.begin(), buf.begin() + n);
  if (piped) pclose(F); else fclose(F);
  return true;
}

static inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; }

// Copy a FASTA body starting at p (input ends at e) to dst, dropping whitespace, until a record
// marker ('>' or '@', anywhere) or the end of the input.  Returns the input position reached; dst is
// advanced.  dst must have room for every remaining input byte plus 32 (blocks are stored whole and
// the write pointer then moves by the bytes kept).
static inline const char *copy_body_scalar(const char *p, const char *e, char *&dst) {
  char *d = dst;
  for (; p < e && *p != '>' && *p != '@'; p++) if (!is_space(*p)) *d++ = *p;
  dst = d;
  return p;
}
__attribute__((target("avx2"))) static const char *copy_body_avx2(const char *p, const char *e, char *&dst) {
  char *d = dst;
  const __m256i lim = _mm256_set1_epi8(0x21), gt = _mm256_set1_epi8('>'), at = _mm256_set1_epi8('@');
  while (p + 32 <= e) {
    const __m256i v = _mm256_loadu_si256((const __m256i *)p);
    _mm256_storeu_si256((__m256i *)d, v);
    // bytes that need a look: 0x00..0x20 (and 0x80..0xff, which compare as negative), '>' and '@'
    const __m256i sp = _mm256_or_si256(_mm256_cmpgt_epi8(lim, v), _mm256_or_si256(_mm256_cmpeq_epi8(v, gt), _mm256_cmpeq_epi8(v, at)));
    const uint32_t m = (uint32_t)_mm256_movemask_epi8(sp);
    if (m == 0) { p += 32; d += 32; continue; }
    const uint32_t n = (uint32_t)__builtin_ctz(m);
    p += n; d += n;                                    // the bytes before it are in place
    const char c = *p;
    if (c == '>' || c == '@') { dst = d; return p; }
    if (!is_space(c)) *d++ = c;
    p++;
  }
  dst = d;
  return copy_body_scalar(p, e, dst);
}
static const char *copy_body(const char *p, const char *e, char *&dst) {
  static const bool avx2 = (__builtin_cpu_init(), __builtin_cpu_supports("avx2"));
  return avx2 ? copy_body_avx2(p, e, dst) : copy_body_scalar(p, e, dst);
}

// The records of a whole file image [data, data + N).
static void parse_records(const char *data, size_t N, SeqSet &out) {
  out.ident.clear(); out.off.assign(1, 0); out.bases.clear();
  out.bases.resize(N + 64);                             // every base is an input byte; + room for whole-block stores
  char *const base = out.bases.data();
  char *dst = base;
  size_t p = 0;
  while (p < N) {
    while (p < N && is_space(data[p])) p++;
    if (p >= N) break;
    const char mark = data[p];
    if (mark != '>' && mark != '@') break;              // not at a record start: stop, like the reference
    size_t h0 = ++p;
    const char *nl = (const char *)memchr(data + p, '\n', N - p);
    p = nl ? (size_t)(nl - data) : N;
    size_t h1 = p;
    while (h0 < h1 && is_space(data[h0])) h0++;          // ident = first word of the header
    size_t he = h0;
    while (he < h1 && !is_space(data[he]) && (unsigned char)data[he] > 32 && (unsigned char)data[he] < 127) he++;
    out.ident.emplace_back(data + h0, he - h0);
    if (p < N) p++;
    if (mark == '>') {
      // body: everything up to the next '>' or '@' (anywhere), whitespace dropped
      p = (size_t)(copy_body(data + p, data + N, dst) - data);
    } else {
      // FASTQ: one sequence line, '+' line, one quality line
      while (p < N && is_space(data[p])) p++;
      while (p < N && data[p] != '\n') { if (!is_space(data[p])) *dst++ = data[p]; p++; }
      while (p < N && is_space(data[p])) p++;
      if (p < N && data[p] == '+') {
        while (p < N && data[p] != '\n') p++;
        while (p < N && is_space(data[p])) p++;
        while (p < N && data[p] != '\n') p++;
      }
    }
    out.off.push_back((uint64_t)(dst - base));
  }
  out.bases.resize((size_t)(dst - base));
}

bool load_sequences(const std::string &path, SeqSet &out, std::string &err) {
  // plain files are parsed straight from a read-only mapping; compressed ones from the decompressor's output
  struct stat st;
  if (!ends_with(path, ".gz") && !ends_with(path, ".bz2") && !ends_with(path, ".xz") &&
      stat(path.c_str(), &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {       // (pipes and devices are read once, below)
    const int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) { err = "cannot open '" + path + "'"; return false; }
    void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (m != MAP_FAILED) {
      madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
      parse_records((const char *)m, (size_t)st.st_size, out);
      munmap(m, (size_t)st.st_size);
      return true;
    }
  }
  std::vector<char> data;
  if (!slurp_file(path, data, err)) return false;
  parse_records(data.data(), data.size(), out);
  return true;
}
