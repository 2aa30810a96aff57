// ASan/UBSan harness for the host packer (test infrastructure): exact-size heap buffers, many record
// layouts and cuts; the packed words are also checked against a per-base loop.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
extern "C" void mfn_pack_host_range(const char *bases, const uint64_t *rec_start, uint32_t nseq, uint64_t g_begin,
                                    uint64_t g_end, uint64_t *seq2_out, uint64_t *ok_out);
int main() {
  uint64_t s = 99;
  auto rnd = [&]() { s = s * 6364136223846793005ull + 1442695040888963407ull; return s >> 33; };
  const uint32_t pool[] = {0, 1, 2, 31, 32, 33, 63, 64, 65, 127, 128, 129, 1000, 4097, 70000};
  for (int iter = 0; iter < 300; iter++) {
    const uint32_t nseq = 1 + rnd() % 40;
    std::vector<uint64_t> rs(nseq + 1);
    uint64_t total = 0;
    std::vector<uint32_t> len(nseq);
    for (uint32_t r = 0; r < nseq; r++) { len[r] = pool[rnd() % 15]; rs[r] = total + r; total += len[r]; }
    rs[nseq] = total + nseq;
    char *bases = (char *)malloc(total ? total : 1);
    // mostly the usual alphabet; one byte in sixteen is any value at all (0xdf and 0xff are not bases)
    for (uint64_t i = 0; i < total; i++) { const uint64_t x = rnd(); bases[i] = (x & 15) == 0 ? (char)(x >> 8) : "ACGTacgtNn"[(x >> 8) % 10]; }
    const uint64_t G = (rs[nseq] + 2047) / 2048 * 2048;
    // expected
    std::vector<uint64_t> w2(G / 32, 0), wo(G / 64, 0);
    for (uint32_t r = 0; r < nseq; r++)
      for (uint32_t i = 0; i < len[r]; i++) {
        const unsigned char c = (unsigned char)bases[rs[r] - r + i], u = c | 0x20;
        if (u == 'a' || u == 'c' || u == 'g' || u == 't') {
          const uint64_t g = rs[r] + i;
          w2[g / 32] |= (uint64_t)((c >> 1) & 3) << (2 * (g % 32));
          wo[g / 64] |= 1ull << (g % 64);
        }
      }
    // packed in 64-aligned pieces, each into its own exact-size allocation
    uint64_t a = 0;
    while (a < G) {
      uint64_t b = a + 64 * (1 + rnd() % 600);
      if (b > G) b = G;
      uint64_t *o2 = (uint64_t *)malloc((b - a) / 32 * 8), *oo = (uint64_t *)malloc((b - a) / 64 * 8);
      memset(o2, 0xAB, (b - a) / 32 * 8); memset(oo, 0xCD, (b - a) / 64 * 8);
      mfn_pack_host_range(bases, rs.data(), nseq, a, b, o2, oo);
      if (memcmp(o2, w2.data() + a / 32, (b - a) / 32 * 8) || memcmp(oo, wo.data() + a / 64, (b - a) / 64 * 8)) { printf("mismatch iter %d [%lu,%lu)\n", iter, (unsigned long)a, (unsigned long)b); return 1; }
      free(o2); free(oo);
      a = b;
    }
    free(bases);
  }
  printf("ok\n");
  return 0;
}
