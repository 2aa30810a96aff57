// Test harness for the k = 33..64 path without a GPU: merfin_b200/csrc/wide.cuh holds the table lookup and the window
// arithmetic of k_scan_wide / k_wide_windows as __host__ __device__ functions; this program runs exactly those
// functions on the CPU, tile by tile, thread slot by thread slot, window by window, the way the kernel walks the packed
// arrays, and writes the (v_read, v_asm, valid) triple of every base position.  tests/test_wide_emulation.py compares
// them with the oracle's pairs.  Test infrastructure only: nothing in the product calls this.
//   wide_emu <asm.fa> <reads.meryl> <asm.meryl> <minV> <maxV> <out prefix>      -> <out>.vr / .va / .ok (raw arrays)
//   wide_emu --count <asm.fa> <k> <out prefix>       canonical k-mers of all windows as k_wide_windows cuts them -> <out>.hi / .lo
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../merfin_b200/csrc/wide.cuh"
#include "../merfin_b200/csrc/host/fasta.h"

struct HostTable {
  std::vector<unsigned long long> hi, lo, index;
  std::vector<uint32_t> val;
  DevWide dev{};
};

static bool load_table(const char *dir, uint32_t minV, uint32_t maxV, HostTable &T) {
  mfn_meryl_info_t mi;
  if (mfn_meryl_info(dir, &mi) != MFN_OK) { fprintf(stderr, "%s\n", mfn_last_error()); return false; }
  uint64_t n = 0;
  if (mfn_meryl_read(dir, nullptr, nullptr, nullptr, 0, &n) != MFN_OK) return false;
  std::vector<uint64_t> lo(n), hi(n);
  std::vector<uint32_t> v(n);
  if (mfn_meryl_read(dir, lo.data(), hi.data(), v.data(), n, &n) != MFN_OK) return false;
  if (minV == 0) minV = 1;
  if (maxV == 0xffffffffu) { maxV = 0; for (uint32_t x : v) maxV = std::max(maxV, x); }
  for (uint64_t i = 0; i < n; i++)
    if (v[i] >= minV && v[i] <= maxV) { T.hi.push_back(hi[i]); T.lo.push_back(lo[i]); T.val.push_back(v[i] - (minV - 1)); }
  T.dev.hi = T.hi.data(); T.dev.lo = T.lo.data(); T.dev.val = T.val.data();
  T.dev.n = T.hi.size(); T.dev.k = mi.k; T.dev.ibits = wide_index_bits(T.dev.n);
  T.index.assign((1ull << T.dev.ibits) + 1, ~0ull);
  for (uint64_t i = 0; i <= T.dev.n; i++) wide_index_fill(T.dev, T.index.data(), i);
  for (unsigned long long x : T.index) if (x == ~0ull) { fprintf(stderr, "index entry never written\n"); return false; }
  T.dev.index = T.index.data();
  return true;
}

static void put(const std::string &path, const void *p, size_t bytes) {
  FILE *F = fopen(path.c_str(), "wb");
  if (!F || fwrite(p, 1, bytes, F) != bytes) { fprintf(stderr, "cannot write %s\n", path.c_str()); exit(1); }
  fclose(F);
}

int main(int argc, char **argv) {
  const bool count = argc > 1 && !strcmp(argv[1], "--count");
  if ((count && argc < 5) || (!count && argc < 7)) return 2;
  SeqSet S; std::string err;
  if (!load_sequences(argv[count ? 2 : 1], S, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
  const uint32_t nseq = (uint32_t)S.nseq();
  std::vector<uint64_t> rec_start(nseq + 1);
  for (uint32_t i = 0; i <= nseq; i++) rec_start[i] = S.off[i] + i;                  // one separator per record (seq.cu)
  const uint64_t G = S.off[nseq] + nseq;
  const uint64_t ntiles = std::max<uint64_t>(1, (G + MFN_TILE - 1) / MFN_TILE);
  const uint64_t gpad = (ntiles + 1) * MFN_TILE + 256;
  std::vector<unsigned long long> seq2(gpad / 32 + 8, 0), ok(gpad / 64 + 8, 0);
  S.bases.push_back(0);
  mfn_pack_host_range(S.bases.data(), rec_start.data(), nseq, 0, ntiles * MFN_TILE, (uint64_t *)seq2.data(), (uint64_t *)ok.data());

  if (count) {
    // k_wide_windows, one window start at a time
    const uint32_t k = (uint32_t)atoi(argv[3]);
    const uint64_t hmask = wide_hmask(k), kb = wide_kbits(k);
    std::vector<unsigned long long> oh, ol;
    for (uint64_t g = 0; g < G; g++) {
      const uint64_t wo = g >> 6, so = g & 63;
      const uint64_t vb = so ? ((ok[wo] >> so) | (ok[wo + 1] << (64 - so))) : ok[wo];
      if ((vb & kb) != kb) continue;
      const uint64_t w = g >> 5, sh = (g & 31) * 2;
      const uint64_t a = seq2[w], b = seq2[w + 1], c = seq2[w + 2];
      const uint64_t xl = sh ? ((a >> sh) | (b << (64 - sh))) : a;
      const uint64_t xh = (sh ? ((b >> sh) | (c << (64 - sh))) : b) & hmask;
      uint64_t fh, fl, rh, rl;
      wide_kmers(k, hmask, xh, xl, fh, fl, rh, rl);
      const bool fwd = fh < rh || (fh == rh && fl <= rl);
      oh.push_back(fwd ? fh : rh); ol.push_back(fwd ? fl : rl);
    }
    put(std::string(argv[4]) + ".hi", oh.data(), oh.size() * 8);
    put(std::string(argv[4]) + ".lo", ol.data(), ol.size() * 8);
    return 0;
  }

  HostTable R, A;
  if (!load_table(argv[2], (uint32_t)strtoul(argv[4], nullptr, 10), (uint32_t)strtoul(argv[5], nullptr, 10), R)) return 1;
  if (!load_table(argv[3], 0, 0xffffffffu, A)) return 1;
  if (R.dev.k != A.dev.k || R.dev.k < 33) { fprintf(stderr, "k mismatch or k <= 32\n"); return 1; }
  const uint32_t k = R.dev.k;
  const uint64_t hmask = wide_hmask(k), kb = wide_kbits(k);
  const uint64_t total = S.off[nseq];
  std::vector<uint32_t> vr(total + 1, 0), va(total + 1, 0);
  std::vector<uint8_t> okb(total + 1, 0);
  for (uint64_t tile = 0; tile < ntiles; tile++) {
    const unsigned long long *tseq = seq2.data() + tile * (MFN_TILE / 32), *tok = ok.data() + tile * (MFN_TILE / 64);   // what the TMA copies
    uint32_t r0;
    {
      const uint64_t gt = tile * MFN_TILE;
      uint32_t a = 0, b = nseq;
      while (a < b) { uint32_t m = (a + b + 1) >> 1; if (rec_start[m] <= gt) a = m; else b = m - 1; }
      r0 = a;
    }
    for (uint32_t tid = 0; tid < 256; tid++) {
      const uint32_t j0 = tid * 8;
      const uint64_t g0 = tile * MFN_TILE + j0;
      const WideSlice Sl = wide_slice(tseq, tok, j0);
      uint32_t rec = r0;
      uint64_t next_start = rec < nseq ? rec_start[rec + 1] : ~0ull;
      for (int uu = 0; uu < 8; uu++) {
        uint64_t xh, xl;
        const bool valid = wide_window(Sl, uu, hmask, kb, xh, xl);
        uint32_t r = 0, a = 0;
        if (valid) wide_pair_of(R.dev, A.dev, k, hmask, xh, xl, r, a);
        const uint64_t g = g0 + uu;
        while (g >= next_start) { rec++; next_start = rec < nseq ? rec_start[rec + 1] : ~0ull; }
        if (rec < nseq && g + 1 < next_start) {
          const uint64_t s = g - rec;
          vr[s] = r; va[s] = a; okb[s] = valid ? 1 : 0;
        }
      }
    }
  }
  put(std::string(argv[6]) + ".vr", vr.data(), total * 4);
  put(std::string(argv[6]) + ".va", va.data(), total * 4);
  put(std::string(argv[6]) + ".ok", okb.data(), total);
  return 0;
}
