"""Host-checkable pieces of the device index arithmetic.  The fused table maps h -> (line, tag) with
line = h / D, tag = h % D for D = d << t; the kernels replace the division by a shift and one 32-bit
multiply-high with the constants in csrc/joint.cu (joint.cuh: joint_split).  The constants are read
from the source and checked here against true division.  CPU only."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _constants():
    src = open(os.path.join(ROOT, "merfin_b200", "csrc", "joint.cu")).read()
    m = re.search(r"FS\[5\]\[3\]\s*=\s*\{(.*?)\};", src, re.S)
    rows = re.findall(r"\{\s*(\d+)\s*,\s*(0x[0-9A-Fa-f]+|0)u?\s*,\s*(\d+)\s*\}", m.group(1))
    assert len(rows) == 5
    return [(int(d), int(mm, 16), int(s)) for d, mm, s in rows]


def _div(x, d, m, s):
    """joint_split's quotient: d == 1 (power-of-two divisor) passes x through, else __umulhi(x, m) >> s."""
    if d == 1:
        return x
    return ((x * np.uint64(m)) >> np.uint64(32)) >> np.uint64(s)


def test_multiply_high_division_is_exact_for_every_32_bit_value():
    rng = np.random.default_rng(0)
    for d, m, s in _constants():
        assert d in (1, 3, 5, 9, 15)
        edge = np.array([0, 1, d - 1, d, d + 1, 2**31 - 1, 2**31, 2**32 - d - 1, 2**32 - d, 2**32 - 2, 2**32 - 1], dtype=np.uint64)
        mult = (np.arange(0, 2**32 // d, 9973, dtype=np.uint64) * np.uint64(d))            # multiples of d and their neighbours
        x = np.concatenate([edge, mult, mult[1:] - np.uint64(1), rng.integers(0, 2**32, size=4_000_000, dtype=np.uint64)])
        q = _div(x, d, m, s)                                                               # __umulhi(x, m) >> s; x*m < 2^64
        assert (q == x // np.uint64(d)).all(), d


def test_split_reassembles_to_h():
    """(line, tag) <-> h is one-to-one: h = line * D + tag with tag < D, for every geometry t >= 2k - 32."""
    rng = np.random.default_rng(1)
    for d, m, s in _constants():
        for k, t in ((21, 10), (21, 13), (16, 0), (16, 9), (31, 30), (32, 32), (32, 40)):
            bits = 2 * k
            assert t >= max(0, bits - 32)
            h = rng.integers(0, 2**63, size=200_000, dtype=np.uint64) >> np.uint64(63 - bits if bits < 63 else 0)
            if bits == 64:
                h = rng.integers(0, 2**64, size=200_000, dtype=np.uint64)
            x = h >> np.uint64(t)
            assert int(x.max()) < 2**32
            q = _div(x, d, m, s)
            rem = x - q * np.uint64(d)
            tag = (rem << np.uint64(t)) | (h & np.uint64((1 << t) - 1))
            D = d << t
            assert (tag < np.uint64(D)).all()
            hh = [int(a) * D + int(b) for a, b in zip(q[:2000], tag[:2000])]
            assert hh == [int(v) for v in h[:2000]]


def test_mix_is_a_bijection_and_unmix_inverts_it(tmp_path):
    """mfn_mix / mfn_unmix are taken verbatim from csrc/table.cuh (CUDA qualifiers dropped) and compiled
    for the host: exhaustive permutation check for k = 4..12, round trip on random k-mers for every k."""
    import subprocess
    src = open(os.path.join(ROOT, "merfin_b200", "csrc", "table.cuh")).read()
    fns = []
    for name in ("mfn_mix", "mfn_unmix"):
        m = re.search(r"__host__ __device__ __forceinline__ (uint64_t %s\(.*?\n\})" % name, src, re.S)
        assert m, name
        fns.append("static inline " + m.group(1))
    prog = r"""
#include <cstdint>
#include <cstdio>
#include <vector>
%s
int main() {
  for (uint32_t k = 4; k <= 12; k++) {
    const uint64_t n = 1ull << (2 * k), kmask = n - 1;
    for (uint64_t seed : {0ull, 0x9e3779b97f4a7c15ull, 0xffffffffffffffffull}) {
      std::vector<uint8_t> seen(n, 0);
      for (uint64_t c = 0; c < n; c++) {
        const uint64_t h = mfn_mix(c, seed, kmask, k);
        if (h >= n || seen[h]) { printf("not a permutation: k=%%u\n", k); return 1; }
        seen[h] = 1;
        if (mfn_unmix(h, seed, kmask, k) != c) { printf("unmix fails: k=%%u\n", k); return 1; }
      }
    }
  }
  uint64_t s = 88172645463325252ull;
  for (uint32_t k = 1; k <= 32; k++) {
    const uint64_t kmask = k == 32 ? ~0ull : (1ull << (2 * k)) - 1;
    for (int i = 0; i < 200000; i++) {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;
      const uint64_t c = s & kmask, seed = s * 0x2545F4914F6CDD1Dull;
      const uint64_t h = mfn_mix(c, seed, kmask, k);
      if ((h & ~kmask) || mfn_unmix(h, seed, kmask, k) != c) { printf("round trip fails: k=%%u\n", k); return 1; }
    }
  }
  printf("ok\n");
  return 0;
}
""" % "\n".join(fns)
    cpp = tmp_path / "mix.cpp"
    cpp.write_text(prog)
    exe = str(tmp_path / "mix")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, str(cpp)])
    p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and p.stdout.strip() == "ok", p.stdout


def test_magic_division_needs_at_most_one_correction():
    """table.cuh mfn_split: q = umul64hi(h, floor(2^64 / D)) (all ones for D = 1), then ONE conditional
    correction.  Checked with exact integers over random and edge (h, D)."""
    import random
    rnd = random.Random(5)
    M = 1 << 64
    Ds = [1, 2, 3, 5, 7, 255, 256, 257, 22369, (1 << 22) - 1, 1 << 22, (1 << 31) + 11, (1 << 40) - 87, (1 << 62) + 1, (1 << 63) - 1, 1 << 63]
    Ds += [rnd.randrange(1, 1 << rnd.randrange(1, 64)) for _ in range(300)]
    for D in Ds:
        magic = M - 1 if D == 1 else M // D
        hs = [0, 1, D - 1, D, D + 1, 2 * D - 1, 2 * D, M - 1, M - D, (M // D) * D - 1, (M // D) * D]
        hs += [rnd.randrange(0, M) for _ in range(200)]
        for h in hs:
            if not 0 <= h < M:
                continue
            q = (h * magic) >> 64
            r = h - q * D
            if r >= D:
                q, r = q + 1, r - D
            assert (q, r) == divmod(h, D), (h, D)


def test_hash_on_32_bit_halves_equals_mfn_mix(tmp_path):
    """The scatter pass hashes 16 < k < 32 on two 32-bit halves (sweep_kernels.cuh: Kmer2) with a literal 0
    seed; the struct is taken verbatim from the source and must equal mfn_mix(c, 0, kmask, k) for every k."""
    import subprocess
    tsrc = open(os.path.join(ROOT, "merfin_b200", "csrc", "table.cuh")).read()
    m = re.search(r"__host__ __device__ __forceinline__ (uint64_t mfn_mix\(.*?\n\})", tsrc, re.S)
    ssrc = open(os.path.join(ROOT, "merfin_b200", "csrc", "sweep_kernels.cuh")).read()
    k2 = re.search(r"(template <int K> struct Kmer2 \{.*?\n\};)", ssrc, re.S)
    assert m and k2
    body = k2.group(1).replace("__device__ __forceinline__", "inline")
    cases = "\n".join("  if (!check<%d>()) return 1;" % k for k in range(17, 32))
    prog = r"""
#include <cstdint>
#include <cstdio>
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) { return (uint32_t)(((((uint64_t)hi) << 32) | lo) >> (s & 31)); }
static inline %s
%s
template <int K> bool check() {
  const uint64_t kmask = (1ull << (2 * K)) - 1;
  uint64_t s = 88172645463325252ull + K;
  for (int i = 0; i < 300000; i++) {
    s ^= s << 13; s ^= s >> 7; s ^= s << 17;
    const uint64_t c = i < 4 ? (i == 0 ? 0 : i == 1 ? kmask : i == 2 ? 1 : kmask - 1) : (s & kmask);
    uint32_t hi = (uint32_t)(c >> 32), lo = (uint32_t)c;
    Kmer2<K>::mix(hi, lo);
    if (((((uint64_t)hi) << 32) | lo) != mfn_mix(c, 0, kmask, K)) { printf("k=%%d differs\n", K); return false; }
  }
  return true;
}
int main() {
%s
  printf("ok\n");
  return 0;
}
""" % (m.group(1), body, cases)
    cpp = tmp_path / "halves.cpp"
    cpp.write_text(prog)
    exe = str(tmp_path / "halves")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, str(cpp)])
    p = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and p.stdout.strip() == "ok", p.stdout
