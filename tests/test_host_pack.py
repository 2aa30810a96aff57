"""Host-side SIMD packer (csrc/host_pack.cpp): ASCII -> 2-bit words + validity bits in global
coordinates, bit-identical to the definition k_pack implements (kmers-tiny.H:69-83,
kmers-iterator.H:97-102).  CPU only: the packer is plain C++ inside the C-ABI library."""
import ctypes as C

import numpy as np
import pytest


def _expected(recs, rec_start, Gp):
    seq2 = np.zeros(Gp // 32, dtype=np.uint64)
    ok = np.zeros(Gp // 64, dtype=np.uint64)
    for r, b in enumerate(recs):
        g = int(rec_start[r]) + np.arange(len(b), dtype=np.uint64)
        lower = b | 0x20
        valid = (lower == ord("a")) | (lower == ord("c")) | (lower == ord("g")) | (lower == ord("t"))
        code = ((b >> 1) & 3).astype(np.uint64)
        gv, cv = g[valid], code[valid]
        np.bitwise_or.at(seq2, (gv // 32).astype(np.int64), cv << (2 * (gv % 32)))
        np.bitwise_or.at(ok, (gv // 64).astype(np.int64), np.uint64(1) << (gv % 64))
    return seq2, ok


@pytest.mark.parametrize("seed,lens", [(1, [1000, 37, 64, 1, 5000, 31, 33, 100000]), (2, [0, 5, 0, 63, 65, 4096]), (3, [300000])])
def test_pack_host_range_matches_definition(seed, lens):
    from merfin_b200 import lib as L
    lib = L.load()
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACGTacgtNnXRY-", dtype=np.uint8)
    p = [.2, .2, .2, .2, .04, .04, .04, .04, .01, .01, .005, .005, .005, .005]
    recs = [rng.choice(alphabet, size=n, p=p) for n in lens]
    bases = np.concatenate(recs) if sum(lens) else np.zeros(1, dtype=np.uint8)
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    rec_start = (off + np.arange(len(lens) + 1, dtype=np.uint64)).astype(np.uint64)
    Gp = (int(rec_start[-1]) + 2047) // 2048 * 2048
    want2, wantok = _expected(recs, rec_start, Gp)
    # pack in several 64-aligned ranges, as the worker threads do
    cuts = sorted({0, Gp} | {int(x) // 64 * 64 for x in rng.integers(0, Gp, size=5)})
    got2 = np.full(Gp // 32, 0xdeadbeef, dtype=np.uint64)
    gotok = np.full(Gp // 64, 0xdeadbeef, dtype=np.uint64)
    for a, b in zip(cuts[:-1], cuts[1:]):
        lib.mfn_pack_host_range(C.c_void_p(bases.ctypes.data), C.c_void_p(rec_start.ctypes.data), len(lens),
                                C.c_uint64(a), C.c_uint64(b), C.c_void_p(got2.ctypes.data + a // 32 * 8),
                                C.c_void_p(gotok.ctypes.data + a // 64 * 8))
    assert (got2 == want2).all()
    assert (gotok == wantok).all()


def test_scalar_avx2_and_avx512_kernels_agree(tmp_path):
    """The packer picks its inner kernel once per process; run the same input through child processes with each
    kernel forced (MFN_PACK_ISA) and compare the packed words.  The input holds EVERY byte value (the AVX2 kernel once
    took 0xdf and 0xff for bases: its "matches nothing" table entry was 0xff) next to the usual alphabet."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import ctypes as C, sys, numpy as np
sys.path.insert(0, %r)
from merfin_b200 import lib as L
lib = L.load()
rng = np.random.default_rng(21)
b = rng.choice(np.frombuffer(b"ACGTacgtNn-*", dtype=np.uint8), size=200_003)
b[1000:1256] = np.arange(256, dtype=np.uint8)
b[150_000:150_000 + 4096] = rng.integers(0, 256, size=4096, dtype=np.uint8)
rs = np.array([0, 70_001, 70_001 + 130_002 + 1], dtype=np.uint64)
G = (int(rs[-1]) + 2047) // 2048 * 2048
s2 = np.zeros(G // 32, dtype=np.uint64); ok = np.zeros(G // 64, dtype=np.uint64)
lib.mfn_pack_host_range(C.c_void_p(b.ctypes.data), C.c_void_p(rs.ctypes.data), 2, C.c_uint64(0), C.c_uint64(G),
                        C.c_void_p(s2.ctypes.data), C.c_void_p(ok.ctypes.data))
valid = [int(b[i]) for i in range(1000, 1256) if (int(ok[i // 64]) >> (i %% 64)) & 1]
sys.stdout.write(lib.mfn_pack_host_isa().decode() + " " + str(int(s2.sum() & np.uint64(0xffffffffffff))) + " " + str(int(np.bitwise_xor.reduce(s2))) + " " + str(int(np.bitwise_xor.reduce(ok))) + " " + ",".join(map(str, valid)))
""" % root
    outs = {}
    for isa in ("scalar", "avx2", "avx512"):
        env = dict(os.environ, MFN_PACK_ISA=isa)
        p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
        assert p.returncode == 0, p.stderr
        got, *digest = p.stdout.split()
        outs[isa] = (got, digest)
    assert outs["scalar"][0] == "scalar"
    assert outs["scalar"][1][3] == ",".join(str(c) for c in sorted(b"ACGTacgt"))          # exactly the eight letters are bases
    assert outs["avx2"][1] == outs["scalar"][1] and outs["avx512"][1] == outs["scalar"][1]
    import subprocess as sp
    flags = open("/proc/cpuinfo").read()
    if " avx2" in flags:
        assert outs["avx2"][0] == "avx2"
    if " avx512bw" in flags:
        assert outs["avx512"][0] == "avx512"                      # (else the request falls back to what the CPU has)


@pytest.mark.parametrize("seed", range(12))
def test_pack_host_range_block_boundaries(seed):
    """No up-front memset: whole 64-position blocks are stored outright, the rest is zeroed and OR-ed.
    Record lengths around the block size, runs of empty / one-base records sharing a block, random
    64-aligned cuts, output arrays pre-filled with garbage."""
    from merfin_b200 import lib as L
    lib = L.load()
    rng = np.random.default_rng(100 + seed)
    pool = [0, 0, 1, 1, 2, 5, 30, 31, 32, 33, 61, 62, 63, 64, 65, 66, 126, 127, 128, 129, 191, 192, 193, 255, 256, 257, 1000, 4096, 4097]
    lens = [int(rng.choice(pool)) for _ in range(int(rng.integers(1, 120)))]
    if seed % 3 == 0:
        lens.insert(int(rng.integers(0, len(lens))), 70_000)
    alphabet = np.frombuffer(b"ACGTacgtN", dtype=np.uint8)
    recs = [rng.choice(alphabet, size=n) for n in lens]
    bases = np.concatenate(recs + [np.zeros(1, dtype=np.uint8)])
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    rec_start = (off + np.arange(len(lens) + 1, dtype=np.uint64)).astype(np.uint64)
    Gp = (int(rec_start[-1]) + 2047) // 2048 * 2048 + (2048 if seed % 2 else 0)      # sometimes a tail of pure padding
    want2, wantok = _expected(recs, rec_start, Gp)
    cuts = sorted({0, Gp} | {int(x) // 64 * 64 for x in rng.integers(0, Gp, size=int(rng.integers(0, 12)))})
    got2 = np.full(Gp // 32, 0xdeadbeefdeadbeef, dtype=np.uint64)
    gotok = np.full(Gp // 64, 0xdeadbeefdeadbeef, dtype=np.uint64)
    order = list(zip(cuts[:-1], cuts[1:]))
    rng.shuffle(order)                                  # ranges are independent: any order
    for a, b in order:
        lib.mfn_pack_host_range(C.c_void_p(bases.ctypes.data), C.c_void_p(rec_start.ctypes.data), len(lens),
                                C.c_uint64(a), C.c_uint64(b), C.c_void_p(got2.ctypes.data + a // 32 * 8),
                                C.c_void_p(gotok.ctypes.data + a // 64 * 8))
    assert (got2 == want2).all()
    assert (gotok == wantok).all()


def test_packer_under_address_sanitizer(tmp_path):
    """host_pack.cpp built with ASan+UBSan and driven by tests/pack_sanitize_harness.cpp: every piece is
    packed into its own exact-size allocation from an exact-size source (no access outside either), and
    compared with a per-base loop; the scalar, the AVX2 and (where the CPU has it) the AVX-512 kernel."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "pk_asan")
    b = subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                        os.path.join(root, "merfin_b200", "csrc", "host_pack.cpp"), os.path.join(root, "tests", "pack_sanitize_harness.cpp"),
                        "-o", exe], capture_output=True, text=True)
    if b.returncode != 0 and ("cannot find -l" in b.stderr or "libasan" in b.stderr):
        pytest.skip("sanitizer runtime not installed")
    assert b.returncode == 0, b.stderr[-2000:]
    for isa in ("avx2", "scalar", "avx512"):
        p = subprocess.run([exe], capture_output=True, text=True, timeout=600, env=dict(os.environ, MFN_PACK_ISA=isa))
        assert p.returncode == 0 and p.stdout.strip() == "ok", (isa, p.stdout[-500:], p.stderr[-3000:])
