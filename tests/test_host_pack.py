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


def test_scalar_and_avx2_kernels_agree(tmp_path):
    """The packer picks its 32-base kernel once per process; run the same input through a child
    process with the scalar kernel forced and compare the packed words."""
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
rs = np.array([0, 70_001, 70_001 + 130_002 + 1], dtype=np.uint64)
G = (int(rs[-1]) + 2047) // 2048 * 2048
s2 = np.zeros(G // 32, dtype=np.uint64); ok = np.zeros(G // 64, dtype=np.uint64)
lib.mfn_pack_host_range(C.c_void_p(b.ctypes.data), C.c_void_p(rs.ctypes.data), 2, C.c_uint64(0), C.c_uint64(G),
                        C.c_void_p(s2.ctypes.data), C.c_void_p(ok.ctypes.data))
sys.stdout.write(lib.mfn_pack_host_isa().decode() + " " + str(int(s2.sum() & np.uint64(0xffffffffffff))) + " " + str(int(np.bitwise_xor.reduce(s2))) + " " + str(int(np.bitwise_xor.reduce(ok))))
""" % root
    outs = {}
    for scalar in ("0", "1"):
        env = dict(os.environ, MFN_PACK_SCALAR=scalar)
        p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
        assert p.returncode == 0, p.stderr
        isa, *digest = p.stdout.split()
        outs[scalar] = (isa, digest)
    assert outs["1"][0] == "scalar"
    assert outs["0"][1] == outs["1"][1]
