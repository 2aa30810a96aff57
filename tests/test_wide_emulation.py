"""The k = 33..64 path (merfin_b200/csrc/wide.cu) without a GPU: the table lookup and the window arithmetic of its
kernels are __host__ __device__ functions in csrc/wide.cuh, and tests/wide_emulation_harness.cu runs exactly those on the
CPU over the packed arrays, in the kernel's tile / thread / window order.  Here their results are compared with the
oracle on the k = 40 and k = 64 golden sets (whose oracle results are pinned to the reference's output files in
test_oracle_golden.py).  The GPU tests (test_gpu_wide.py) then only have the launch itself left to prove."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "wide_emu")


@pytest.fixture(scope="module")
def emu():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    so_dir = os.path.join(ROOT, "merfin_b200")
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-O2", "-std=c++17", "-arch=sm_100a", "-o", EXE,
                           os.path.join(ROOT, "tests", "wide_emulation_harness.cu"),
                           os.path.join(ROOT, "merfin_b200", "csrc", "host", "fasta.cpp"),
                           "-L" + so_dir, "-lmerfin_b200", "-Xlinker", "-rpath," + so_dir])
    return EXE


@pytest.mark.parametrize("name,k", [("g4_k40", 40), ("g5_k64", 64)])
@pytest.mark.parametrize("minmax", [(0, 0xFFFFFFFF), (2, 12)])
def test_pairs_of_every_position(emu, golden_dir, tmp_path, name, k, minmax):
    d = os.path.join(golden_dir, name)
    out = str(tmp_path / "emu")
    subprocess.check_call([emu, os.path.join(d, "asm.fa"), os.path.join(d, "reads.meryl"), os.path.join(d, "asm.meryl"),
                           str(minmax[0]), str(minmax[1]), out])
    vr, va, ok = np.fromfile(out + ".vr", dtype=np.uint32), np.fromfile(out + ".va", dtype=np.uint32), np.fromfile(out + ".ok", dtype=np.uint8)
    R = O.Lookup(O.DB(os.path.join(d, "reads.meryl")), *minmax)
    A = O.Lookup(O.DB(os.path.join(d, "asm.meryl")))
    at = 0
    nvalid = 0
    for _, bases in O.read_fasta(os.path.join(d, "asm.fa")):
        wr, wa, wok = O.pairs(R, A, bases)
        n = len(wr)
        assert (ok[at:at + n] == wok).all()
        assert (vr[at:at + n] == wr * wok).all() and (va[at:at + n] == wa * wok).all()
        assert not ok[at + n:at + len(bases)].any()          # the last k - 1 positions start no window
        nvalid += int(wok.sum())
        at += len(bases)
    assert at == len(vr) and nvalid > 1000


@pytest.mark.parametrize("name,k", [("g4_k40", 40), ("g5_k64", 64)])
def test_window_kmers_are_the_database(emu, golden_dir, tmp_path, name, k):
    """The canonical k-mers cut out of the packed assembly (k_wide_windows' arithmetic), sorted and run-length
    counted, are the reference's `meryl count` of the same file."""
    import gzip
    d = os.path.join(golden_dir, name)
    out = str(tmp_path / "cnt")
    subprocess.check_call([emu, "--count", os.path.join(d, "asm.fa"), str(k), out])
    hi, lo = np.fromfile(out + ".hi", dtype=np.uint64), np.fromfile(out + ".lo", dtype=np.uint64)
    got = {}
    for h, l in zip(hi, lo):
        key = (int(h) << 64) | int(l)
        got[key] = got.get(key, 0) + 1
    want = {O.kmer_from_string(a): int(b) for a, b in (l.split() for l in gzip.open(os.path.join(d, "asm.print.gz"), "rt"))}
    assert got == want
