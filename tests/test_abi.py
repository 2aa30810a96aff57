"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol the
header declares, fails loudly without a GPU, and the meryl writer is readable by the oracle."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from merfin_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "merfin_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mfn_[a-z_0-9A-Z]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    names = _declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/merfin_b200.h but not exported"
    # the ctypes stub covers the same set
    assert sorted(L.SIGNATURES) == names


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = L.load()
    h = C.c_void_p()
    st = lib.mfn_ctx_create(C.byref(h), 0)
    assert st == L.MFN_ERR_NO_DEVICE
    assert b"CUDA" in lib.mfn_last_error() or b"device" in lib.mfn_last_error()
    with pytest.raises(L.MfnError):
        L.check(st)


def test_product_does_not_reference_oracle():
    """The product path must not import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "merfin_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                text = open(os.path.join(dp, fn), errors="replace").read()
                assert "mfn_oracle" not in text and "oracle_lib" not in text and "oracle/" not in text, os.path.join(dp, fn)
    out = subprocess.run(["ldd", L.SO_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_host_replay_matches_oracle_kmodel():
    lib = L.load()
    K = np.array([0, 0, 1, 1, 1, 2], dtype=np.uint32)
    P = np.array([0.9, 0.8, 0.7, 0.75, 0.9, 0.5])
    km = L.KModel(7.5, len(K), K.ctypes.data_as(C.POINTER(C.c_uint32)), P.ctypes.data_as(C.POINTER(C.c_double)))
    okm = O.KModel(7.5, K, P)
    a, b, c = C.c_double(), C.c_double(), C.c_double()
    for vr in list(range(0, 64)) + [100, 1000, 4095, 65536, 2 ** 31, 2 ** 32 - 1]:
        for va in (0, 1, 2, 3, 7, 1000):
            lib.mfn_getK(C.byref(km), vr, va, C.byref(a), C.byref(b), C.byref(c))
            assert (a.value, b.value, c.value) == okm.getK(vr, va)
            assert lib.mfn_getKmetric(a.value, b.value) == O.L.o_getKmetric(a.value, b.value) or \
                (np.isnan(lib.mfn_getKmetric(a.value, b.value)) and np.isnan(O.L.o_getKmetric(a.value, b.value)))


def test_host_hist_report_matches_oracle():
    """mfn_hist_report / mfn_hist_format replay of the K* bins == the oracle's per-window accumulation."""
    from merfin_b200 import api
    rng = np.random.default_rng(5)
    vr = rng.integers(0, 90, size=4000).astype(np.uint32)
    va = rng.integers(0, 6, size=4000).astype(np.uint32)
    vr[:50] = 0
    K = np.array([0, 1, 1], dtype=np.uint32); P = np.array([0.5, 0.25, 0.75])
    for probK, probP in (((), ()), (K, P)):
        okm = O.KModel(9.0, probK, probP)
        h = O.Hist()
        for a, b in zip(vr, va):
            h.add_pair(okm, a, b)
        keys, counts = np.unique(np.stack([va, vr], 1), axis=0, return_counts=True)
        pairs = np.zeros(len(keys), dtype=api.PAIR_DTYPE)
        pairs["v_asm"], pairs["v_read"], pairs["count"] = keys[:, 0], keys[:, 1], counts
        res = api.HistResult(["x"], np.zeros(1, dtype=api.COUNTS_DTYPE), pairs, 21, api.KModel(9.0, probK, probP))
        assert res.file_text() == h.report()
        s, _, _ = res.report()
        assert (s.kasm, s.kmissing) == (h.kasm, h.kmissing)
        assert abs(s.koverCpy - h.koverCpy) < 1e-9
        assert res.summary_text() == h.summary(21)


def test_meryl_writer_roundtrip(tmp_path):
    """mfn_meryl_write output is decoded identically by the oracle reader (and, where the
    reference build exists, by the reference's own `meryl print`)."""
    from merfin_b200 import api
    rng = np.random.default_rng(3)
    for k, n in ((21, 5000), (6, 300), (31, 2000), (16, 70000)):
        kmers = np.unique(rng.integers(0, 1 << min(62, 2 * k), size=n, dtype=np.uint64))
        values = rng.integers(1, 300, size=len(kmers)).astype(np.uint32)
        values[:3] = [1, 70000, 4000000000]
        d = str(tmp_path / f"w{k}.meryl")
        api.meryl_write(d, k, kmers, values)
        db = O.DB(d)
        assert db.k == k and db.n == len(kmers)
        assert (db.lo == kmers).all() and (db.hi == 0).all() and (db.val == values).all()
        vals, cnts = np.unique(values, return_counts=True)
        assert (db.histValue == vals).all() and (db.histCount == cnts).all()
        meryl = os.path.join(ROOT, "oracle", "_ref", "meryl")
        if os.path.exists(meryl):
            p = subprocess.run([meryl, "threads=1", "print", d], capture_output=True, text=True)
            assert p.returncode == 0, p.stderr
            lines = [l.split("\t") for l in p.stdout.splitlines()]
            assert len(lines) == len(kmers)
            assert [O.kmer_from_string(a) for a, _ in lines] == [int(x) for x in kmers]
            assert [int(b) for _, b in lines] == [int(x) for x in values]


@pytest.mark.parametrize("name,k", [("g1_k21", 21), ("g2_k6", 6), ("g3_k31", 31)])
def test_meryl_info_matches_the_reference_listing(golden_dir, name, k):
    """mfn_meryl_info / mfn_meryl_histogram (host only) against the reference's `meryl print` of the
    same databases (committed as *.print.gz): k, distinct / unique / total and the value histogram."""
    import gzip
    from merfin_b200 import api
    for db in ("reads", "asm"):
        d = os.path.join(golden_dir, name, db + ".meryl")
        vals = np.array([int(l.split()[1]) for l in gzip.open(os.path.join(golden_dir, name, db + ".print.gz"), "rt")], dtype=np.uint64)
        info = api.meryl_info(d)
        assert info["k"] == k and info["prefix_bits"] + info["suffix_bits"] == 2 * k
        assert info["version"] == 3 and info["files"] == 64 and info["multiset"] == 0
        assert info["n_distinct"] == len(vals)
        assert info["n_unique"] == int((vals == 1).sum())
        assert info["n_total"] == int(vals.sum())
        hv, hc = api.meryl_histogram(d)
        uv, uc = np.unique(vals, return_counts=True)
        assert info["hist_len"] == len(uv)
        assert (hv == uv).all() and (hc == uc.astype(np.uint64)).all()


@pytest.mark.parametrize("name,k", [("g1_k21", 21), ("g2_k6", 6), ("g3_k31", 31), ("g4_k40", 40), ("g5_k64", 64)])
def test_meryl_read_matches_meryl_print(golden_dir, name, k):
    """mfn_meryl_read (host decoder, any k <= 64: 128-bit k-mers as two halves) against the reference's
    `meryl print` of the same databases, k-mer by k-mer in stored order."""
    import gzip
    from merfin_b200 import api
    for db in ("reads", "asm"):
        lines = [l.split() for l in gzip.open(os.path.join(golden_dir, name, db + ".print.gz"), "rt")]
        lo, hi, v = api.meryl_read(os.path.join(golden_dir, name, db + ".meryl"))
        assert api.meryl_info(os.path.join(golden_dir, name, db + ".meryl"))["k"] == k
        assert len(lines) == len(lo)
        assert [O.kmer_from_string(a) for a, _ in lines] == [(int(h) << 64) | int(l) for h, l in zip(hi, lo)]
        assert [int(b) for _, b in lines] == [int(x) for x in v]
        if k <= 32:
            assert not hi.any()


def test_meryl_info_errors(tmp_path):
    info = L.MerylInfo()
    assert L.load().mfn_meryl_info(str(tmp_path).encode(), C.byref(info)) == L.MFN_ERR_IO
    assert b"doesn't appear to be a meryl database" in L.load().mfn_last_error()
    assert L.load().mfn_meryl_info(None, C.byref(info)) == L.MFN_ERR_INVALID


def test_header_is_plain_c_and_example_links(tmp_path):
    """include/merfin_b200.h is a C header (no C++): a C99 -pedantic translation unit that uses
    the -hist entry points compiles and links against the library (examples/hist_example.c); run
    without a GPU it must fail with the library's own 'no device' error, not crash."""
    import subprocess
    exe = str(tmp_path / "hist_example")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "hist_example.c"), "-L" + os.path.join(ROOT, "merfin_b200"),
                           "-lmerfin_b200", "-Wl,-rpath," + os.path.join(ROOT, "merfin_b200"), "-o", exe])
    import torch
    if not torch.cuda.is_available():
        g = os.path.join(ROOT, "tests", "golden", "g1_k21")
        p = subprocess.run([exe, os.path.join(g, "reads.meryl"), os.path.join(g, "asm.meryl"), "8", "ACGT" * 10],
                           capture_output=True, text=True)
        assert p.returncode == 1 and "no usable CUDA device" in p.stderr


@pytest.mark.parametrize("threads", [1, 2, 3, 7, 64])
def test_vcf_loader_edge_cases(tmp_path, monkeypatch, threads):
    """mfn_vcf_load / _info / _header / _cluster (host only; vcf.C:93-246) on a file with CRLF and blank lines, runs of
    tabs, header lines in the middle, records with fewer than ten columns, a last line without a newline -- the same
    counts whatever number of parsing threads the byte ranges are cut for."""
    from merfin_b200 import api
    recs = []
    for i in range(400):
        chrom = "chrA" if i % 3 else "chrB"
        recs.append("%s\t%d\t.\tA\tC,G\t%d\tPASS\t.\tGT:DP\t%s:7" % (chrom, 10 + 37 * i, 20 + i % 30, ("1/1", "0/1", "1/2", "0/0", "./.")[i % 5]))
    lines = ["##fileformat=VCFv4.2", "##contig=<ID=chrA,length=100000>", "##contig=<ID=chrB,length=100000>",
             "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS"]
    lines += recs[:100] + ["", "##late=header"] + recs[100:200] + ["chrA\t5\t.\tA\tC\t1\tPASS\t.\tGT"]          # 9 columns: excluded
    lines += [recs[200].replace("\t", "\t\t", 3)] + [l + "\r" for l in recs[201:300]] + ["\r", "chrB\t7"] + recs[300:]
    p = tmp_path / "odd.vcf"
    p.write_bytes("\n".join(lines).encode())                      # no newline after the last record
    monkeypatch.setenv("MFN_VCF_THREADS", str(threads))
    monkeypatch.setenv("MFN_QUIET", "1")
    v = api.Vcf(str(p))
    i = v.info
    assert (i.headers, i.records, i.excluded, i.contigs, i.contig_ids) == (5, 400, 2, 2, 2)
    assert v.header() == ("\n".join(lines[:4] + ["##late=header"]) + "\n").encode()
    assert v.clusters_of("chrA") + v.clusters_of("chrB") == 400 and v.clusters_of("chrC") == 0
    v.cluster(21)
    # chrA's records come in pairs 37 bp apart (closer than 2k = 42: merged), chrB's are 111 bp apart (left alone)
    assert v.clusters_of("chrA") == 133 and v.clusters_of("chrB") == 134
    counts = (v.clusters_of("chrA"), v.clusters_of("chrB"))
    v.close()
    monkeypatch.setenv("MFN_VCF_THREADS", "1")
    v1 = api.Vcf(str(p))
    v1.cluster(21)
    assert (v1.clusters_of("chrA"), v1.clusters_of("chrB")) == counts
    v1.close()
