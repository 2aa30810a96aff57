"""mfn_dump_write: the multi-threaded -dump formatter must produce exactly the text the
reference's outputDump does (merfin-dump.C:87-93); checked here against a straightforward
per-line formatting with the same getK / getKmetric calls.  CPU only."""
import ctypes as C
import os

import numpy as np


def test_dump_write_matches_line_by_line_formatting(tmp_path):
    from merfin_b200 import api, lib as L
    lib = L.load()
    rng = np.random.default_rng(5)
    n = 2_500_000                                     # several formatter chunks
    vr = rng.choice(np.array([0, 1, 2, 7, 8, 15, 16, 31, 100, 70000], dtype=np.uint32), size=n)
    va = rng.choice(np.array([0, 1, 2, 3, 9], dtype=np.uint32), size=n)
    ok = (rng.random(n) > 0.05).astype(np.uint8)
    km = api.KModel(8.0, [0, 0, 1, 1, 2], [0.5, 0.6, 0.7, 0.8, 0.9])
    out = str(tmp_path / "dump.txt")
    fd = os.open(out, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
    cnt = np.zeros(1, dtype=api.COUNTS_DTYPE)
    L.check(lib.mfn_dump_write(fd, b"scaffold_7", n, C.c_void_p(vr.ctypes.data), C.c_void_p(va.ctypes.data),
                               C.c_void_p(ok.ctypes.data), C.byref(km.c), 3, C.c_void_p(cnt.ctypes.data)))
    os.close(fd)
    want, per = api.dump_text(["scaffold_7"], np.array([0, n], dtype=np.uint64), vr, va, ok, km, 21)
    assert open(out, "rb").read() == want
    kmiss, _, kasm = per.strip().split("\t")[1:]
    assert (int(cnt["kasm"][0]), int(cnt["kmissing"][0])) == (int(kasm), int(kmiss))
    # count-only mode (-skipMissing): nothing written, same counters
    cnt2 = np.zeros(1, dtype=api.COUNTS_DTYPE)
    L.check(lib.mfn_dump_write(-1, b"scaffold_7", n, C.c_void_p(vr.ctypes.data), C.c_void_p(va.ctypes.data),
                               C.c_void_p(ok.ctypes.data), C.byref(km.c), 0, C.c_void_p(cnt2.ctypes.data)))
    assert cnt2[0] == cnt[0]


def _dump(lib, L, api, ident, vr, va, ok, km, threads, path):
    fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
    cnt = np.zeros(1, dtype=api.COUNTS_DTYPE)
    try:
        L.check(lib.mfn_dump_write(fd, ident.encode(), len(vr), C.c_void_p(vr.ctypes.data), C.c_void_p(va.ctypes.data),
                                   C.c_void_p(ok.ctypes.data), C.byref(km.c), threads, C.c_void_p(cnt.ctypes.data)))
    finally:
        os.close(fd)
    return open(path, "rb").read(), (int(cnt["kasm"][0]), int(cnt["kmissing"][0]))


def test_dump_write_corner_cases(tmp_path):
    """Digit roll-overs across chunk borders and gaps, names longer than the fixed-width copy, pairs
    outside the shared tail table, tails longer than its entries (tiny peak), changing K* models
    between calls (the table is per model), every thread count, empty and one-line records."""
    from merfin_b200 import api, lib as L
    lib = L.load()
    rng = np.random.default_rng(11)
    n = 1_200_000                                          # 10 chunks of 2^17, crosses 9 -> 10 ... 999999 -> 1000000
    vr = rng.choice(np.array([0, 1, 3, 14, 15, 16, 29, 30, 45, 2047, 2048, 5000, 4_000_000_000], dtype=np.uint32), size=n)
    va = rng.choice(np.array([0, 1, 2, 3, 31, 32, 33, 70000], dtype=np.uint32), size=n)
    ok = (rng.random(n) > 0.02).astype(np.uint8)
    ok[5000:9000] = 0                                      # a gap (an N run)
    ok[131072 - 3:131072 + 3] = 1                          # consecutive positions across a chunk border
    models = [api.KModel(15.0), api.KModel(1e-7), api.KModel(15.0, [0, 1, 1, 2, 7], [0.1, 0.2, 0.3, 0.4, 0.5]), api.KModel(15.0)]
    names = ["s", "contig_with_a_name_longer_than_thirty_two_characters_0001", "x" * 32, "y" * 33]
    for i, km in enumerate(models):
        ident = names[i % len(names)]
        want, per = api.dump_text([ident], np.array([0, n], dtype=np.uint64), vr, va, ok, km, 21)
        kmiss, _, kasm = per.strip().split("\t")[1:]
        for threads in (1, 2, 7):
            got, cnt = _dump(lib, L, api, ident, vr, va, ok, km, threads, str(tmp_path / "d.txt"))
            assert got == want, (i, threads)
            assert cnt == (int(kasm), int(kmiss))
    km = models[0]
    for m in (0, 1, 2, 131072, 131073):                    # empty, tiny, exactly one chunk, one more
        want, _ = api.dump_text(["r"], np.array([0, m], dtype=np.uint64), vr[:m], va[:m], ok[:m], km, 21)
        got, _ = _dump(lib, L, api, "r", vr[:m].copy(), va[:m].copy(), ok[:m].copy(), km, 4, str(tmp_path / "e.txt"))
        assert got == want, m


def test_dump_write_concurrent_calls_with_different_models(tmp_path):
    """Two host threads formatting at once with different K* models must not see each other's tails."""
    import threading
    from merfin_b200 import api, lib as L
    lib = L.load()
    rng = np.random.default_rng(12)
    n = 600_000
    vr = rng.integers(0, 60, size=n).astype(np.uint32)
    va = rng.integers(0, 4, size=n).astype(np.uint32)
    ok = np.ones(n, dtype=np.uint8)
    kms = [api.KModel(10.0), api.KModel(23.0)]
    wants = [api.dump_text(["c%d" % i], np.array([0, n], dtype=np.uint64), vr, va, ok, km, 21)[0] for i, km in enumerate(kms)]
    assert wants[0] != wants[1].replace(b"c1", b"c0")
    results = {}

    def work(i):
        for rep in range(4):
            results[(i, rep)] = _dump(lib, L, api, "c%d" % i, vr, va, ok, kms[i], 3, str(tmp_path / ("t%d.txt" % i)))[0]

    th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    for (i, rep), got in results.items():
        assert got == wants[i], (i, rep)


def test_dump_formatter_under_sanitizers(tmp_path):
    """report.cpp built for the host with ASan+UBSan and with TSan, driven by tests/dump_sanitize_harness.cpp
    (exact-size heap inputs, short / 32-byte / long names, long and 'inf' tails, 1 / 3 / 8 threads): the
    fixed-width copies stay inside their buffers and the shared tail table / chunk hand-off are race-free."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    srcs = [os.path.join(root, "merfin_b200", "csrc", "report.cpp"), os.path.join(root, "tests", "dump_sanitize_harness.cpp")]
    for name, flags in (("asan", ["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined"]), ("tsan", ["-fsanitize=thread"])):
        exe = str(tmp_path / name)
        b = subprocess.run(["g++", "-O1", "-g", "-std=c++17", *flags, "-I/usr/local/cuda/include", "-x", "c++", *srcs, "-o", exe, "-pthread"],
                           capture_output=True, text=True)
        if b.returncode != 0 and "sanitizer" in b.stderr.lower() or "cannot find -l" in b.stderr:
            import pytest
            pytest.skip("sanitizer runtime not installed")
        assert b.returncode == 0, b.stderr[-2000:]
        p = subprocess.run([exe], capture_output=True, text=True, timeout=600)
        assert p.returncode == 0 and p.stdout.strip().endswith("ok"), (name, p.stdout[-500:], p.stderr[-3000:])


def test_cli_dump_driver_batches_match_reference_files(tmp_path, golden_dir):
    """The CLI's -dump driver (csrc/host/dump_driver.cpp) with the CPU oracle standing in for mfn_dump:
    whatever the batch size (one batch, one record per batch, a few records per batch) and thread count,
    the file is the reference's ref.dump byte for byte and the per-record log lines are ref.dump.perseq."""
    import gzip
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "dump_driver_harness")
    obj = str(tmp_path / "oracle.o")
    subprocess.check_call(["gcc", "-O2", "-w", "-std=gnu11", "-c", os.path.join(root, "oracle", "mfn_oracle.c"), "-o", obj])
    so_dir = os.path.join(root, "merfin_b200")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-w", "-pthread", "-o", exe, os.path.join(root, "tests", "dump_driver_harness.cpp"),
                           os.path.join(root, "merfin_b200", "csrc", "host", "dump_driver.cpp"),
                           os.path.join(root, "merfin_b200", "csrc", "host", "fasta.cpp"), obj,
                           "-L" + so_dir, "-lmerfin_b200", "-Wl,-rpath," + so_dir, "-lm"])
    for name, peak in (("g1_k21", 8), ("g2_k6", 3), ("g3_k31", 6)):
        d = os.path.join(golden_dir, name)
        want = gzip.open(os.path.join(d, "ref.dump.gz"), "rb").read()
        want_per = open(os.path.join(d, "ref.dump.perseq")).read()
        ncalls = set()
        for batch, threads in ((0, 0), (1, 1), (7000, 3), (20000, 2), (10**9, 4)):
            out, per = str(tmp_path / "o.dump"), str(tmp_path / "o.perseq")
            p = subprocess.run([exe, os.path.join(d, "asm.fa"), os.path.join(d, "reads.meryl"), os.path.join(d, "asm.meryl"), str(peak),
                                out, per, str(batch), str(threads)], capture_output=True, text=True)
            assert p.returncode == 0, p.stderr
            assert open(out, "rb").read() == want, (name, batch)
            assert open(per).read() == want_per, (name, batch)
            ncalls.add(int(p.stdout))
        assert len(ncalls) > 1                       # the batch sizes above really produced different batchings
