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
