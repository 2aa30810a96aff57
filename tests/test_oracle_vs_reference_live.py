"""Differential test of the CPU oracle against the UNMODIFIED reference binaries run NOW on fresh
random inputs (oracle/_ref/{meryl,merfin}; skipped where the reference sources were not available
to build them).  The committed golden files pin three fixed cases; this widens the pin to more
mer sizes (odd, even, the 32-mer maximum), peaks, -min/-max windows and -prob tables.  CPU only."""
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MERFIN = os.path.join(ROOT, "oracle", "_ref", "merfin")
MERYL = os.path.join(ROOT, "oracle", "_ref", "meryl")

pytestmark = pytest.mark.skipif(not (os.path.exists(MERFIN) and os.path.exists(MERYL)),
                                reason="oracle/_ref not built (reference sources absent)")


def _run(cmd, cwd):
    p = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stderr


CASES = [  # k, total bases, records, coverage, peak, (min, max), with -prob
    (9, 6000, 3, 6, 2.5, (0, None), False),
    (16, 30000, 4, 8, 6.0, (2, None), False),
    (21, 40000, 3, 10, 8.0, (0, 9), True),
    (24, 30000, 2, 7, 5.0, (0, None), True),
    (32, 30000, 3, 9, 7.0, (3, 30), False),
]


@pytest.mark.parametrize("k,total,nrec,cov,peak,mm,prob", CASES)
def test_oracle_equals_reference_on_fresh_input(tmp_path, k, total, nrec, cov, peak, mm, prob):
    from merfin_b200 import synth
    d = str(tmp_path)
    seed = 1000 + k
    truth = synth.make_assembly(total, nrec, seed, n_run=20, lower_run=40, dup_len=min(2000, total // 10))
    if k % 2 == 0:                                  # plant reverse-palindromes (Q2) in record 0
        rng = np.random.default_rng(seed)
        half = synth.random_bases(k // 2, rng)
        pal = np.concatenate([half, synth.revcomp(half)])
        b = truth[0][1].copy()
        for pos in (100, 100 + 3 * k, len(b) // 2):
            b[pos:pos + k] = pal
        truth[0] = (truth[0][0], b)
    synth.write_fasta(os.path.join(d, "asm.fa"), truth, width=61)
    reads = synth.simulate_reads(truth, cov, 100, 0.01, seed + 1, skip=[(0, 200, 200 + 4 * k)])
    synth.write_fasta(os.path.join(d, "reads.fa"), reads, prefix="r")
    for fa, db in (("reads.fa", "reads.meryl"), ("asm.fa", "asm.meryl")):
        _run([MERYL, "threads=1", "count", f"k={k}", "memory=1", fa, "output", db], d)
    base = [MERFIN, "-sequence", "asm.fa", "-seqmers", "asm.meryl", "-readmers", "reads.meryl", "-threads", "1",
            "-memory", "4", "-peak", str(peak)]
    mn, mx = mm
    if mn:
        base += ["-min", str(mn)]
    if mx:
        base += ["-max", str(mx)]
    probfile = None
    if prob:
        probfile = os.path.join(d, "prob.csv")
        with open(probfile, "w") as f:
            for v in range(1, 40):
                f.write("%d,%.3f\n" % (0 if v <= 2 else max(1, round(v / peak)), max(0.1, 1.0 - v / 60.0)))
        base += ["-prob", "prob.csv"]
    err_h = _run(base + ["-hist", "-output", "ref.hist"], d)
    _run(base + ["-dump", "-output", "ref.dump"], d)
    err_c = _run(base + ["-completeness"], d)

    S, A = O.DB(os.path.join(d, "reads.meryl")), O.DB(os.path.join(d, "asm.meryl"))
    assert S.k == k
    recs = O.read_fasta(os.path.join(d, "asm.fa"))
    R, Am = O.Lookup(S, mn or 0, mx or 0xFFFFFFFF), O.Lookup(A)
    km = O.KModel.from_file(peak, probfile) if prob else O.KModel(peak)
    h = O.Hist()
    dump = []
    for ident, b in recs:
        h.add_record(R, Am, km, b)
        dump.append(O.dump_record(R, Am, km, ident, b)[0])
    assert h.report() == open(os.path.join(d, "ref.hist"), "rb").read()
    want = "\n".join(l for l in err_h.splitlines() if re.match(r"^(K-mers|Missing QV|Merfin QV)", l)) + "\n"
    assert h.summary(k) == want
    assert b"".join(dump) == open(os.path.join(d, "ref.dump"), "rb").read()
    # -completeness streams raw values (Q4: no -min/-max)
    _, _, t, u = O.completeness(S, A, km)
    tail = ("TOTAL readK:   %15.2f\nTOTAL undrcpy:    %15.5f\nCOMPLETENESS:             %0.5f\n" % (t, u, 1.0 - u / t))
    assert tail in err_c


@pytest.mark.parametrize("k,n", [(21, 200_000), (31, 50_000), (12, 30_000)])
def test_reference_meryl_reads_databases_written_by_the_library(tmp_path, k, n):
    """mfn_meryl_write (the generator side of the benchmarks) produces databases the unmodified
    reference tool lists exactly: same k-mers, same values, same order."""
    from merfin_b200 import api
    rng = np.random.default_rng(k)
    km = np.unique(rng.integers(0, (1 << (2 * k)) - 1, size=n, dtype=np.uint64, endpoint=True))
    vv = rng.integers(1, 3000, size=len(km), dtype=np.uint32)
    vv[::500] = rng.integers(70000, 4_000_000, size=len(vv[::500]), dtype=np.uint32)
    d = os.path.join(str(tmp_path), "w.meryl")
    api.meryl_write(d, k, km, vv)
    p = subprocess.run([MERYL, "threads=1", "print", d], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = p.stdout.split("\n")[:-1]
    assert len(lines) == len(km)
    dec = "ACTG"                                            # kmers-tiny.H:69-83: A=0 C=1 T=2 G=3
    def text(x):
        return "".join(dec[(int(x) >> (2 * (k - 1 - i))) & 3] for i in range(k))
    idx = rng.integers(0, len(km), size=2000)
    for i in idx:                                           # spot-check the text form, then every value and the order
        assert lines[i] == "%s\t%d" % (text(km[i]), vv[i])
    got_v = np.array([int(l.split("\t")[1]) for l in lines], dtype=np.uint32)
    assert (got_v == vv).all()
    # the statistics block the writer stored is the one the reference reports, and mfn_meryl_info reads it back
    p = subprocess.run([MERYL, "statistics", d], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    stats = {w: int(n) for w, n in re.findall(r"^\s+(unique|distinct|present)\s+(\d+)", p.stdout, re.M)}
    info = api.meryl_info(d)
    assert stats == {"unique": info["n_unique"], "distinct": info["n_distinct"], "present": info["n_total"]}
    assert info["n_distinct"] == len(km) and info["n_total"] == int(vv.astype(np.uint64).sum())


def test_oracle_decodes_databases_made_by_meryl_set_operations(tmp_path):
    """Read databases in the wild are usually `meryl union-sum` of per-run counts, written by meryl's
    streaming writer (kmers-writer-stream.C) rather than the block writer `count` uses, and filtered
    ones (`greater-than`) have sparse prefixes.  The oracle's decoder against `meryl print` of each,
    and mfn_meryl_info against `meryl statistics`."""
    from merfin_b200 import synth, api
    d = str(tmp_path)
    k = 21
    t1 = synth.make_assembly(60000, 3, 1, n_run=10, lower_run=20, dup_len=500)
    t2 = synth.make_assembly(50000, 2, 2, n_run=10, lower_run=20, dup_len=500)
    synth.write_fasta(os.path.join(d, "a.fa"), t1)
    synth.write_fasta(os.path.join(d, "b.fa"), t2 + t1[:1])
    for fa, db in (("a.fa", "a.meryl"), ("b.fa", "b.meryl")):
        _run([MERYL, "threads=2", "count", f"k={k}", "memory=1", fa, "output", db], d)
    _run([MERYL, "threads=2", "union-sum", "a.meryl", "b.meryl", "output", "u.meryl"], d)
    _run([MERYL, "threads=1", "greater-than", "1", "u.meryl", "output", "g.meryl"], d)
    _run([MERYL, "threads=1", "intersect-min", "a.meryl", "b.meryl", "output", "i.meryl"], d)
    dec = "ACTG"
    for db in ("u.meryl", "g.meryl", "i.meryl"):
        p = subprocess.run([MERYL, "threads=1", "print", os.path.join(d, db)], capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stderr[-2000:]
        lines = p.stdout.splitlines()
        D = O.DB(os.path.join(d, db))
        assert D.n == len(lines) > 1000, db
        want_v = np.array([int(l.split()[1]) for l in lines], dtype=np.uint32)
        assert (D.val == want_v).all(), db
        for i in np.random.default_rng(1).integers(0, D.n, size=500):
            x = int(D.lo[i])
            assert "".join(dec[(x >> (2 * (k - 1 - j))) & 3] for j in range(k)) == lines[i].split()[0]
        info = api.meryl_info(os.path.join(d, db))
        assert info["k"] == k and info["n_distinct"] == D.n and info["n_total"] == int(want_v.astype(np.uint64).sum())
