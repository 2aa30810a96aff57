"""GPU parity tests of the k = 33..64 path (merfin_b200/csrc/wide.cu: tables of sorted 128-bit k-mers, k_scan_wide,
k_join_wide, the sort-based count) through the C ABI and the CLI, against the CPU oracle and the reference's own output
files for k = 40 and k = 64 (tests/golden/g4_k40 with the variant modes, g5_k64).  Bit-exact integers, byte-identical text."""
import gzip
import os
import re
import subprocess

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MERFIN = os.environ.get("MFN_TEST_CLI", os.path.join(ROOT, "merfin_b200", "bin", "merfin"))
SETS = [("g4_k40", 40, 6.0), ("g5_k64", 64, 6.0)]


@pytest.fixture(scope="module")
def ctx():
    from merfin_b200 import api
    c = api.Context(0)
    yield c
    c.close()


def _golden(golden_dir, name):
    d = os.path.join(golden_dir, name)
    return d, O.read_fasta(os.path.join(d, "asm.fa"))


@pytest.mark.parametrize("name,k,peak", SETS)
def test_wide_table_load_and_value(ctx, golden_dir, name, k, peak):
    """mfn_table_load + mfn_table_lookup_wide == merylExactLookup::load + value() for 128-bit k-mers, with -min / -max."""
    from merfin_b200 import api, lib as L
    d, recs = _golden(golden_dir, name)
    for sub in ("reads.meryl", "asm.meryl"):
        db = O.DB(os.path.join(d, sub))
        assert db.k == k
        for mn, mx in ((0, 0xFFFFFFFF), (2, 0xFFFFFFFF), (3, 9), (50, 40)):
            t = api.MerylExactLookup.load(ctx, os.path.join(d, sub), mn, mx)
            ref = O.Lookup(db, mn, mx)
            assert t.nKmers() == ref.n
            los, his = [db.lo], [db.hi]
            for _, b in recs:                                   # forward and reverse-complement k-mer of every window
                f, fh, r, rh, ok = O.windows(b, k)
                los += [f[ok == 1], r[ok == 1]]; his += [fh[ok == 1], rh[ok == 1]]
            rng = np.random.default_rng(1)
            hbits = 2 * k - 64
            los.append(rng.integers(0, 1 << 63, size=5000, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=5000, dtype=np.uint64))
            his.append(rng.integers(0, 1 << min(hbits, 63), size=5000, dtype=np.uint64))
            los.append(np.array([0, 0xFFFFFFFFFFFFFFFF], dtype=np.uint64))
            his.append(np.array([0, (1 << hbits) - 1], dtype=np.uint64))
            lo, hi = np.concatenate(los), np.concatenate(his)
            assert (t.values_wide(lo, hi) == ref.values(lo, hi)).all()
            info = t.info
            assert info.k == k and info.canonical == 0 and info.slot_bits == 128 and info.n_kmers == ref.n
            assert info.min_value == max(mn, 1) and info.value_offset == max(mn, 1) - 1
            # the entry points that exist only for 64-bit k-mers say so
            with pytest.raises(L.MfnError) as e:
                t.values([1, 2, 3])
            assert e.value.status == L.MFN_ERR_UNSUPPORTED
            with pytest.raises(L.MfnError):
                t.export()
            with pytest.raises(L.MfnError):
                t.fuse(t)
            t.close()


def _hist_check(ctx, d, recs, k, peak, mn, mx, prob, tag):
    from merfin_b200 import api
    R = api.MerylExactLookup.load(ctx, os.path.join(d, "reads.meryl"), mn, mx)
    A = api.MerylExactLookup.load(ctx, os.path.join(d, "asm.meryl"))
    km = api.KModel.load_Kmetric(peak, prob)
    res = api.hist(ctx, R, A, recs, km)
    if tag:
        assert res.file_text() == open(os.path.join(d, f"{tag}.hist"), "rb").read()
        assert res.summary_text() == open(os.path.join(d, f"{tag}.hist.summary")).read()
        assert res.per_seq_text() == open(os.path.join(d, f"{tag}.hist.perseq")).read()
    oR, oA = O.Lookup(O.DB(os.path.join(d, "reads.meryl")), mn, mx), O.Lookup(O.DB(os.path.join(d, "asm.meryl")))
    okm = O.KModel(peak, km.K, km.P)
    h = O.Hist()
    for _, b in recs:
        h.add_record(oR, oA, okm, b)
    s, _, _ = res.report()
    assert (s.kasm, s.kmissing) == (h.kasm, h.kmissing)
    assert res.file_text() == h.report()
    assert abs(s.qv_star - O.histoQV(h.kmissing + h.koverCpy, h.kasm, k)) < 1e-6
    assert abs(s.qv_missing - O.histoQV(h.kmissing, h.kasm, k)) < 1e-6
    # the same through a resident sequence set cut into shards (what the multi-GPU callers do)
    seq = api.SeqSet(ctx, recs)
    acc = api.Accum(ctx, len(recs))
    for sh in range(3):
        api.hist_resident(ctx, R, A, seq, km, acc, sh, 3)
    per, pairs = acc.fetch()
    key = lambda p: sorted((int(a), int(b), int(c)) for a, b, c in zip(p["v_read"], p["v_asm"], p["count"]))
    assert key(pairs) == key(res.pairs) and (per == res.per_seq).all()
    acc.close(); seq.close(); R.close(); A.close()


@pytest.mark.parametrize("name,k,peak", SETS)
def test_wide_hist_matches_reference_files(ctx, golden_dir, name, k, peak):
    d, recs = _golden(golden_dir, name)
    _hist_check(ctx, d, recs, k, peak, 0, 0xFFFFFFFF, None, "ref")
    _hist_check(ctx, d, recs, k, peak, 2, 12, None, None)                      # -min / -max against the oracle
    if name == "g4_k40":
        _hist_check(ctx, d, recs, k, peak, 0, 0xFFFFFFFF, os.path.join(d, "prob.csv"), "ref.prob")


@pytest.mark.parametrize("name,k,peak", SETS)
def test_wide_dump_matches_reference_files(ctx, golden_dir, name, k, peak):
    from merfin_b200 import api
    d, recs = _golden(golden_dir, name)
    R = api.MerylExactLookup.load(ctx, os.path.join(d, "reads.meryl"))
    A = api.MerylExactLookup.load(ctx, os.path.join(d, "asm.meryl"))
    idents, off, vr, va, ok = api.dump(ctx, R, A, recs)
    text, per = api.dump_text(idents, off, vr, va, ok, api.KModel(peak), k)
    assert text == gzip.open(os.path.join(d, "ref.dump.gz"), "rb").read()
    assert per == open(os.path.join(d, "ref.dump.perseq")).read()
    if name == "g4_k40":
        km = api.KModel.load_Kmetric(peak, os.path.join(d, "prob.csv"))
        text, _ = api.dump_text(idents, off, vr, va, ok, km, k)
        assert text == gzip.open(os.path.join(d, "ref.prob.dump.gz"), "rb").read()
    oR, oA = O.Lookup(O.DB(os.path.join(d, "reads.meryl"))), O.Lookup(O.DB(os.path.join(d, "asm.meryl")))
    for i, (_, b) in enumerate(recs):
        evr, eva, eok = O.pairs(oR, oA, b)
        a = int(off[i]); n = len(eok)
        assert (ok[a:a + n] == eok).all() and (vr[a:a + n] == evr).all() and (va[a:a + n] == eva).all()
        assert (ok[a + n:int(off[i + 1])] == 0).all()
    # odd batches: empty, shorter than k, all N, lower case, and the records again
    extra = [("e", b""), ("s", b"ACGT" * 5), ("n", b"N" * 300), ("lc", recs[0][1][:500].lower())] + list(recs)
    idents, off, vr, va, ok = api.dump(ctx, R, A, extra)
    for i, (_, b) in enumerate(extra):
        evr, eva, eok = O.pairs(oR, oA, b)
        a = int(off[i]); n = len(eok)
        assert (ok[a:a + n] == eok).all() and (vr[a:a + n] == evr * eok).all() and (va[a:a + n] == eva * eok).all()
        assert (ok[a + n:int(off[i + 1])] == 0).all()
    R.close(); A.close()


@pytest.mark.parametrize("name,k,peak", SETS)
def test_wide_completeness_matches_reference(ctx, golden_dir, name, k, peak):
    from merfin_b200 import api
    d, _ = _golden(golden_dir, name)
    t, u, pairs = api.completeness(ctx, os.path.join(d, "reads.meryl"), os.path.join(d, "asm.meryl"), api.KModel(peak))
    assert open(os.path.join(d, "ref.completeness")).read().endswith(api.completeness_text(t, u))
    assert int(pairs["count"].sum()) == O.DB(os.path.join(d, "reads.meryl")).n
    if name == "g4_k40":
        km = api.KModel.load_Kmetric(peak, os.path.join(d, "prob.csv"))
        t2, u2, _ = api.completeness(ctx, os.path.join(d, "reads.meryl"), os.path.join(d, "asm.meryl"), km)
        assert open(os.path.join(d, "ref.prob.completeness")).read().endswith(api.completeness_text(t2, u2))
    # the join over two resident tables gives the same sums
    R = api.MerylExactLookup.load(ctx, os.path.join(d, "reads.meryl"))
    A = api.MerylExactLookup.load(ctx, os.path.join(d, "asm.meryl"))
    t3, u3, _ = api.completeness_resident(ctx, R, A, api.KModel(peak))
    assert (t3, u3) == (t, u)
    R.close(); A.close()


@pytest.mark.parametrize("name,k,peak", SETS)
def test_wide_count_sequence_equals_meryl_count(ctx, golden_dir, name, k, peak):
    """mfn_table_count_seq for k > 32 (canonical k-mers of all windows, two radix sorts, run lengths) == `meryl count`."""
    from merfin_b200 import api
    d, recs = _golden(golden_dir, name)
    db = O.DB(os.path.join(d, "asm.meryl"))
    seq = api.SeqSet(ctx, recs)
    t = api.MerylExactLookup.count_sequence(ctx, k, seq)
    assert t.nKmers() == db.n
    assert (t.values_wide(db.lo, db.hi) == db.val).all()
    rd = O.DB(os.path.join(d, "reads.meryl"))
    want = O.Lookup(db).values(rd.lo, rd.hi)
    assert (t.values_wide(rd.lo, rd.hi) == want).all()               # (read k-mers the assembly does not hold: 0)
    empty = api.SeqSet(ctx, [("n", b"N" * 100)])
    t0 = api.MerylExactLookup.count_sequence(ctx, k, empty)
    assert t0.nKmers() == 0 and not t0.values_wide(db.lo[:10], db.hi[:10]).any()
    t0.close(); empty.close(); t.close(); seq.close()


def test_wide_k_mismatch_is_an_error(ctx, golden_dir):
    from merfin_b200 import api, lib as L
    R = api.MerylExactLookup.load(ctx, os.path.join(golden_dir, "g4_k40", "reads.meryl"))
    for other in ("g5_k64", "g3_k31"):
        A = api.MerylExactLookup.load(ctx, os.path.join(golden_dir, other, "asm.meryl"))
        with pytest.raises(L.MfnError) as e:
            api.hist(ctx, R, A, [("x", b"ACGT" * 50)], api.KModel(6.0))
        assert e.value.status == L.MFN_ERR_K_MISMATCH
        A.close()
    R.close()


# ---------------------------------------------------------------------------------- the CLI
def _run(args, cwd):
    p = subprocess.run([MERFIN] + args, cwd=cwd, capture_output=True, text=True, timeout=150)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stderr


def _summary(stderr):
    return "\n".join(l for l in stderr.splitlines() if re.match(r"^(K-mers|Missing QV|Merfin QV)", l)) + "\n"


@pytest.mark.parametrize("name,k,peak", SETS)
def test_wide_cli_hist_dump_completeness(golden_dir, tmp_path, name, k, peak):
    d = os.path.join(golden_dir, name)
    base = ["-sequence", os.path.join(d, "asm.fa"), "-seqmers", os.path.join(d, "asm.meryl"),
            "-readmers", os.path.join(d, "reads.meryl"), "-peak", str(int(peak))]
    err = _run(base + ["-hist", "-output", "out.hist"], str(tmp_path))
    assert open(tmp_path / "out.hist", "rb").read() == open(os.path.join(d, "ref.hist"), "rb").read()
    assert _summary(err) == open(os.path.join(d, "ref.hist.summary")).read()
    assert open(os.path.join(d, "ref.hist.perseq")).read() in err
    _run(base + ["-dump", "-output", "out.dump"], str(tmp_path))
    assert open(tmp_path / "out.dump", "rb").read() == gzip.open(os.path.join(d, "ref.dump.gz"), "rb").read()
    err = _run(base + ["-completeness"], str(tmp_path))
    tail = [l for l in open(os.path.join(d, "ref.completeness")).read().splitlines() if not l.startswith("thread")]
    for l in tail:
        assert l in err
    # without -seqmers the assembly k-mers are counted on the GPU; two contexts share the work
    err = _run(base[:2] + base[4:] + ["-hist", "-output", "out2.hist"], str(tmp_path))
    assert open(tmp_path / "out2.hist", "rb").read() == open(os.path.join(d, "ref.hist"), "rb").read()
    err = _run(base + ["-hist", "-output", "out3.hist", "-devices", "0,0"], str(tmp_path))
    assert open(tmp_path / "out3.hist", "rb").read() == open(os.path.join(d, "ref.hist"), "rb").read()
    assert _summary(err) == open(os.path.join(d, "ref.hist.summary")).read()
    err = _run(base[:2] + base[4:] + ["-completeness", "-devices", "0,0"], str(tmp_path))
    for l in tail:
        assert l in err


@pytest.mark.parametrize("mode", ["polish", "filter", "better", "strict", "loose"])
def test_wide_cli_variant_modes(golden_dir, tmp_path, mode):
    d = os.path.join(golden_dir, "g4_k40")
    args = ["-sequence", os.path.join(d, "asm.fa"), "-seqmers", os.path.join(d, "asm.meryl"),
            "-readmers", os.path.join(d, "reads.meryl"), "-peak", "6", "-vcf", os.path.join(d, "calls.vcf"),
            "-" + mode, "-output", "out"]
    if mode == "polish":
        args.append("-debug")
    _run(args, str(tmp_path))
    produced = "out.polish.vcf" if mode == "polish" else "out.filter.vcf"      # Q7
    assert open(tmp_path / produced).read() == open(os.path.join(d, f"ref.{mode}.vcf")).read()
    if mode == "polish":
        assert gzip.open(tmp_path / "out.00.debug.gz", "rt").read() == gzip.open(os.path.join(d, "ref.polish.debug.gz"), "rt").read()
        _run(args[:-1] + ["-prob", os.path.join(d, "prob.csv")], str(tmp_path))
        assert open(tmp_path / "out.polish.vcf").read() == open(os.path.join(d, "ref.probpolish.vcf")).read()
