"""GPU parity tests: the CUDA path through the C ABI vs the CPU oracle and the committed
reference outputs.  Integer results must be bit-exact; text files byte-identical; QV within 1e-6."""
import gzip
import os

import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu

SETS = [("g1_k21", 21, 8.0), ("g2_k6", 6, 3.0), ("g3_k31", 31, 6.0)]


@pytest.fixture(scope="module")
def ctx():
    from merfin_b200 import api
    c = api.Context(0)
    yield c
    c.close()


def _golden(golden_dir, name):
    d = os.path.join(golden_dir, name)
    return d, O.read_fasta(os.path.join(d, "asm.fa"))


def _all_window_kmers(recs, k):
    fs, rs = [], []
    for _, b in recs:
        f, fh, r, rh, ok = O.windows(b, k)
        fs.append(f[ok == 1]); rs.append(r[ok == 1])
    return np.concatenate(fs), np.concatenate(rs)


# ---------------------------------------------------------------------------------- tables
@pytest.mark.parametrize("name,k,peak", SETS)
def test_table_load_and_value(ctx, golden_dir, name, k, peak):
    """mfn_table_load + mfn_table_lookup == merylExactLookup::load + value()."""
    from merfin_b200 import api
    d, recs = _golden(golden_dir, name)
    for sub in ("reads.meryl", "asm.meryl"):
        db = O.DB(os.path.join(d, sub))
        for mn, mx in ((0, 0xFFFFFFFF), (2, 0xFFFFFFFF), (3, 9), (50, 40)):
            t = api.MerylExactLookup.load(ctx, os.path.join(d, sub), mn, mx)
            ref = O.Lookup(db, mn, mx)
            assert t.nKmers() == ref.n
            f, r = _all_window_kmers(recs, k)
            rng = np.random.default_rng(1)
            rnd = rng.integers(0, 1 << (2 * k), size=5000, dtype=np.uint64)
            q = np.concatenate([db.lo, f, r, rnd, np.array([0, (1 << (2 * k)) - 1], dtype=np.uint64)])
            assert (t.values(q) == ref.values(q)).all()
            info = t.info
            assert info.k == k and info.canonical == 1
            assert info.n_level1 + info.n_level2 + info.n_stash == info.n_kmers
            km, vv = t.export()
            keep = (db.val >= max(mn, 1)) & (db.val <= (mx if mx != 0xFFFFFFFF else db.val.max()))
            assert (km == db.lo[keep]).all() and (vv == db.val[keep] - (max(mn, 1) - 1)).all()
            t.close()


def test_table_levels_stash_and_big_values(ctx, monkeypatch):
    """Overloaded buckets spill to level 2 and the side table; huge values keep their excess exactly."""
    from merfin_b200 import api
    monkeypatch.setenv("MFN_TABLE_LOAD", "1.9")
    rng = np.random.default_rng(7)
    for k in (15, 21, 32):
        n = 300000
        kmers = np.unique(rng.integers(0, (1 << (2 * k)) - 1, size=n, dtype=np.uint64, endpoint=True))
        values = rng.integers(1, 50, size=len(kmers)).astype(np.uint32)
        values[::97] = rng.integers(1 << 20, 0xFFFFFFFF, size=len(values[::97]), dtype=np.uint64).astype(np.uint32)
        values[5] = 0xFFFFFFFF
        t = api.MerylExactLookup.from_kmers(ctx, k, kmers, values)
        info = t.info
        assert info.n_kmers == len(kmers)
        assert info.n_level2 > 0
        if k != 21:      # k=21 at this size gets 64-bit slots with full 32-bit values: nothing saturates
            assert info.n_bigvalue > 0 and info.value_bits < 20
        assert (t.values(kmers) == values).all()
        absent = np.setdiff1d(rng.integers(0, (1 << (2 * k)) - 1, size=20000, dtype=np.uint64, endpoint=True), kmers)
        assert (t.values(absent) == 0).all()
        km, vv = t.export()
        assert (km == kmers).all() and (vv == values).all()
        t.close()


def test_count_sequence_equals_meryl_count(ctx, golden_dir):
    """mfn_table_count_seq builds what `meryl count` of the assembly builds (asm.meryl)."""
    from merfin_b200 import api
    for name, k, _ in SETS:
        d, recs = _golden(golden_dir, name)
        db = O.DB(os.path.join(d, "asm.meryl"))
        seq = api.SeqSet(ctx, recs)
        t = api.MerylExactLookup.count_sequence(ctx, k, seq)
        km, vv = t.export()
        assert (km == db.lo).all() and (vv == db.val).all()
        t.close(); seq.close()


# ---------------------------------------------------------------------------------- -hist
def _hist_check(ctx, d, recs, k, peak, mn, mx, prob, tag):
    from merfin_b200 import api
    R = api.MerylExactLookup.load(ctx, os.path.join(d, "reads.meryl"), mn, mx)
    A = api.MerylExactLookup.load(ctx, os.path.join(d, "asm.meryl"))
    km = api.KModel.load_Kmetric(peak, prob)
    res0 = api.hist(ctx, R, A, recs, km)           # one probe per database
    R.fuse(A)                                      # one probe of the fused table per window
    assert R.joint_info.fused == 1
    res = api.hist(ctx, R, A, recs, km)
    assert (res0.pairs == res.pairs).all() and (res0.per_seq == res.per_seq).all()
    assert res.file_text() == open(os.path.join(d, f"{tag}.hist"), "rb").read()
    assert res.summary_text() == open(os.path.join(d, f"{tag}.hist.summary")).read()
    per = os.path.join(d, f"{tag}.hist.perseq")
    if os.path.exists(per):
        assert res.per_seq_text() == open(per).read()
    # full-precision QV against the oracle's position-order accumulation (reference prints %.2f only)
    oR, oA = O.Lookup(O.DB(os.path.join(d, "reads.meryl")), mn, mx), O.Lookup(O.DB(os.path.join(d, "asm.meryl")))
    okm = O.KModel(peak, km.K, km.P)
    h = O.Hist()
    for _, b in recs:
        h.add_record(oR, oA, okm, b)
    s, _, _ = res.report()
    assert (s.kasm, s.kmissing) == (h.kasm, h.kmissing)
    assert abs(s.qv_star - O.histoQV(h.kmissing + h.koverCpy, h.kasm, k)) < 1e-6
    assert abs(s.qv_missing - O.histoQV(h.kmissing, h.kasm, k)) < 1e-6
    R.close(); A.close()


@pytest.mark.parametrize("name,k,peak", SETS)
def test_hist_matches_reference_files(ctx, golden_dir, name, k, peak):
    d, recs = _golden(golden_dir, name)
    _hist_check(ctx, d, recs, k, peak, 0, 0xFFFFFFFF, None, "ref")


def test_hist_min_max_and_prob(ctx, golden_dir):
    d, recs = _golden(golden_dir, "g1_k21")
    _hist_check(ctx, d, recs, 21, 1.0, 3, 0xFFFFFFFF, None, "ref.min3")
    _hist_check(ctx, d, recs, 21, 8.0, 2, 12, None, "ref.min2max12")
    _hist_check(ctx, d, recs, 21, 8.0, 0, 0xFFFFFFFF, os.path.join(d, "prob.csv"), "ref.prob")


# ---------------------------------------------------------------------------------- -dump
@pytest.mark.parametrize("name,k,peak", SETS)
def test_dump_matches_reference_files(ctx, golden_dir, name, k, peak):
    from merfin_b200 import api
    d, recs = _golden(golden_dir, name)
    R = api.MerylExactLookup.load(ctx, os.path.join(d, "reads.meryl"))
    A = api.MerylExactLookup.load(ctx, os.path.join(d, "asm.meryl"))
    idents, off, vr0, va0, ok0 = api.dump(ctx, R, A, recs)
    R.fuse(A)
    idents, off, vr, va, ok = api.dump(ctx, R, A, recs)
    assert (vr0 == vr).all() and (va0 == va).all() and (ok0 == ok).all()
    text, per = api.dump_text(idents, off, vr, va, ok, api.KModel(peak), k)
    assert text == gzip.open(os.path.join(d, "ref.dump.gz"), "rb").read()
    assert per == open(os.path.join(d, "ref.dump.perseq")).read()
    if name == "g1_k21":
        km = api.KModel.load_Kmetric(peak, os.path.join(d, "prob.csv"))
        text, _ = api.dump_text(idents, off, vr, va, ok, km, k)
        assert text == gzip.open(os.path.join(d, "ref.prob.dump.gz"), "rb").read()
    # integer pairs against the oracle, position by position
    oR, oA = O.Lookup(O.DB(os.path.join(d, "reads.meryl"))), O.Lookup(O.DB(os.path.join(d, "asm.meryl")))
    for i, (_, b) in enumerate(recs):
        evr, eva, eok = O.pairs(oR, oA, b)
        a = int(off[i]); n = len(eok)
        assert (ok[a:a + n] == eok).all() and (vr[a:a + n] == evr).all() and (va[a:a + n] == eva).all()
        assert (ok[a + n:int(off[i + 1])] == 0).all()
    R.close(); A.close()


# ---------------------------------------------------------------------------------- -completeness
@pytest.mark.parametrize("name,k,peak", SETS)
def test_completeness_matches_reference(ctx, golden_dir, name, k, peak):
    from merfin_b200 import api
    d, _ = _golden(golden_dir, name)
    t, u, pairs = api.completeness(ctx, os.path.join(d, "reads.meryl"), os.path.join(d, "asm.meryl"), api.KModel(peak))
    assert open(os.path.join(d, "ref.completeness")).read().endswith(api.completeness_text(t, u))
    assert int(pairs["count"].sum()) == O.DB(os.path.join(d, "reads.meryl")).n
    if name == "g1_k21":
        km = api.KModel.load_Kmetric(peak, os.path.join(d, "prob.csv"))
        t, u, _ = api.completeness(ctx, os.path.join(d, "reads.meryl"), os.path.join(d, "asm.meryl"), km)
        assert open(os.path.join(d, "ref.prob.completeness")).read().endswith(api.completeness_text(t, u))


# ---------------------------------------------------------------------------------- path scoring
def test_score_paths_matches_oracle(ctx, golden_dir):
    from merfin_b200 import api
    d, recs = _golden(golden_dir, "g1_k21")
    R = api.MerylExactLookup.load(ctx, os.path.join(d, "reads.meryl"))
    A = api.MerylExactLookup.load(ctx, os.path.join(d, "asm.meryl"))
    oR, oA = O.Lookup(O.DB(os.path.join(d, "reads.meryl"))), O.Lookup(O.DB(os.path.join(d, "asm.meryl")))
    rng = np.random.default_rng(9)
    b = recs[0][1]
    paths = []
    for i in range(400):
        p = int(rng.integers(0, len(b) - 80))
        s = bytearray(b[p:p + int(rng.integers(1, 80))])
        if len(s) > 30 and i % 3 == 0:
            s[int(rng.integers(0, len(s)))] = ord("ACGTN"[int(rng.integers(0, 5))])
        paths.append(bytes(s))
    paths += [b"", b"A", b"N" * 30, b"acgt" * 10]
    off, vr, va, ok = api.score_paths(ctx, R, A, paths)
    for i, s in enumerate(paths):
        evr, eva, eok = O.pairs(oR, oA, s)
        a = int(off[i]); n = len(eok)
        assert (ok[a:a + n] == eok).all() and (vr[a:a + n] == evr).all() and (va[a:a + n] == eva).all()
    R.close(); A.close()


# ---------------------------------------------------------------------------------- bigger / edge cases
def _np_revcomp(x, k):
    """Reverse complement of 2-bit packed k-mers (A=0 C=1 T=2 G=3: complement = xor 2)."""
    x = x.astype(np.uint64) ^ np.uint64(0xAAAAAAAAAAAAAAAA)
    for sh, m in ((2, 0x3333333333333333), (4, 0x0F0F0F0F0F0F0F0F), (8, 0x00FF00FF00FF00FF),
                  (16, 0x0000FFFF0000FFFF)):
        x = ((x >> np.uint64(sh)) & np.uint64(m)) | ((x & np.uint64(m)) << np.uint64(sh))
    x = (x >> np.uint64(32)) | (x << np.uint64(32))
    return x >> np.uint64(64 - 2 * k)


def _synthetic_case(k, total, seed, noncanonical=False):
    from merfin_b200 import synth
    recs = synth.make_assembly(total, 5, seed, dup_len=2000)
    recs = [(n, b.tobytes()) for n, b in recs]
    recs.insert(2, ("empty", b""))
    recs.insert(3, ("short", b"ACGTACGTAC"[:k - 1]))
    recs.append(("allN", b"N" * 100))
    f, r = _all_window_kmers(recs, k)
    canon = np.minimum(f, r)
    rng = np.random.default_rng(seed)
    # asm table = true counts; read table = noisy multiples with dropouts and a few huge values
    ak, ac = np.unique(canon, return_counts=True)
    keep = rng.random(len(ak)) > 0.02
    rk = ak[keep]
    rv = (ac[keep] * rng.poisson(12, size=keep.sum())).astype(np.uint32) + 1
    rv[::1000] = 5000 + rng.integers(0, 100000, size=len(rv[::1000])).astype(np.uint32)
    extra = rng.integers(0, (1 << (2 * k)) - 1, size=len(ak) // 2, dtype=np.uint64, endpoint=True)
    if not noncanonical:            # error k-mers of a canonical read DB are canonical too
        extra = np.minimum(extra, _np_revcomp(extra, k))
    extra = np.setdiff1d(extra, ak)
    rk = np.concatenate([rk, extra]); rv = np.concatenate([rv, np.ones(len(extra), dtype=np.uint32)])
    if noncanonical:        # store the strand the assembly shows, split counts between strands
        fk, fc = np.unique(f, return_counts=True)
        return recs, (fk, fc.astype(np.uint32)), (rk, rv)
    return recs, (ak, ac.astype(np.uint32)), (rk, rv)


class _ArrLookup:
    """Oracle-side lookup over explicit arrays (same value semantics as o_lookup)."""
    def __init__(self, k, kmers, vals):
        o = np.argsort(kmers); self.k = k; self.km = kmers[o]; self.v = vals[o]

    def values(self, q):
        i = np.searchsorted(self.km, q)
        i[i >= len(self.km)] = 0
        hit = self.km[i] == q
        return np.where(hit, self.v[i], 0).astype(np.uint32)


def _expected_pairs(recs, k, Rl, Al):
    exp = {}
    per = []
    for _, b in recs:
        f, _, r, _, ok = O.windows(b, k)
        f, r = f[ok == 1], r[ok == 1]
        vr = Rl.values(f) + Rl.values(r); va = Al.values(f) + Al.values(r)
        per.append((len(f), int((vr == 0).sum())))
        keys, cnt = np.unique(np.stack([vr, va], 1), axis=0, return_counts=True)
        for (a, b2), c in zip(keys, cnt):
            exp[(int(a), int(b2))] = exp.get((int(a), int(b2)), 0) + int(c)
    return exp, per


@pytest.mark.parametrize("k,noncanon", [(21, False), (16, False), (32, False), (21, True)])
def test_hist_joint_histogram_2mbp(ctx, k, noncanon):
    """Joint (v_read, v_asm) histogram and per-record counters on a 2 Mbp set with empty/short/all-N
    records, dense-tier overflow pairs, palindromes (even k) and a non-canonical assembly DB."""
    from merfin_b200 import api
    recs, (ak, ac), (rk, rv) = _synthetic_case(k, 2_000_000, 100 + k, noncanon)
    R = api.MerylExactLookup.from_kmers(ctx, k, rk, rv)
    A = api.MerylExactLookup.from_kmers(ctx, k, ak, ac)
    assert A.info.canonical == (0 if noncanon else 1) and R.info.canonical == (0 if noncanon else 1)
    exp, per = _expected_pairs(recs, k, _ArrLookup(k, rk, rv), _ArrLookup(k, ak, ac))
    res = api.hist(ctx, R, A, recs, api.KModel(12.0))
    got = {(int(p["v_read"]), int(p["v_asm"])): int(p["count"]) for p in res.pairs}
    assert got == exp
    if not noncanon:
        R.fuse(A)
        ji = R.joint_info
        assert ji.fused and ji.n_home + ji.n_spill + ji.n_stash == ji.n_kmers == len(np.union1d(rk, ak))
        q = np.concatenate([rk[::7], ak[::5], np.array([0, 1, (1 << (2 * k)) - 1], dtype=np.uint64)])
        jr, ja = R.joint_values(q)
        assert (jr == R.values(q)).all() and (ja == A.values(q)).all()
        res = api.hist(ctx, R, A, recs, api.KModel(12.0))
        got = {(int(p["v_read"]), int(p["v_asm"])): int(p["count"]) for p in res.pairs}
        assert got == exp
    assert [(int(c["kasm"]), int(c["kmissing"])) for c in res.per_seq] == per
    assert any(key[0] >= 4096 for key in got)          # the sparse side table was exercised

    # resident + sharded: shards add up to the same accumulator
    seq = api.SeqSet(ctx, recs)
    acc = api.Accum(ctx, seq.nseq)
    for s in range(3):
        api.hist_resident(ctx, R, A, seq, None, acc, s, 3)
    per2, pairs2 = acc.fetch()
    got2 = {(int(p["v_read"]), int(p["v_asm"])): int(p["count"]) for p in pairs2}
    assert got2 == exp
    assert [(int(c["kasm"]), int(c["kmissing"])) for c in per2] == per
    R.close(); A.close(); seq.close(); acc.close()


def test_joint_table_spill_and_saturation(ctx, monkeypatch):
    """Fused table under pressure: lines overflow into the side table, both value fields saturate."""
    from merfin_b200 import api
    monkeypatch.setenv("MFN_JOINT_LOAD", "0.97")
    rng = np.random.default_rng(17)
    for k in (21, 31):
        n = 400000
        keys = np.unique(rng.integers(0, (1 << (2 * k)) - 1, size=n, dtype=np.uint64, endpoint=True))
        keys = np.unique(np.minimum(keys, _np_revcomp(keys, k)))
        rk = keys[rng.random(len(keys)) < 0.8]
        ak = keys[rng.random(len(keys)) < 0.6]
        rv = rng.integers(1, 40, size=len(rk)).astype(np.uint32)
        av = rng.integers(1, 4, size=len(ak)).astype(np.uint32)
        rv[::50] = rng.integers(1 << 21, 0xFFFFFFFF, size=len(rv[::50]), dtype=np.uint64).astype(np.uint32)
        av[::60] = rng.integers(70, 1 << 30, size=len(av[::60]), dtype=np.uint64).astype(np.uint32)
        R = api.MerylExactLookup.from_kmers(ctx, k, rk, rv)
        A = api.MerylExactLookup.from_kmers(ctx, k, ak, av)
        R.fuse(A)
        ji = R.joint_info
        assert ji.n_kmers == len(np.union1d(rk, ak))
        assert ji.n_spill > 0 and ji.n_stash > 0 and ji.n_bigvalue > 0
        q = np.concatenate([keys, rng.integers(0, (1 << (2 * k)) - 1, size=50000, dtype=np.uint64, endpoint=True)])
        jr, ja = R.joint_values(q)
        assert (jr == R.values(q)).all() and (ja == A.values(q)).all()
        R.close(); A.close()


def test_k_mismatch_is_an_error(ctx, golden_dir):
    from merfin_b200 import api
    from merfin_b200.lib import MfnError, MFN_ERR_K_MISMATCH
    R = api.MerylExactLookup.load(ctx, os.path.join(golden_dir, "g1_k21", "reads.meryl"))
    A = api.MerylExactLookup.load(ctx, os.path.join(golden_dir, "g3_k31", "asm.meryl"))
    with pytest.raises(MfnError) as e:
        api.hist(ctx, R, A, [b"ACGT" * 20], api.KModel(5.0))
    assert e.value.status == MFN_ERR_K_MISMATCH
    with pytest.raises(MfnError):
        api.MerylExactLookup.load(ctx, "/nonexistent/db.meryl")
    R.close(); A.close()


# ---------------------------------------------------------------------------------- sweep (range-partitioned scan)
@pytest.mark.parametrize("k", [21, 16, 6, 31, 15, 32, 11])
def test_sweep_equals_direct_scan(ctx, monkeypatch, k):
    """The range-partitioned -hist scan gives the same joint histogram and per-record counters as
    the direct scan (forced on here; it is chosen automatically for large batches)."""
    from merfin_b200 import api
    recs, (ak, ac), (rk, rv) = _synthetic_case(k, 1_500_000 if k > 6 else 300_000, 200 + k)
    # many tiny records in one chunk (record-delta side list) and a long tail of records
    rng = np.random.default_rng(k)
    recs += [(f"t{i}", bytes(rng.choice(list(b"ACGT"), size=int(rng.integers(k, 3 * k))).astype(np.uint8))) for i in range(400)]
    R = api.MerylExactLookup.from_kmers(ctx, k, rk, rv)
    A = api.MerylExactLookup.from_kmers(ctx, k, ak, ac)
    R.fuse(A)
    km = api.KModel(12.0, [0, 0, 1], [0.5, 0.6, 0.7])        # readK == 0 for v_read 1 and 2 too
    monkeypatch.setenv("MFN_SWEEP", "0")
    d = api.hist(ctx, R, A, recs, km)
    monkeypatch.setenv("MFN_SWEEP", "1")
    s = api.hist(ctx, R, A, recs, km)
    assert len(d.pairs) == len(s.pairs) and (d.pairs == s.pairs).all()
    assert (d.per_seq == s.per_seq).all()
    assert int(d.per_seq["kmissing"].sum()) > 0
    # the host-buffer path cut into several sub-sweeps (probe of one under the upload of the next)
    monkeypatch.setenv("MFN_SUBSWEEPS", "3")
    s3 = api.hist(ctx, R, A, recs, km)
    monkeypatch.delenv("MFN_SUBSWEEPS")
    assert len(d.pairs) == len(s3.pairs) and (d.pairs == s3.pairs).all() and (d.per_seq == s3.per_seq).all()
    exp, per = _expected_pairs(recs, k, _ArrLookup(k, rk, rv), _ArrLookup(k, ak, ac))
    got = {(int(p["v_read"]), int(p["v_asm"])): int(p["count"]) for p in s.pairs}
    assert got == exp
    assert [int(c["kasm"]) for c in s.per_seq] == [p[0] for p in per]
    # the fused table restricted to the k-mers of the scanned sequences (what the sweep streams when it
    # exists): same result; and restricted to HALF of them, so that the other half of the windows (and
    # a record of k-mers neither database has) must fall back to the per-database tables
    seq = api.SeqSet(ctx, recs)
    R.restrict(A, seq)
    ri = R.restricted_info
    assert ri.fused and 0 < ri.n_kmers <= R.joint_info.n_kmers
    r1 = api.hist(ctx, R, A, recs, km)
    assert len(d.pairs) == len(r1.pairs) and (d.pairs == r1.pairs).all() and (d.per_seq == r1.per_seq).all()
    R.restrict(A, seq, 0, 2)
    # (with k = 6 either half already holds every one of the 2080 canonical 6-mers)
    assert R.restricted_info.n_kmers < ri.n_kmers or (k <= 8 and R.restricted_info.n_kmers == ri.n_kmers)
    extra = recs + [("novel", bytes(rng.choice(list(b"ACGT"), size=50_000).astype(np.uint8)))]
    monkeypatch.setenv("MFN_SWEEP", "0")
    d2 = api.hist(ctx, R, A, extra, km)
    monkeypatch.setenv("MFN_SWEEP", "1")
    r2 = api.hist(ctx, R, A, extra, km)
    assert len(d2.pairs) == len(r2.pairs) and (d2.pairs == r2.pairs).all() and (d2.per_seq == r2.per_seq).all()
    R.unrestrict()
    assert not R.restricted_info.fused
    seq.close()
    R.close(); A.close()


def test_sweep_on_golden_files(ctx, golden_dir, monkeypatch):
    monkeypatch.setenv("MFN_SWEEP", "1")
    for name, k, peak in SETS[:2]:
        d, recs = _golden(golden_dir, name)
        _hist_check(ctx, d, recs, k, peak, 0, 0xFFFFFFFF, None, "ref")
    d, recs = _golden(golden_dir, "g1_k21")
    _hist_check(ctx, d, recs, 21, 8.0, 0, 0xFFFFFFFF, os.path.join(d, "prob.csv"), "ref.prob")


def test_host_packing_lane_equals_device_packing(ctx, monkeypatch):
    """mfn_hist over host buffers with half of the pieces packed by the host threads (ingest.cu,
    host_pack.cpp) and half uploaded raw gives exactly the result of the all-device path."""
    from merfin_b200 import api
    k = 21
    rng = np.random.default_rng(77)
    recs = []
    for i, n in enumerate([3_000_001, 2_500_000, 17, 4_000_003, 1_999_999]):      # ~11.5 Mbp: 6 pieces of 1024 tiles
        b = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n)
        b[rng.integers(0, n, size=max(1, n // 5000))] = ord("N")
        if n > 1000:
            b[100:200] |= 0x20                                                       # a lowercase run
        recs.append((f"r{i}", b.tobytes()))
    seq = api.SeqSet(ctx, recs)
    A = api.MerylExactLookup.count_sequence(ctx, k, seq)
    R = api.MerylExactLookup.synth_reads(ctx, A, 15.0, 1_000_000, 5)
    R.fuse(A)
    km = api.KModel(15.0)
    monkeypatch.setenv("MFN_SWEEP", "1")
    monkeypatch.setenv("MFN_HOST_PACK_SHARE", "0")
    d = api.hist(ctx, R, A, recs, km)
    monkeypatch.setenv("MFN_HOST_PACK_SHARE", "0.5")
    monkeypatch.setenv("MFN_HOST_PACK_THREADS", "3")
    h = api.hist(ctx, R, A, recs, km)
    monkeypatch.setenv("MFN_HOST_PACK_SHARE", "0.9")
    monkeypatch.setenv("MFN_SUBSWEEPS", "2")
    h2 = api.hist(ctx, R, A, recs, km)
    for x in (h, h2):
        assert len(d.pairs) == len(x.pairs) and (d.pairs == x.pairs).all() and (d.per_seq == x.per_seq).all()
    assert int(d.per_seq["kasm"].sum()) > 11_000_000
    seq.close(); R.close(); A.close()


@pytest.mark.parametrize("name", ["g1_k21", "g2_k6", "g3_k31"])
def test_gpu_decode_equals_host_decode(ctx, golden_dir, monkeypatch, name):
    """The device decoder of the on-disk meryl blocks (meryl_gpu.cu: one thread per block walks the
    unary/binary chain) yields exactly the (k-mer, value) pairs of the host decoder, which the
    CPU tests pin to `meryl print`."""
    from merfin_b200 import api
    for db in ("reads.meryl", "asm.meryl"):
        path = os.path.join(golden_dir, name, db)
        monkeypatch.setenv("MFN_HOST_DECODE", "1")
        H = api.MerylExactLookup.load(ctx, path)
        monkeypatch.delenv("MFN_HOST_DECODE")
        G = api.MerylExactLookup.load(ctx, path)
        hk, hv = H.export()
        gk, gv = G.export()
        ho, go = np.argsort(hk), np.argsort(gk)
        assert len(hk) == len(gk) > 0
        assert (hk[ho] == gk[go]).all() and (hv[ho] == gv[go]).all()
        H.close(); G.close()


def test_gpu_decode_written_database(ctx, tmp_path):
    """A larger database with many blocks per file, written by mfn_meryl_write (which the
    reference reads back in the CPU tests), decoded on the device."""
    from merfin_b200 import api
    k = 21
    rng = np.random.default_rng(11)
    km = np.unique(rng.integers(0, 1 << 42, size=3_000_000, dtype=np.uint64))
    vv = rng.integers(1, 70000, size=len(km), dtype=np.uint32)
    vv[::1000] = 0xFFFFFFF0                                   # values that need the excess table
    d = str(tmp_path / "w.meryl")
    api.meryl_write(d, k, km, vv)
    T = api.MerylExactLookup.load(ctx, d)
    gk, gv = T.export()
    o = np.argsort(gk)
    assert (gk[o] == km).all() and (gv[o] == vv).all()
    T.close()


def _run_reference_hist(workdir, args):
    import subprocess
    ref = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "merfin")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/merfin not built (reference sources absent)")
    p = subprocess.run([ref] + args, cwd=workdir, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stderr


@pytest.mark.parametrize("k,mn,prob", [(21, 0, False), (21, 2, True), (31, 0, False)])
def test_midsize_hist_against_reference_binary(ctx, tmp_path, monkeypatch, k, mn, prob):
    """12 Mbp through every fast path at once (host-buffer call, both ingest lanes, sub-sweeps,
    range-partitioned scan, device decode of the databases) against the UNMODIFIED reference
    binary run here on the same FASTA and the same on-disk meryl databases: -hist file byte for
    byte, missing / found counts exactly, the printed QV lines as printed."""
    from merfin_b200 import api, synth
    recs = [(n, b.tobytes()) for n, b in synth.make_assembly(12_000_000, 7, 100 + k, dup_len=5000)]
    seq = api.SeqSet(ctx, recs)
    A0 = api.MerylExactLookup.count_sequence(ctx, k, seq)
    R0 = api.MerylExactLookup.synth_reads(ctx, A0, 12.0, 4_000_000, 9)
    d = str(tmp_path)
    with open(os.path.join(d, "asm.fa"), "wb") as f:
        for n, b in recs:
            f.write(b">" + n.encode() + b" some description\n")
            for i in range(0, len(b), 100000):
                f.write(b[i:i + 100000] + b"\n")
    for t, name in ((A0, "asm.meryl"), (R0, "reads.meryl")):
        km_, vv_ = t.export()
        api.meryl_write(os.path.join(d, name), k, km_, vv_)
    seq.close(); A0.close(); R0.close()
    args = ["-hist", "-sequence", "asm.fa", "-seqmers", "asm.meryl", "-readmers", "reads.meryl", "-peak", "12",
            "-threads", "8", "-memory", "32", "-output", "ref.hist"]
    probfile = None
    if prob:
        probfile = os.path.join(d, "prob.csv")
        with open(probfile, "w") as f:      # readK 0 for multiplicities 1-2, then K = round(v/12) with decreasing confidence
            for v in range(1, 120):
                f.write("%d,%.3f\n" % (0 if v <= 2 else max(1, round(v / 12)), max(0.05, 1.0 - v / 150)))
        args += ["-prob", "prob.csv"]
    if mn:
        args += ["-min", str(mn)]
    err = _run_reference_hist(d, args)
    want_file = open(os.path.join(d, "ref.hist"), "rb").read()
    import re
    want_summary = "\n".join(l for l in err.splitlines() if re.match(r"^(K-mers|Missing QV|Merfin QV)", l)) + "\n"

    R = api.MerylExactLookup.load(ctx, os.path.join(d, "reads.meryl"), mn, 0xFFFFFFFF)
    A = api.MerylExactLookup.load(ctx, os.path.join(d, "asm.meryl"))
    R.fuse(A)
    km = api.KModel.load_Kmetric(12.0, probfile)
    monkeypatch.setenv("MFN_SWEEP", "1")
    monkeypatch.setenv("MFN_SUBSWEEPS", "2")
    monkeypatch.setenv("MFN_HOST_PACK_SHARE", "0.5")
    monkeypatch.setenv("MFN_HOST_PACK_THREADS", "4")
    res = api.hist(ctx, R, A, recs, km)
    assert res.file_text() == want_file
    assert res.summary_text() == want_summary
    monkeypatch.setenv("MFN_SWEEP", "0")
    res2 = api.hist(ctx, R, A, recs, km)
    assert (res.pairs == res2.pairs).all() and (res.per_seq == res2.per_seq).all()
    R.close(); A.close()


@pytest.mark.parametrize("name,peak", [("g1_k21", 8), ("g2_k6", 3), ("g3_k31", 6)])
def test_completeness_over_resident_tables(ctx, golden_dir, name, peak):
    """The join over two resident tables (mfn_completeness_resident) gives the pairs of the join over the databases
    on disk, hence the reference's TOTAL / COMPLETENESS lines."""
    from merfin_b200 import api
    d = os.path.join(golden_dir, name)
    km = api.KModel(peak)
    t0, u0, pairs0 = api.completeness(ctx, os.path.join(d, "reads.meryl"), os.path.join(d, "asm.meryl"), km)
    R = api.MerylExactLookup.load(ctx, os.path.join(d, "reads.meryl"))
    A = api.MerylExactLookup.load(ctx, os.path.join(d, "asm.meryl"))
    t1, u1, pairs1 = api.completeness_resident(ctx, R, A, km)
    assert len(pairs0) == len(pairs1) and (pairs0 == pairs1).all() and (t0, u0) == (t1, u1)
    assert open(os.path.join(d, "ref.completeness")).read().endswith(api.completeness_text(t1, u1))
    R.close(); A.close()


def test_completeness_shards_sum_to_the_whole(ctx, golden_dir):
    """-completeness cut into shards of read-database files (what N ranks do before the all-reduce)
    accumulates exactly the joint histogram of the single-call version."""
    import ctypes as C
    from merfin_b200 import api, lib as L
    lib = L.load()
    d = os.path.join(golden_dir, "g1_k21")
    _, _, whole = api.completeness(ctx, os.path.join(d, "reads.meryl"), os.path.join(d, "asm.meryl"), api.KModel(8.0))
    A = api.MerylExactLookup.load(ctx, os.path.join(d, "asm.meryl"))          # raw values
    acc = api.Accum(ctx, 0)
    for s in range(3):
        L.check(lib.mfn_completeness_shard(ctx.h, os.path.join(d, "reads.meryl").encode(), A.h, s, 3, acc.h))
    _, pairs = acc.fetch()
    key = lambda p: sorted((int(x["v_read"]), int(x["v_asm"]), int(x["count"])) for x in p)
    assert key(pairs) == key(whole) and len(pairs) > 3
    acc.close(); A.close()


@pytest.mark.parametrize("name", ["g1_k21", "g2_k6"])
def test_streamed_read_database(ctx, golden_dir, name):
    """A read database streamed through the device into the restricted table (never resident: the answer to a
    database bigger than the HBM) gives the -hist result of the resident tables, whole and in two shards,
    also with -min/-max; -dump and path scoring refuse the handle."""
    from merfin_b200 import api, lib as L
    d = os.path.join(golden_dir, name)
    recs = O.read_fasta(os.path.join(d, "asm.fa"))
    km = api.KModel(8.0)
    A = api.MerylExactLookup.load(ctx, os.path.join(d, "asm.meryl"))
    seq = api.SeqSet(ctx, recs)
    for minV, maxV in ((0, 0xFFFFFFFF), (2, 12)):
        R = api.MerylExactLookup.load(ctx, os.path.join(d, "reads.meryl"), minV, maxV)
        want = api.hist(ctx, R, A, recs, km)
        S, st = api.MerylExactLookup.restrict_stream(ctx, os.path.join(d, "reads.meryl"), A, seq, minV=minV, maxV=maxV)
        assert st.db_kmers == api.meryl_info(os.path.join(d, "reads.meryl"))["n_distinct"] and 0 < st.kept_kmers <= st.shard_kmers
        acc = api.Accum(ctx, len(recs))
        api.hist_resident(ctx, S, A, seq, km, acc)
        per, pairs = acc.fetch()
        assert len(pairs) == len(want.pairs) and (pairs == want.pairs).all() and (per == want.per_seq).all()
        # two shards, each with the table of its own half of the windows
        acc.reset()
        for sh in range(2):
            S2, _ = api.MerylExactLookup.restrict_stream(ctx, os.path.join(d, "reads.meryl"), A, seq, sh, 2, minV=minV, maxV=maxV)
            api.hist_resident(ctx, S2, A, seq, km, acc, sh, 2)
            ctx.sync()
            S2.close()
        per, pairs = acc.fetch()
        assert len(pairs) == len(want.pairs) and (pairs == want.pairs).all() and (per == want.per_seq).all()
        with pytest.raises(L.MfnError) as e:
            api.dump(ctx, S, A, recs)
        assert e.value.status == L.MFN_ERR_UNSUPPORTED
        acc.close(); S.close(); R.close()
    seq.close(); A.close()


@pytest.mark.parametrize("k", [21, 31, 16])
def test_streamed_synthetic_database_equals_resident(ctx, k):
    """mfn_tables_restrict_synth (the benchmark's database of any size, error k-mers generated in chunks and pushed
    through the semi-join) against the same synthetic database built as a resident table: same -hist result
    (k = 31: 64-bit slots and wide entries; k = 16: k-mers that are their own reverse complement)."""
    from merfin_b200 import api
    rng = np.random.default_rng(8)
    base = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=1_500_000)
    recs = [("a", base[:800_000].tobytes()), ("b", base[700_000:].tobytes()), ("n", b"ACGTN" * 2000)]
    seq = api.SeqSet(ctx, recs)
    A = api.MerylExactLookup.count_sequence(ctx, k, seq)
    R = api.MerylExactLookup.synth_reads(ctx, A, 12.0, 3_000_000, 5)
    km = api.KModel(12.0)
    want = api.hist(ctx, R, A, recs, km)
    S, st = api.MerylExactLookup.restrict_synth(ctx, A, seq, 12.0, 3_000_000, 5, chunk_kmers=400_000)
    assert st.db_kmers == 3_000_000 + A.nKmers() and st.shard_kmers == A.nKmers()
    acc = api.Accum(ctx, len(recs))
    api.hist_resident(ctx, S, A, seq, km, acc)
    per, pairs = acc.fetch()
    assert len(pairs) == len(want.pairs) and (pairs == want.pairs).all() and (per == want.per_seq).all()
    acc.close(); S.close(); R.close(); A.close(); seq.close()


def _big_records(rng):
    """110 Mbp in records that exercise every path of the k = 21 scatter kernel at a size where the tables have
    32-bit slots (>= 2^22 lines): N runs, lower case, empty / short / all-N records, and a crowd of tiny records
    inside one chunk (more records than the staging word tells apart: side list)."""
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    a = rng.choice(acgt, size=80_000_000)
    a[500_000:500_700] = ord("N")
    a[900_000:900_050] += 32                               # lower case
    b = rng.choice(acgt, size=30_000_000)
    b[100:130] = ord("N")
    tiny = [("t%d" % i, rng.choice(acgt, size=int(rng.integers(25, 90))).tobytes()) for i in range(60)]
    return ([("a", a.tobytes()), ("empty", b""), ("short", b"ACGTACGTAC")] + tiny +
            [("b", b.tobytes()), ("dup", a[:300_000].tobytes()), ("n", b"N" * 5000)])


@pytest.mark.parametrize("restricted", [False, True])
def test_scatter3_hist_equals_direct_scan(ctx, monkeypatch, restricted):
    """-hist through k_sweep_scatter3 (k = 21, narrow entries: needs tables with 32-bit slots, i.e. >= 70 M k-mers)
    against the direct scan: joint histogram and per-record counters, full fused table (division geometry) and
    restricted table (bit-field geometry); also in two shards and through the host-buffer path."""
    from merfin_b200 import api
    recs = _big_records(np.random.default_rng(22))
    seq = api.SeqSet(ctx, recs)
    A = api.MerylExactLookup.count_sequence(ctx, 21, seq)
    R = api.MerylExactLookup.synth_reads(ctx, A, 10.0, 2_000_000, 9)
    R.fuse(A)
    assert R.joint_info.slot_bits == 32
    if restricted:
        R.restrict(A, seq)
    km = api.KModel(10.0)
    acc = api.Accum(ctx, len(recs))
    monkeypatch.setenv("MFN_SWEEP", "0")
    api.hist_resident(ctx, R, A, seq, km, acc)
    per0, pairs0 = acc.fetch()
    monkeypatch.setenv("MFN_SWEEP", "1")
    acc.reset()
    l0 = ctx.launches
    api.hist_resident(ctx, R, A, seq, km, acc)
    assert ctx.launches - l0 >= 3                            # scatter, probe, side list
    per1, pairs1 = acc.fetch()
    assert len(pairs0) == len(pairs1) and (pairs0 == pairs1).all() and (per0 == per1).all()
    acc.reset()
    for sh in range(2):
        api.hist_resident(ctx, R, A, seq, km, acc, sh, 2)
    per2, pairs2 = acc.fetch()
    assert len(pairs0) == len(pairs2) and (pairs0 == pairs2).all() and (per0 == per2).all()
    h = api.hist(ctx, R, A, recs, km)                        # host buffers: both ingest lanes, sub-sweeps
    assert len(pairs0) == len(h.pairs) and (pairs0 == h.pairs).all() and (per0 == h.per_seq).all()
    acc.close(); seq.close(); R.close(); A.close()


@pytest.mark.parametrize("restricted", [False, True])
def test_dump_through_the_range_partitioned_scan(ctx, monkeypatch, restricted):
    """-dump's pairs per position through scatter -> probe (entries overwritten by their pairs) -> gather equal the
    direct scan's, for records with N runs, lower case, short and empty records; with the full fused table
    (division geometry) and with the restricted one (bit-field geometry)."""
    from merfin_b200 import api
    recs = _big_records(np.random.default_rng(21))
    seq = api.SeqSet(ctx, recs)
    A = api.MerylExactLookup.count_sequence(ctx, 21, seq)
    R = api.MerylExactLookup.synth_reads(ctx, A, 10.0, 2_000_000, 9)
    R.fuse(A)
    assert R.joint_info.slot_bits == 32
    if restricted:
        R.restrict(A, seq)
    monkeypatch.setenv("MFN_DUMP_SWEEP", "0")
    l0 = ctx.launches
    _, _, vr0, va0, ok0 = api.dump(ctx, R, A, recs)
    direct_launches = ctx.launches - l0
    monkeypatch.setenv("MFN_DUMP_SWEEP", "1")
    l0 = ctx.launches
    _, _, vr1, va1, ok1 = api.dump(ctx, R, A, recs)
    assert ctx.launches - l0 > direct_launches              # scatter + probe + gather ran
    assert (ok0 == ok1).all() and (vr0 == vr1).all() and (va0 == va1).all()
    assert int(ok1.sum()) > 100_000_000
    seq.close(); R.close(); A.close()


def test_sweep_overflow_falls_back_to_the_direct_scan(ctx, monkeypatch):
    """A batch dominated by one k-mer (5 Mbp of poly-A) sends nearly every window to a single
    bin; the range-partitioned scan runs out of room there and mfn_hist redoes the batch with the
    direct scan instead of failing."""
    from merfin_b200 import api
    rng = np.random.default_rng(4)
    rnd = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=1_000_000).tobytes()
    recs = [("polyA", b"A" * 5_000_000), ("rnd", rnd), ("polyA2", b"a" * 1_000_000)]
    seq = api.SeqSet(ctx, recs)
    A = api.MerylExactLookup.count_sequence(ctx, 21, seq)
    R = api.MerylExactLookup.synth_reads(ctx, A, 10.0, 100_000, 3)
    R.fuse(A)
    km = api.KModel(10.0)
    monkeypatch.setenv("MFN_SWEEP", "0")
    d = api.hist(ctx, R, A, recs, km)
    monkeypatch.setenv("MFN_SWEEP", "1")
    s = api.hist(ctx, R, A, recs, km)
    assert len(d.pairs) == len(s.pairs) and (d.pairs == s.pairs).all() and (d.per_seq == s.per_seq).all()
    assert int(s.per_seq["kasm"][0]) == 5_000_000 - 20
    # the entry points that return before the device is done (mfn_hist_resident / mfn_hist_host) report the
    # loss at fetch; the caller recovers with mfn_ctx_set_sweep(ctx, 0) and the same call again
    from merfin_b200 import lib as L
    acc = api.Accum(ctx, len(recs))
    api.hist_resident(ctx, R, A, seq, km, acc)
    with pytest.raises(L.MfnError) as e:
        acc.fetch()
    assert e.value.status == L.MFN_ERR_CAPACITY
    ctx.set_sweep(0)
    acc.reset()
    api.hist_resident(ctx, R, A, seq, km, acc)
    ctx.set_sweep(-1)
    per, pairs = acc.fetch()
    assert len(pairs) == len(d.pairs) and (pairs == d.pairs).all() and (per == d.per_seq).all()
    acc.close(); seq.close(); R.close(); A.close()


def test_cli_dump_in_batches(golden_dir, tmp_path):
    """The CLI's -dump with one record per batch (MFN_DUMP_BATCH=1): every batch after the first is
    computed on a worker thread while the previous one is formatted (host/dump_driver.cpp); same file
    as the reference's.  (Kept last: it is the only GPU test of the multi-batch path.)"""
    import gzip
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    merfin = os.path.join(root, "merfin_b200", "bin", "merfin")
    for name, peak in (("g1_k21", 8), ("g3_k31", 6)):
        d = os.path.join(golden_dir, name)
        args = [merfin, "-dump", "-sequence", os.path.join(d, "asm.fa"), "-seqmers", os.path.join(d, "asm.meryl"),
                "-readmers", os.path.join(d, "reads.meryl"), "-peak", str(peak), "-output", "b.dump"]
        p = subprocess.run(args, cwd=str(tmp_path), capture_output=True, text=True, timeout=150, env=dict(os.environ, MFN_DUMP_BATCH="1"))
        assert p.returncode == 0, p.stderr[-2000:]
        assert open(tmp_path / "b.dump", "rb").read() == gzip.open(os.path.join(d, "ref.dump.gz"), "rb").read()
        assert open(os.path.join(d, "ref.dump.perseq")).read() in p.stderr
