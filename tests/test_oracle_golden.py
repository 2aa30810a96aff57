"""Pin the CPU oracle against outputs of the unmodified reference binaries (tests/golden/)."""
import gzip
import os
import re

import numpy as np
import pytest

import oracle_lib as O

SETS = [("g1_k21", 21, 8.0), ("g2_k6", 6, 3.0), ("g3_k31", 31, 6.0), ("g4_k40", 40, 6.0), ("g5_k64", 64, 6.0)]


def _load(golden_dir, name):
    d = os.path.join(golden_dir, name)
    S = O.DB(os.path.join(d, "reads.meryl"))
    A = O.DB(os.path.join(d, "asm.meryl"))
    recs = O.read_fasta(os.path.join(d, "asm.fa"))
    return d, S, A, recs


@pytest.mark.parametrize("name,k,peak", SETS)
def test_db_decode_matches_meryl_print(golden_dir, name, k, peak):
    d, S, A, _ = _load(golden_dir, name)
    for db, fn in ((S, "reads.print.gz"), (A, "asm.print.gz")):
        assert db.k == k
        lines = gzip.open(os.path.join(d, fn), "rt").read().split("\n")
        lines = [l for l in lines if l]
        assert db.n == len(lines) == db.numDistinct
        exp_k = np.array([O.kmer_from_string(l.split("\t")[0]) for l in lines], dtype=object)
        exp_v = np.array([int(l.split("\t")[1]) for l in lines], dtype=np.uint32)
        got_k = np.array([(int(h) << 64) | int(l) for h, l in zip(db.hi, db.lo)], dtype=object)
        assert (got_k == exp_k).all()
        assert (db.val == exp_v).all()
        # master-index statistics agree with the data
        vals, cnts = np.unique(db.val, return_counts=True)
        assert (db.histValue == vals).all() and (db.histCount == cnts).all()
        assert db.numTotal == int(db.val.astype(np.uint64).sum())


@pytest.mark.parametrize("name,k,peak", SETS)
def test_hist_matches_reference(golden_dir, name, k, peak):
    d, S, A, recs = _load(golden_dir, name)
    R, Am = O.Lookup(S), O.Lookup(A)
    km = O.KModel(peak)
    h = O.Hist()
    per = []
    cum = 0
    for ident, b in recs:
        kasm, kmiss, _ = h.add_record(R, Am, km, b)
        cum += kmiss
        per.append("%s\t%d\t%d\t%d\t%.2f" % (ident, kmiss, cum, kasm, O.histoQV(kmiss, kasm, k)))
    assert h.report() == open(os.path.join(d, "ref.hist"), "rb").read()
    assert h.summary(k) == open(os.path.join(d, "ref.hist.summary")).read()
    assert "\n".join(per) + "\n" == open(os.path.join(d, "ref.hist.perseq")).read()


def test_hist_min_max_filter(golden_dir):
    d, S, A, recs = _load(golden_dir, "g1_k21")
    for tag, peak, mn, mx in (("min3", 1.0, 3, 0xFFFFFFFF), ("min2max12", 8.0, 2, 12)):
        R, Am = O.Lookup(S, mn, mx), O.Lookup(A)
        km = O.KModel(peak)
        h = O.Hist()
        for _, b in recs:
            h.add_record(R, Am, km, b)
        assert h.report() == open(os.path.join(d, f"ref.{tag}.hist"), "rb").read()
        assert h.summary(21) == open(os.path.join(d, f"ref.{tag}.hist.summary")).read()


def test_hist_dump_completeness_with_prob_table(golden_dir):
    d, S, A, recs = _load(golden_dir, "g1_k21")
    R, Am = O.Lookup(S), O.Lookup(A)
    km = O.KModel.from_file(8.0, os.path.join(d, "prob.csv"))
    assert len(km.K) == 7          # the malformed line is skipped
    h = O.Hist()
    per, cum = [], 0
    for ident, b in recs:
        kasm, kmiss, _ = h.add_record(R, Am, km, b)
        cum += kmiss
        per.append("%s\t%d\t%d\t%d\t%.2f" % (ident, kmiss, cum, kasm, O.histoQV(kmiss, kasm, 21)))
    assert h.report() == open(os.path.join(d, "ref.prob.hist"), "rb").read()
    assert h.summary(21) == open(os.path.join(d, "ref.prob.hist.summary")).read()
    assert "\n".join(per) + "\n" == open(os.path.join(d, "ref.prob.hist.perseq")).read()
    text = b"".join(O.dump_record(R, Am, km, ident, b)[0] for ident, b in recs)
    assert text == gzip.open(os.path.join(d, "ref.prob.dump.gz"), "rb").read()
    _check_completeness(d, "ref.prob.completeness", S, A, km)


@pytest.mark.parametrize("name,k,peak", SETS)
def test_dump_matches_reference(golden_dir, name, k, peak):
    d, S, A, recs = _load(golden_dir, name)
    R, Am = O.Lookup(S), O.Lookup(A)
    km = O.KModel(peak)
    out, per, cm, ca = [], [], 0, 0
    for ident, b in recs:
        text, kasm, kmiss = O.dump_record(R, Am, km, ident, b)
        out.append(text)
        cm += kmiss; ca += kasm
        per.append("%s\t%d\t%d\t%d" % (ident, kmiss, cm, ca))
    assert b"".join(out) == gzip.open(os.path.join(d, "ref.dump.gz"), "rb").read()
    assert "\n".join(per) + "\n" == open(os.path.join(d, "ref.dump.perseq")).read()


def _check_completeness(d, fn, S, A, km):
    t64, u64, t, u = O.completeness(S, A, km)
    ref = open(os.path.join(d, fn)).read()
    for line in ref.splitlines():
        m = re.match(r"thread\s+(\d+) total\s+([\d.]+) underc\s+([\d.]+) completeness", line)
        if m:
            i = int(m.group(1))
            assert "%12.2f" % t64[i] == "%12.2f" % float(m.group(2))
            assert "%15.5f" % u64[i] == "%15.5f" % float(m.group(3))
    tail = ("TOTAL readK:   %15.2f\nTOTAL undrcpy:    %15.5f\nCOMPLETENESS:             %0.5f\n"
            % (t, u, 1.0 - u / t))
    assert ref.endswith(tail)


@pytest.mark.parametrize("name,k,peak", SETS)
def test_completeness_matches_reference(golden_dir, name, k, peak):
    d, S, A, _ = _load(golden_dir, name)
    _check_completeness(d, "ref.completeness", S, A, O.KModel(peak))


def test_palindromes_count_double(golden_dir):
    """Q2: even k, f == r -> both strand probes hit the same entry."""
    d, S, A, _ = _load(golden_dir, "g2_k6")
    R = O.Lookup(S)
    pal = O.kmer_from_string("ACGCGT")
    v = R.value(pal)
    vr, va, ok = O.pairs(R, O.Lookup(A), b"ACGCGT")
    assert ok[0] == 1 and vr[0] == 2 * v


def test_window_iteration_edges():
    f, _, r, _, ok = O.windows(b"ACGTNACGTACGT", 4)
    assert len(ok) == 10
    assert list(ok) == [1, 0, 0, 0, 0, 1, 1, 1, 1, 1]
    assert f[0] == O.kmer_from_string("ACGT") and r[0] == O.kmer_from_string("ACGT")
    assert f[5] == O.kmer_from_string("ACGT") and f[6] == O.kmer_from_string("CGTA")
    assert r[6] == O.kmer_from_string("TACG")
    assert len(O.windows(b"ACG", 4)[4]) == 0
    assert len(O.windows(b"", 4)[4]) == 0
    # lower case is valid and encodes like upper case
    assert O.windows(b"acgt", 4)[0][0] == O.kmer_from_string("ACGT")
    # only ACGTacgt are bases: IUPAC codes break windows
    assert list(O.windows(b"ACGTRACGT", 4)[4]) == [1, 0, 0, 0, 0, 1]
