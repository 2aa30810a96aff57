"""Host-side variant logic of the product (VCF parsing, clustering, path enumeration, score replay,
selectors: merfin_b200/csrc/host/variants.cpp) against the reference's golden VCFs, with the CPU
oracle standing in for the GPU probes (tests/variant_harness.cpp).  Runs without a GPU."""
import gzip
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "build", "variant_harness")


@pytest.fixture(scope="module")
def harness():
    srcs = [os.path.join(ROOT, "tests", "variant_harness.cpp"),
            os.path.join(ROOT, "merfin_b200", "csrc", "host", "variants.cpp"),
            os.path.join(ROOT, "merfin_b200", "csrc", "host", "fasta.cpp"),
            os.path.join(ROOT, "oracle", "mfn_oracle.c")]
    so_dir = os.path.join(ROOT, "merfin_b200")
    os.makedirs(os.path.dirname(HARNESS), exist_ok=True)
    obj = os.path.join(ROOT, "build", "mfn_oracle_h.o")
    subprocess.check_call(["gcc", "-O2", "-w", "-std=gnu11", "-c", srcs[3], "-o", obj])
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-w", "-o", HARNESS, *srcs[:3], obj,
                           "-L" + so_dir, "-lmerfin_b200", "-Wl,-rpath," + so_dir, "-lm"])
    return HARNESS


def _run(harness, d, mode, peak, *extra, cwd=None):
    p = subprocess.run([harness, os.path.join(d, "asm.fa"), os.path.join(d, "reads.meryl"), os.path.join(d, "asm.meryl"),
                        os.path.join(d, "calls.vcf"), mode, str(peak), *extra], capture_output=True, text=True, cwd=cwd)
    assert p.returncode == 0, p.stderr
    return p.stdout


@pytest.mark.parametrize("mode", ["polish", "filter", "better", "strict", "loose"])
def test_variant_modes_match_reference(harness, golden_dir, mode):
    d = os.path.join(golden_dir, "g1_k21")
    assert _run(harness, d, mode, 8) == open(os.path.join(d, f"ref.{mode}.vcf")).read()


def test_polish_with_prob_table(harness, golden_dir):
    d = os.path.join(golden_dir, "g1_k21")
    assert _run(harness, d, "polish", 8, os.path.join(d, "prob.csv")) == open(os.path.join(d, "ref.probpolish.vcf")).read()


def test_polish_debug_log_matches_reference(harness, golden_dir, tmp_path):
    """Per-path numM, min/max/median/avg |k*| and the delta sum, for every candidate path."""
    d = os.path.join(golden_dir, "g1_k21")
    _run(harness, d, "polish", 8, "--debug", cwd=str(tmp_path))
    got = open(tmp_path / "harness.debug").read()
    assert got == gzip.open(os.path.join(d, "ref.polish.debug.gz"), "rt").read()


@pytest.mark.parametrize("extra", [("--shards", "3"), ("--batch", "2"), ("--shards", "4", "--batch", "1")])
def test_cluster_ranges_and_batches_concatenate(harness, golden_dir, tmp_path, extra):
    """A record processed as consecutive cluster ranges (the multi-GPU sharding of -polish) and in small
    scorer batches gives the same VCF and the same per-path debug records as in one piece."""
    d = os.path.join(golden_dir, "g1_k21")
    for mode in ("polish", "filter", "loose"):
        assert _run(harness, d, mode, 8, *extra) == open(os.path.join(d, f"ref.{mode}.vcf")).read()
    assert _run(harness, d, "polish", 8, os.path.join(d, "prob.csv"), *extra) == open(os.path.join(d, "ref.probpolish.vcf")).read()
    _run(harness, d, "polish", 8, "--debug", *extra, cwd=str(tmp_path))
    assert open(tmp_path / "harness.debug").read() == gzip.open(os.path.join(d, "ref.polish.debug.gz"), "rt").read()


def test_threaded_vcf_parse_keeps_file_order(harness, golden_dir, monkeypatch):
    """The VCF is parsed on several host threads; clustering starts from the file order (the
    reference's sort is not stable), so the output must not depend on the thread count."""
    d = os.path.join(golden_dir, "g1_k21")
    monkeypatch.setenv("MFN_VCF_THREADS", "5")
    assert _run(harness, d, "polish", 8) == open(os.path.join(d, "ref.polish.vcf")).read()
    assert _run(harness, d, "filter", 8) == open(os.path.join(d, "ref.filter.vcf")).read()


MERFIN = os.path.join(ROOT, "oracle", "_ref", "merfin")
MERYL = os.path.join(ROOT, "oracle", "_ref", "meryl")

LIVE = [  # k, bases per record, records, edits per record, bogus SNVs per record, peak, extra reference flags
    (21, 30000, 3, 150, 20, 10.0, []),
    (15, 20000, 2, 200, 30, 8.0, ["-comb", "3"]),
    (31, 25000, 2, 120, 10, 12.0, ["-nosplit"]),
    (18, 20000, 4, 100, 10, 9.0, ["-comb", "6", "-nosplit"]),
]


@pytest.mark.skipif(not (os.path.exists(MERFIN) and os.path.exists(MERYL)), reason="oracle/_ref not built (reference sources absent)")
@pytest.mark.parametrize("k,length,nrec,nedit,nfalse,peak,flags", LIVE)
def test_variant_modes_against_reference_on_fresh_input(harness, tmp_path, k, length, nrec, nedit, nfalse, peak, flags):
    """All five selectors against the unmodified reference run NOW (one thread, so its output order is
    the file order) on a fresh assembly with planted errors and a VCF of proposed corrections: mixed
    genotypes, clustered neighbours, multi-allelic records, 0/0 and ./. records, bogus SNVs;
    -comb / -nosplit change how clusters are split."""
    from merfin_b200 import synth
    d = str(tmp_path)
    truth = synth.make_assembly(length * nrec, nrec, 4000 + k, n_run=15, lower_run=30, dup_len=1500)
    asm, vcf = synth.plant_errors(truth, nedit, 77 + k, false_per_rec=nfalse)
    synth.write_fasta(os.path.join(d, "asm.fa"), asm, width=70)
    open(os.path.join(d, "calls.vcf"), "w").write(vcf)
    reads = synth.simulate_reads(truth, 12, 120, 0.005, 5000 + k)
    synth.write_fasta(os.path.join(d, "reads.fa"), reads, prefix="r")
    for fa, db in (("reads.fa", "reads.meryl"), ("asm.fa", "asm.meryl")):
        p = subprocess.run([MERYL, "threads=1", "count", f"k={k}", "memory=1", fa, "output", db], cwd=d, capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stderr[-2000:]
    hflags = [f.replace("-comb", "--comb").replace("-nosplit", "--nosplit") for f in flags]
    nrecords = 0
    for mode in ("polish", "filter", "better", "strict", "loose"):
        p = subprocess.run([MERFIN, "-" + mode, "-sequence", "asm.fa", "-seqmers", "asm.meryl", "-readmers", "reads.meryl",
                            "-vcf", "calls.vcf", "-peak", str(peak), "-threads", "1", "-memory", "4", "-output", "ref." + mode] + flags,
                           cwd=d, capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        want = open(os.path.join(d, "ref.%s.%s.vcf" % (mode, "polish" if mode == "polish" else "filter"))).read()
        got = _run(harness, d, mode, peak, *hflags)
        assert got == want, mode
        nrecords += sum(1 for l in want.splitlines() if not l.startswith("#"))
    assert nrecords > 100                       # the comparison is not vacuous


@pytest.mark.skipif(not (os.path.exists(MERFIN) and os.path.exists(MERYL)), reason="oracle/_ref not built (reference sources absent)")
@pytest.mark.parametrize("seed,k,flags", [(13, 15, ("-comb", "4", "-nosplit")), (30, 21, ())])
def test_adversarial_vcf_against_reference(harness, tmp_path, seed, k, flags):
    """One seed of tests/fuzz_variants.py per case: records crowded at record ends and on top of each
    other, overlapping REF spans, duplicate positions, multi-allelic and lower-case REF, odd genotypes."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fuzz_variants
    res = fuzz_variants.run(seed, k=k, flags=flags, nvar=60, L=4000, work=str(tmp_path))
    assert all(v.startswith("ok ") for v in res.values()), res
    assert sum(int(v.split()[1]) for v in res.values()) > 200


@pytest.mark.skipif(not (os.path.exists(MERFIN) and os.path.exists(MERYL)), reason="oracle/_ref not built (reference sources absent)")
def test_bench_vcf_generator_against_reference(harness, tmp_path):
    """The synthetic VCF of `bench.py --mode polish` (bench_modes.synth_vcf: SNVs, 1-4 bp insertions and deletions,
    1/2 genotypes, neighbours within 15 bp) through the reference and through the product's front-end: same
    .polish.vcf / .filter.vcf, in one piece and as 3 cluster ranges per record."""
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_polish as bench_modes
    from merfin_b200 import synth
    d = str(tmp_path)
    k = 21
    truth = synth.make_assembly(120000, 3, 99, n_run=15, lower_run=0, dup_len=1500)
    asm, _ = synth.plant_errors(truth, 60, 5, false_per_rec=0)
    synth.write_fasta(os.path.join(d, "asm.fa"), asm, width=80)
    names = [n for n, _ in asm]
    lens = [len(b) for _, b in asm]
    asm_np = np.concatenate([np.asarray(b, dtype=np.uint8) for _, b in asm])
    n = bench_modes.synth_vcf(os.path.join(d, "calls.vcf"), asm_np, lens, names, 1500, seed=3)
    assert n > 1000
    reads = synth.simulate_reads(truth, 12, 120, 0.005, 17)
    synth.write_fasta(os.path.join(d, "reads.fa"), reads, prefix="r")
    for fa, db in (("reads.fa", "reads.meryl"), ("asm.fa", "asm.meryl")):
        p = subprocess.run([MERYL, "threads=1", "count", f"k={k}", "memory=1", fa, "output", db], cwd=d, capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stderr[-2000:]
    total = 0
    for mode in ("polish", "filter"):
        p = subprocess.run([MERFIN, "-" + mode, "-sequence", "asm.fa", "-seqmers", "asm.meryl", "-readmers", "reads.meryl",
                            "-vcf", "calls.vcf", "-peak", "10", "-threads", "1", "-memory", "4", "-output", "ref." + mode],
                           cwd=d, capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        want = open(os.path.join(d, "ref.%s.%s.vcf" % (mode, mode))).read()
        assert _run(harness, d, mode, 10) == want, mode
        assert _run(harness, d, mode, 10, "--shards", "3", "--batch", "50") == want, mode
        total += sum(1 for l in want.splitlines() if not l.startswith("#"))
    assert total > 10                          # (random edits are mostly rejected by the selectors)


@pytest.mark.skipif(not os.path.exists(MERFIN), reason="oracle/_ref not built (reference sources absent)")
def test_many_alleles_and_crowded_clusters_against_reference(harness, golden_dir, tmp_path):
    """Genotypes that name three and four ALT alleles (a record's allele list longer than the three kept in line by
    the front-end's small vectors) and runs of five variants inside one cluster (more than the two kept in line),
    added to the golden VCF: same output as the unmodified reference run now, all five selectors."""
    import shutil
    d0 = os.path.join(golden_dir, "g1_k21")
    d = str(tmp_path)
    for f in ("asm.fa", "prob.csv"):
        shutil.copy(os.path.join(d0, f), d)
    for sub in ("reads.meryl", "asm.meryl"):
        shutil.copytree(os.path.join(d0, sub), os.path.join(d, sub))
    recs = dict(O_read_fasta(os.path.join(d0, "asm.fa")))
    lines = open(os.path.join(d0, "calls.vcf")).read().splitlines()
    head = [l for l in lines if l.startswith("#")]
    body = [l for l in lines if not l.startswith("#")]
    extra = []
    for name, at, gt in (("seq1", 1200, "1/2/3"), ("seq1", 4100, "1|2|3|4"), ("seq2", 700, "3/1/2"), ("seq2", 3300, "4/4/2/1")):
        ref = chr(recs[name][at]).upper()
        alts = [b for b in "ACGT" if b != ref] + [ref + "TT"]
        extra.append("%s\t%d\t.\t%s\t%s\t33\tPASS\t.\tGT\t%s" % (name, at + 1, ref, ",".join(alts), gt))
    for name, at in (("seq1", 5000), ("seq2", 5200)):          # five SNVs 6 bp apart: one cluster of five
        for j in range(5):
            p = at + 6 * j
            ref = chr(recs[name][p]).upper()
            extra.append("%s\t%d\t.\t%s\t%s\t40\tPASS\t.\tGT\t%s" % (name, p + 1, ref, "A" if ref != "A" else "C", ("1/1", "0/1")[j % 2]))
    # POS / QUAL spellings beyond plain integers (the loader's short cut for plain decimals must agree with strtoul / strtod)
    quals = ["30.", ".5", "1e1", "007.25", "+4", "12345678901234567", "45.123456789012345", "0", ".", "3.0e-1", "99.95", "1234567890.12345"]
    body = ["\t".join(f[:5] + [quals[i % len(quals)]] + f[6:]) if i % 2 else "\t".join(f) for i, f in enumerate(l.split("\t") for l in body)]
    body = ["\t".join([f[0], "0" * (i % 3) + f[1]] + f[2:]) for i, f in enumerate(l.split("\t") for l in body)]
    allrecs = sorted(body + extra, key=lambda l: (l.split("\t")[0], int(l.split("\t")[1])))
    open(os.path.join(d, "calls.vcf"), "w").write("\n".join(head + allrecs) + "\n")
    for mode in ("polish", "filter", "better", "strict", "loose"):
        p = subprocess.run([MERFIN, "-" + mode, "-sequence", "asm.fa", "-seqmers", "asm.meryl", "-readmers", "reads.meryl",
                            "-vcf", "calls.vcf", "-peak", "8", "-threads", "1", "-memory", "4", "-output", "ref." + mode],
                           cwd=d, capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        want = open(os.path.join(d, "ref.%s.%s.vcf" % (mode, "polish" if mode == "polish" else "filter"))).read()
        assert _run(harness, d, mode, 8) == want, mode
        assert _run(harness, d, mode, 8, "--shards", "2", "--batch", "3") == want, mode


def O_read_fasta(path):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    return oracle_lib.read_fasta(path)


def test_front_end_under_address_and_ub_sanitizers(golden_dir, tmp_path):
    """variants.cpp + fasta.cpp built with ASan + UBSan: the golden cases (k = 21 with shards / small batches / the
    threaded parse / -debug, and the k = 40 set) run clean and still give the reference's files.  The records are spans
    of the mapped file and the paths live in flat arrays: any over-read would show here."""
    exe = str(tmp_path / "variant_harness_asan")
    src = os.path.join(ROOT, "merfin_b200", "csrc", "host")
    so_dir = os.path.join(ROOT, "merfin_b200")
    obj = str(tmp_path / "mfn_oracle.o")
    subprocess.check_call(["gcc", "-O1", "-w", "-std=gnu11", "-c", os.path.join(ROOT, "oracle", "mfn_oracle.c"), "-o", obj])
    subprocess.check_call(["g++", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-std=c++17", "-ffp-contract=off", "-w",
                           "-o", exe, os.path.join(ROOT, "tests", "variant_harness.cpp"), os.path.join(src, "variants.cpp"),
                           os.path.join(src, "fasta.cpp"), obj, "-L" + so_dir, "-lmerfin_b200", "-Wl,-rpath," + so_dir, "-lm", "-pthread"])
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0", MFN_VCF_THREADS="3")       # (the harness never frees its oracle tables)
    cases = [("g1_k21", "polish", 8, ()), ("g1_k21", "filter", 8, ("--shards", "3", "--batch", "2")), ("g1_k21", "loose", 8, ()),
             ("g1_k21", "polish", 8, ("--debug",)), ("g4_k40", "polish", 6, ()), ("g4_k40", "better", 6, ("--batch", "1"))]
    for name, mode, peak, extra in cases:
        d = os.path.join(golden_dir, name)
        p = subprocess.run([exe, os.path.join(d, "asm.fa"), os.path.join(d, "reads.meryl"), os.path.join(d, "asm.meryl"),
                            os.path.join(d, "calls.vcf"), mode, str(peak), *extra], capture_output=True, text=True, env=env, cwd=str(tmp_path))
        assert p.returncode == 0, p.stderr[-3000:]
        assert "AddressSanitizer" not in p.stderr and "runtime error" not in p.stderr, p.stderr[-3000:]
        assert p.stdout == open(os.path.join(d, f"ref.{mode}.vcf")).read(), (name, mode)
