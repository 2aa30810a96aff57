#!/usr/bin/env python
"""Regenerate tests/golden/ by running the UNMODIFIED reference binaries.

Needs oracle/_ref/{merfin,meryl} (make -C oracle ref; only possible where
/root/reference exists).  Everything written here is committed so that the
tests run without the reference tree.

    python tests/golden/make_golden.py
"""
import gzip
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from merfin_b200 import synth  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
MERFIN, MERYL = os.path.join(REF, "merfin"), os.path.join(REF, "meryl")


def run(cmd, cwd, stderr_to=None):
    p = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if p.returncode != 0:
        sys.stderr.write(p.stderr)
        raise SystemExit(f"FAILED: {' '.join(cmd)}")
    if stderr_to:
        with open(os.path.join(cwd, stderr_to), "w") as f:
            f.write(p.stderr)
    return p


def summary_block(stderr_text):
    """The final summary merfin -hist prints (merfin-histogram.C:167-172)."""
    keep = [l for l in stderr_text.splitlines()
            if re.match(r"^(K-mers|Missing QV|Merfin QV)", l)]
    return "\n".join(keep) + "\n"


def per_seq_lines(stderr_text, names):
    return "\n".join(l for l in stderr_text.splitlines() if l.split("\t")[0] in names and "\t" in l) + "\n"


def meryl_count(d, k, fasta, out):
    run([MERYL, "threads=1", "count", f"k={k}", "memory=1", fasta, "output", out], d)


def meryl_print(d, db, out):
    p = run([MERYL, "threads=1", "print", db], d)
    with gzip.GzipFile(os.path.join(d, out), "wb", mtime=0) as f:
        f.write(p.stdout.encode())


def make_set(name, k, total, nrec, seed, cov, peak, dup_len, extra=()):
    d = os.path.join(HERE, name)
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    truth = synth.make_assembly(total, nrec, seed, n_run=50 if total >= 4000 else 7,
                                lower_run=100 if total >= 4000 else 11, dup_len=dup_len)
    recs = truth
    if "variants" in extra:          # the assembly carries planted errors; the VCF proposes corrections
        recs, vcf = synth.plant_errors(truth, 14, seed + 7, false_per_rec=5)
        open(os.path.join(d, "calls.vcf"), "w").write(vcf)
    synth.write_fasta(os.path.join(d, "asm.fa"), recs, width=70)
    # a stretch of record 0 is never sequenced -> missing k-mers
    skip = [(0, len(truth[0][1]) // 3, len(truth[0][1]) // 3 + 3 * k)]
    reads = synth.simulate_reads(truth, cov, 150 if total >= 4000 else 40, 0.005, seed + 1, skip=skip)
    synth.write_fasta(os.path.join(d, "reads.fa"), reads, prefix="r")
    meryl_count(d, k, "reads.fa", "reads.meryl")
    meryl_count(d, k, "asm.fa", "asm.meryl")
    os.remove(os.path.join(d, "reads.fa"))
    meryl_print(d, "reads.meryl", "reads.print.gz")
    meryl_print(d, "asm.meryl", "asm.print.gz")
    names = [n for n, _ in recs]
    base = [MERFIN, "-sequence", "asm.fa", "-seqmers", "asm.meryl", "-readmers", "reads.meryl",
            "-threads", "1", "-memory", "4"]
    p = run(base + ["-hist", "-peak", str(peak), "-output", "ref.hist"], d)
    open(os.path.join(d, "ref.hist.summary"), "w").write(summary_block(p.stderr))
    open(os.path.join(d, "ref.hist.perseq"), "w").write(per_seq_lines(p.stderr, names))
    p = run(base + ["-dump", "-peak", str(peak), "-output", "ref.dump"], d)
    open(os.path.join(d, "ref.dump.perseq"), "w").write(per_seq_lines(p.stderr, names))
    with open(os.path.join(d, "ref.dump"), "rb") as f, gzip.GzipFile(os.path.join(d, "ref.dump.gz"), "wb", mtime=0) as g:
        g.write(f.read())
    os.remove(os.path.join(d, "ref.dump"))
    p = run(base + ["-completeness", "-peak", str(peak)], d)
    open(os.path.join(d, "ref.completeness"), "w").write(
        "\n".join(l for l in p.stderr.splitlines() if re.match(r"^(TOTAL|COMPLETENESS|thread)", l)) + "\n")
    if "min" in extra:      # Q1: -min offsets every reported read multiplicity
        p = run(base + ["-hist", "-peak", "1", "-min", "3", "-output", "ref.min3.hist"], d)
        open(os.path.join(d, "ref.min3.hist.summary"), "w").write(summary_block(p.stderr))
        p = run(base + ["-hist", "-peak", str(peak), "-min", "2", "-max", "12", "-output", "ref.min2max12.hist"], d)
        open(os.path.join(d, "ref.min2max12.hist.summary"), "w").write(summary_block(p.stderr))
    if "prob" in extra:     # -prob table overrides readK / prob for small multiplicities
        with open(os.path.join(d, "prob.csv"), "w") as f:
            f.write("0,0.95\n0,0.80\n1,0.60\n1,0.75\nbad line\n1,0.90\n1,0.97\n2,0.55\n")
        p = run(base + ["-hist", "-peak", str(peak), "-prob", "prob.csv", "-output", "ref.prob.hist"], d)
        open(os.path.join(d, "ref.prob.hist.summary"), "w").write(summary_block(p.stderr))
        open(os.path.join(d, "ref.prob.hist.perseq"), "w").write(per_seq_lines(p.stderr, names))
        run(base + ["-dump", "-peak", str(peak), "-prob", "prob.csv", "-output", "ref.prob.dump"], d)
        with open(os.path.join(d, "ref.prob.dump"), "rb") as f, gzip.GzipFile(os.path.join(d, "ref.prob.dump.gz"), "wb", mtime=0) as g:
            g.write(f.read())
        os.remove(os.path.join(d, "ref.prob.dump"))
        p = run(base + ["-completeness", "-peak", str(peak), "-prob", "prob.csv"], d)
        open(os.path.join(d, "ref.prob.completeness"), "w").write(
            "\n".join(l for l in p.stderr.splitlines() if re.match(r"^(TOTAL|COMPLETENESS|thread)", l)) + "\n")
    if "variants" in extra:
        vbase = base + ["-vcf", "calls.vcf", "-peak", str(peak)]
        for mode in ("polish", "filter", "better", "strict", "loose"):
            run(vbase + ["-" + mode, "-output", "ref." + mode] + (["-debug"] if mode == "polish" else []), d)
            produced = "ref.polish.polish.vcf" if mode == "polish" else f"ref.{mode}.filter.vcf"     # Q7
            os.rename(os.path.join(d, produced), os.path.join(d, f"ref.{mode}.vcf"))
        dbg = os.path.join(d, "ref.polish.00.debug.gz")
        with gzip.open(dbg, "rb") as f, gzip.GzipFile(os.path.join(d, "ref.polish.debug.gz"), "wb", mtime=0) as g:
            g.write(f.read())
        os.remove(dbg)
        # with a -prob table as well (prob weights enter the tie-break sums)
        run(vbase + ["-polish", "-prob", "prob.csv", "-output", "ref.probpolish"], d)
        os.rename(os.path.join(d, "ref.probpolish.polish.vcf"), os.path.join(d, "ref.probpolish.vcf"))
    # per-file .merylIndex files are never read on the sequential path; drop them to keep fixtures small
    for sub in ("reads.meryl", "asm.meryl"):
        for fn in os.listdir(os.path.join(d, sub)):
            if fn.endswith(".merylIndex"):
                os.remove(os.path.join(d, sub, fn))


if __name__ == "__main__":
    if not (os.path.exists(MERFIN) and os.path.exists(MERYL)):
        raise SystemExit("build the reference first: make -C oracle ref")
    # G1: k=21, 2 records, N run + lowercase run + duplicated segment, 10x reads
    make_set("g1_k21", 21, 12000, 2, 11, 10, 8, 300, extra=("min", "prob", "variants"))
    # G2: even k with palindromes (Q2), tiny
    make_set("g2_k6", 6, 1500, 2, 12, 6, 3, 40)
    # G3: k=31 (62-bit k-mers, 64-bit table slots)
    make_set("g3_k31", 31, 6000, 3, 13, 8, 6, 200)
    # G4: k=40 (80-bit k-mers: the 128-bit path), with a -prob table and the variant modes
    make_set("g4_k40", 40, 6000, 3, 14, 8, 6, 200, extra=("prob", "variants"))
    # G5: k=64, the widest k-mer the reference's 128-bit kmdata holds (even k: palindromes)
    make_set("g5_k64", 64, 4000, 2, 15, 8, 6, 150)
    print("golden sets written under", HERE)
