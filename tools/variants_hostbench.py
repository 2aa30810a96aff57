"""Host-side cost of the variant modes without a GPU (tools/variants_hostbench.cpp): builds the harness, writes a
synthetic diploid assembly + the `bench.py --mode polish` VCF recipe at the same density (1 record per 600 bp) and
prints the harness's JSON line per mode.

    python tools/variants_hostbench.py [--mbp 40] [--threads 1] [--modes polish,filter] [--repeat 3]
"""
import argparse
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.bench_polish import het_sites, synth_vcf  # noqa: E402


def build():
    exe = os.path.join(ROOT, "build", "variants_hostbench")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    h = os.path.join(ROOT, "merfin_b200", "csrc", "host")
    subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-ffp-contract=off", "-pthread", "-o", exe,
                           os.path.join(ROOT, "tools", "variants_hostbench.cpp"), os.path.join(h, "variants.cpp"),
                           os.path.join(h, "fasta.cpp"), "-L" + os.path.join(ROOT, "merfin_b200"), "-lmerfin_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "merfin_b200")])      # (mfn_getK comes from the library)
    return exe


def make_inputs(d, mbp, seed=11):
    fa, vcf = os.path.join(d, "asm.fa"), os.path.join(d, "calls.vcf")
    if os.path.exists(fa) and os.path.exists(vcf):
        return fa, vcf
    rng = np.random.default_rng(seed)
    npairs = 2
    L = int(mbp * 1e6) // (2 * npairs)
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)
    haps = [bases[rng.integers(0, 4, size=L)] for _ in range(npairs)]
    recs = list(haps)
    for h in haps:                                          # the partner haplotype: 0.1 % SNVs
        b = h.copy()
        at = rng.integers(0, L, size=L // 1000)
        b[at] = bases[(np.searchsorted(bases, b[at]) + rng.integers(1, 4, size=len(at))) % 4]
        recs.append(b)
    names = ["chr%d_A" % (i + 1) for i in range(npairs)] + ["chr%d_B" % (i + 1) for i in range(npairs)]
    lens = [len(r) for r in recs]
    asm = np.concatenate(recs)
    with open(fa, "wb") as f:
        for n, r in zip(names, recs):
            f.write(b">" + n.encode() + b"\n" + r.tobytes() + b"\n")
    n = synth_vcf(vcf, asm, lens, names, int(sum(lens) / 600), het=het_sites(asm, lens))
    print("wrote %d records over %.1f Mbp" % (n, sum(lens) / 1e6), file=sys.stderr)
    return fa, vcf


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mbp", type=float, default=40.0)
    ap.add_argument("--threads", type=int, default=1)
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--modes", default="polish,filter,loose")
    ap.add_argument("--dir", default="/tmp/mfn_variants_hostbench")
    a = ap.parse_args()
    os.makedirs(a.dir, exist_ok=True)
    exe = build()
    fa, vcf = make_inputs(a.dir, a.mbp)
    for mode in a.modes.split(","):
        subprocess.check_call([exe, fa, vcf, mode, "--threads", str(a.threads), "--repeat", str(a.repeat)])


if __name__ == "__main__":
    main()
