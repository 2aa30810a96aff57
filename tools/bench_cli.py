#!/usr/bin/env python
"""BASELINE.json configs[1] as a user would run it: the merfin-compatible command line
(merfin_b200/bin/merfin) against the unmodified reference binary (oracle/_ref/merfin) on the same
FASTA file and the same on-disk meryl databases, -hist and -dump, wall clock per command, outputs
compared byte for byte.  One JSON line.

    python tools/bench_cli.py [--mbp 100] [--k 21] [--no-dump]
"""
import argparse
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B


def md5(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 24), b""):
            h.update(blk)
    return h.hexdigest()


def run(cmd, cwd, repeat=1):
    """Wall clock of the whole command (best of `repeat`: the first command on a fresh box also
    pays driver and page-cache warm-up)."""
    best = None
    for _ in range(repeat):
        t0 = time.time()
        p = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True)
        dt = time.time() - t0
        if p.returncode != 0:
            raise RuntimeError("%s failed:\n%s" % (" ".join(cmd), p.stderr[-3000:]))
        best = dt if best is None else min(best, dt)
    return best, p.stderr


def summary(err):
    import re
    return [l for l in err.splitlines() if re.match(r"^(K-mers|Missing QV|Merfin QV)", l)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mbp", type=float, default=100.0)
    ap.add_argument("--k", type=int, default=21)
    ap.add_argument("--no-dump", action="store_true")
    ap.add_argument("--variants", type=int, default=200000, help="records of the synthetic VCF for -polish / -filter (0 = skip)")
    ap.add_argument("--keep", action="store_true")
    ap.add_argument("--repeat", type=int, default=2, help="runs per timed command (best is reported)")
    args = ap.parse_args()
    import torch
    from merfin_b200 import api
    dev = torch.device("cuda", 0)
    total = int(args.mbp * 1e6)
    asm_dev, lens = B.make_assembly(total, dev)
    asm = asm_dev.cpu().numpy()
    del asm_dev
    off = B.offsets_of(lens)
    d = tempfile.mkdtemp(prefix="mfn_cli_")
    out = {"config": "BASELINE configs[1]: %.0f Mbp synthetic assembly (%d records) + 30x read meryl DB, k=%d, merfin CLI" % (total / 1e6, len(lens), args.k),
           "host_cores": os.cpu_count()}
    try:
        ctx = api.Context(0)
        import numpy as np
        from merfin_b200 import synth
        truth = [("rec%d" % (i + 1), asm[int(off[i]):int(off[i + 1])]) for i in range(len(lens))]
        if args.variants:
            # the assembly = the truth with planted SNV / indel errors; the VCF proposes the corrections
            # (mixed genotypes, clustered neighbours, some 0/0 and ./. records) plus bogus SNVs
            t0 = time.time()
            per = max(1, args.variants // len(lens))
            asm_recs, vcf = synth.plant_errors(truth, per * 9 // 10, 7, false_per_rec=per // 10)
            open(os.path.join(d, "calls.vcf"), "w").write(vcf)
            out["vcf_records"] = sum(1 for l in vcf.splitlines() if l and not l.startswith("#"))
            out["vcf_write_s"] = round(time.time() - t0, 2)
        else:
            asm_recs = truth
        recs = [(n, np.ascontiguousarray(b).tobytes()) for n, b in asm_recs]
        with open(os.path.join(d, "asm.fa"), "wb") as f:
            for n, b in recs:
                f.write(b">" + n.encode() + b"\n")
                for i in range(0, len(b), 1 << 20):
                    f.write(b[i:i + (1 << 20)] + b"\n")
        # reads come from the truth, the assembly database from the (erroneous) assembly
        tseq = api.SeqSet(ctx, [(n, np.ascontiguousarray(b).tobytes()) for n, b in truth])
        T = api.MerylExactLookup.count_sequence(ctx, args.k, tseq)
        R = api.MerylExactLookup.synth_reads(ctx, T, B.COV_PER_COPY, int(total * 0.67), 7)
        seq = api.SeqSet(ctx, recs)
        A = api.MerylExactLookup.count_sequence(ctx, args.k, seq)
        t0 = time.time()
        for t, name in ((A, "asm.meryl"), (R, "reads.meryl")):
            km, vv = t.export()
            api.meryl_write(os.path.join(d, name), args.k, km, vv)
        out["db_kmers"] = {"asm": int(A.nKmers()), "reads": int(R.nKmers())}
        out["db_write_s"] = round(time.time() - t0, 2)
        seq.close(); tseq.close(); A.close(); R.close(); T.close(); ctx.close()

        ours = os.path.join(ROOT, "merfin_b200", "bin", "merfin")
        ref = os.path.join(ROOT, "oracle", "_ref", "merfin")
        common = ["-sequence", "asm.fa", "-seqmers", "asm.meryl", "-readmers", "reads.meryl", "-peak", str(B.PEAK)]
        res = {}
        os.environ["MFN_TIMELINE"] = "1"
        t, err = run([ours, "-hist"] + common + ["-output", "ours.hist"], d, repeat=args.repeat)
        res["ours_hist_s"] = round(t, 2); ours_sum = summary(err)
        res["ours_hist_timeline"] = [l.replace("[timeline]", "").strip() for l in err.splitlines() if l.startswith("[timeline]")]
        if not args.no_dump:
            t, err = run([ours, "-dump"] + common + ["-output", "ours.dump"], d)
            res["ours_dump_s"] = round(t, 2)
        if os.path.exists(ref):
            th = str(min(os.cpu_count() or 1, len(lens)))
            t, err = run([ref, "-hist"] + common + ["-threads", th, "-memory", "64", "-output", "ref.hist"], d, repeat=args.repeat)
            res["ref_hist_s"] = round(t, 2); ref_sum = summary(err)
            res["hist_file_identical"] = open(os.path.join(d, "ours.hist"), "rb").read() == open(os.path.join(d, "ref.hist"), "rb").read()
            res["hist_summary_identical"] = ours_sum == ref_sum
            res["summary"] = ref_sum
            if not args.no_dump:
                t, err = run([ref, "-dump"] + common + ["-threads", th, "-memory", "64", "-output", "ref.dump"], d)
                res["ref_dump_s"] = round(t, 2)
                res["dump_bytes"] = os.path.getsize(os.path.join(d, "ref.dump"))
                res["dump_file_identical"] = md5(os.path.join(d, "ours.dump")) == md5(os.path.join(d, "ref.dump"))
            if args.variants:
                for mode, produced in (("-polish", "%s.polish.vcf"), ("-filter", "%s.filter.vcf")):
                    t, err = run([ours, mode] + common + ["-vcf", "calls.vcf", "-output", "ours"], d, repeat=args.repeat)
                    res["ours_%s_s" % mode[1:]] = round(t, 2)
                    res["ours_%s_timeline" % mode[1:]] = [l.replace("[timeline]", "").strip() for l in err.splitlines() if l.startswith("[timeline]")]
                    t, err = run([ref, mode] + common + ["-vcf", "calls.vcf", "-threads", th, "-memory", "64", "-output", "ref"], d, repeat=args.repeat)
                    res["ref_%s_s" % mode[1:]] = round(t, 2)
                    a, b = open(os.path.join(d, produced % "ours"), "rb").read(), open(os.path.join(d, produced % "ref"), "rb").read()
                    # the reference writes clusters in the order its worker threads finish them: compare as sets of lines too
                    res["%s_vcf_identical" % mode[1:]] = a == b
                    res["%s_vcf_same_lines" % mode[1:]] = sorted(a.splitlines()) == sorted(b.splitlines())
                    res["%s_vcf_records" % mode[1:]] = sum(1 for l in b.splitlines() if not l.startswith(b"#"))
        else:
            res["reference"] = "oracle/_ref/merfin not present"
        out.update(res)
        print(json.dumps(out))
    finally:
        if not args.keep:
            shutil.rmtree(d, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
