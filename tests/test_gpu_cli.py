"""The merfin-compatible CLI (merfin_b200/bin/merfin) on a B200 against the reference's golden
outputs: same flags, byte-identical output files, same final summary block."""
import gzip
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MERFIN = os.environ.get("MFN_TEST_CLI", os.path.join(ROOT, "merfin_b200", "bin", "merfin"))


def _run(args, cwd):
    p = subprocess.run([MERFIN] + args, cwd=cwd, capture_output=True, text=True, timeout=150)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stderr


def _summary(stderr):
    return "\n".join(l for l in stderr.splitlines() if re.match(r"^(K-mers|Missing QV|Merfin QV)", l)) + "\n"


@pytest.mark.parametrize("name,peak", [("g1_k21", 8), ("g2_k6", 3), ("g3_k31", 6)])
def test_cli_hist_dump_completeness(golden_dir, tmp_path, name, peak):
    d = os.path.join(golden_dir, name)
    base = ["-sequence", os.path.join(d, "asm.fa"), "-seqmers", os.path.join(d, "asm.meryl"),
            "-readmers", os.path.join(d, "reads.meryl"), "-peak", str(peak)]
    err = _run(base + ["-hist", "-output", "out.hist"], str(tmp_path))
    assert open(tmp_path / "out.hist", "rb").read() == open(os.path.join(d, "ref.hist"), "rb").read()
    assert _summary(err) == open(os.path.join(d, "ref.hist.summary")).read()
    per = open(os.path.join(d, "ref.hist.perseq")).read()
    assert per in err
    _run(base + ["-dump", "-output", "out.dump"], str(tmp_path))
    assert open(tmp_path / "out.dump", "rb").read() == gzip.open(os.path.join(d, "ref.dump.gz"), "rb").read()
    err = _run(base + ["-completeness"], str(tmp_path))
    assert err.split("\n\n")[-1].startswith("TOTAL") or "TOTAL readK" in err
    tail = [l for l in open(os.path.join(d, "ref.completeness")).read().splitlines() if not l.startswith("thread")]
    for l in tail:
        assert l in err
    # without -seqmers the assembly k-mers are counted on the GPU (the reference shells out to meryl)
    err = _run(base[:2] + base[4:] + ["-hist", "-output", "out2.hist"], str(tmp_path))
    assert open(tmp_path / "out2.hist", "rb").read() == open(os.path.join(d, "ref.hist"), "rb").read()
    err = _run(base[:2] + base[4:] + ["-completeness"], str(tmp_path))      # likewise for -completeness
    for l in tail:
        assert l in err


def _device_lists():
    import torch
    n = torch.cuda.device_count()
    lists = ["0,0", "0,0,0"]                      # several contexts on one GPU: the whole multi-device path of the CLI
    if n >= 2:
        lists.append(",".join(str(i) for i in range(min(n, 8))))
    return lists


@pytest.mark.parametrize("name,peak", [("g1_k21", 8), ("g2_k6", 3), ("g3_k31", 6)])
def test_cli_devices_same_output(golden_dir, tmp_path, name, peak):
    """-devices a,b,...: -hist / -completeness shard their work over several contexts (tiles resp. data files of
    the read database), the accumulators are summed device to device, and the output is byte for byte the
    single-device one (= the reference's).  Runs with one context per listed device, so '0,0' exercises the
    path on a one-GPU box; with >= 2 GPUs the real device list is used as well."""
    d = os.path.join(golden_dir, name)
    base = ["-sequence", os.path.join(d, "asm.fa"), "-seqmers", os.path.join(d, "asm.meryl"),
            "-readmers", os.path.join(d, "reads.meryl"), "-peak", str(peak)]
    tail = [l for l in open(os.path.join(d, "ref.completeness")).read().splitlines() if not l.startswith("thread")]
    for devs in _device_lists():
        err = _run(base + ["-hist", "-output", "out.hist", "-devices", devs], str(tmp_path))
        assert open(tmp_path / "out.hist", "rb").read() == open(os.path.join(d, "ref.hist"), "rb").read(), devs
        assert _summary(err) == open(os.path.join(d, "ref.hist.summary")).read(), devs
        assert open(os.path.join(d, "ref.hist.perseq")).read() in err
        err = _run(base + ["-completeness", "-devices", devs], str(tmp_path))
        for l in tail:
            assert l in err, devs
        err = _run(base[:2] + base[4:] + ["-completeness", "-devices", devs], str(tmp_path))
        for l in tail:
            assert l in err, devs


def test_cli_devices_variants(golden_dir, tmp_path, monkeypatch):
    """Variant modes over several contexts, and with every record cut into pieces of 7 clusters that are scored as
    separate tasks (the CLI does that with 50 000 clusters per piece to keep its host threads busy)."""
    d = os.path.join(golden_dir, "g1_k21")
    monkeypatch.setenv("MFN_VARIANT_PIECE", "7")
    base = ["-sequence", os.path.join(d, "asm.fa"), "-seqmers", os.path.join(d, "asm.meryl"),
            "-readmers", os.path.join(d, "reads.meryl"), "-peak", "8", "-vcf", os.path.join(d, "calls.vcf")]
    for devs in ["0"] + _device_lists():
        for mode, produced in (("polish", "polish"), ("filter", "filter"), ("loose", "filter")):
            _run(base + ["-" + mode, "-output", "o", "-devices", devs], str(tmp_path))
            assert open(tmp_path / f"o.{produced}.vcf").read() == open(os.path.join(d, f"ref.{mode}.vcf")).read(), (devs, mode)


def test_cli_min_max_prob_skipmissing(golden_dir, tmp_path):
    d = os.path.join(golden_dir, "g1_k21")
    base = ["-sequence", os.path.join(d, "asm.fa"), "-seqmers", os.path.join(d, "asm.meryl"),
            "-readmers", os.path.join(d, "reads.meryl")]
    err = _run(base + ["-hist", "-peak", "1", "-min", "3", "-output", "a.hist"], str(tmp_path))
    assert open(tmp_path / "a.hist", "rb").read() == open(os.path.join(d, "ref.min3.hist"), "rb").read()
    assert _summary(err) == open(os.path.join(d, "ref.min3.hist.summary")).read()
    err = _run(base + ["-hist", "-peak", "8", "-min", "2", "-max", "12", "-output", "b.hist"], str(tmp_path))
    assert open(tmp_path / "b.hist", "rb").read() == open(os.path.join(d, "ref.min2max12.hist"), "rb").read()
    err = _run(base + ["-hist", "-peak", "8", "-prob", os.path.join(d, "prob.csv"), "-output", "c.hist"], str(tmp_path))
    assert open(tmp_path / "c.hist", "rb").read() == open(os.path.join(d, "ref.prob.hist"), "rb").read()
    assert _summary(err) == open(os.path.join(d, "ref.prob.hist.summary")).read()
    _run(base + ["-dump", "-peak", "8", "-skipMissing", "-output", "none.dump"], str(tmp_path))
    assert not os.path.exists(tmp_path / "none.dump")          # Q9: -skipMissing writes no file at all


@pytest.mark.parametrize("mode", ["polish", "filter", "better", "strict", "loose"])
def test_cli_variant_modes(golden_dir, tmp_path, mode):
    d = os.path.join(golden_dir, "g1_k21")
    args = ["-sequence", os.path.join(d, "asm.fa"), "-seqmers", os.path.join(d, "asm.meryl"),
            "-readmers", os.path.join(d, "reads.meryl"), "-peak", "8", "-vcf", os.path.join(d, "calls.vcf"),
            "-" + mode, "-output", "out"]
    if mode == "polish":
        args.append("-debug")
    _run(args, str(tmp_path))
    produced = "out.polish.vcf" if mode == "polish" else "out.filter.vcf"      # Q7
    assert open(tmp_path / produced).read() == open(os.path.join(d, f"ref.{mode}.vcf")).read()
    if mode == "polish":
        assert gzip.open(tmp_path / "out.00.debug.gz", "rt").read() == gzip.open(os.path.join(d, "ref.polish.debug.gz"), "rt").read()


def test_cli_variant_shards_concatenate(golden_dir, tmp_path):
    """-shard i/n scores cluster range i of n of every record (the multi-GPU sharding of -polish / -filter):
    the shard files, concatenated per record, hold exactly the lines of the unsharded output."""
    d = os.path.join(golden_dir, "g1_k21")
    base = ["-sequence", os.path.join(d, "asm.fa"), "-seqmers", os.path.join(d, "asm.meryl"),
            "-readmers", os.path.join(d, "reads.meryl"), "-peak", "8", "-vcf", os.path.join(d, "calls.vcf")]
    for mode, produced in (("polish", "polish"), ("filter", "filter")):
        want = open(os.path.join(d, f"ref.{mode}.vcf")).read().splitlines()
        got = []
        for i in range(3):
            _run(base + ["-" + mode, "-output", f"s{i}", "-shard", f"{i}/3"], str(tmp_path))
            got += open(tmp_path / f"s{i}.{produced}.vcf").read().splitlines()
        hdr = [l for l in want if l.startswith("#")]
        assert [l for l in got if l.startswith("#")] == hdr                  # the header comes from shard 0 only
        # records of a contig are written shard by shard: same multiset, and the same order inside every contig
        body_w, body_g = [l for l in want if not l.startswith("#")], [l for l in got if not l.startswith("#")]
        assert sorted(body_w) == sorted(body_g)
        for c in {l.split("\t")[0] for l in body_w}:
            assert [l for l in body_w if l.split("\t")[0] == c] == [l for l in body_g if l.split("\t")[0] == c]


def test_variants_abi_shards_and_threads(golden_dir):
    """mfn_variants_process through the ctypes stub: every record in 1 piece and in 4 shards, called from
    several host threads at once (the device part of mfn_score_paths is serialised inside the library)."""
    import threading
    import oracle_lib as O
    from merfin_b200 import api
    d = os.path.join(golden_dir, "g1_k21")
    recs = O.read_fasta(os.path.join(d, "asm.fa"))
    ctx = api.Context(0)
    R = api.MerylExactLookup.load(ctx, os.path.join(d, "reads.meryl"))
    A = api.MerylExactLookup.load(ctx, os.path.join(d, "asm.meryl"))
    R.fuse(A)
    km = api.KModel(8.0)
    vcf = api.Vcf(os.path.join(d, "calls.vcf"))
    vcf.cluster(21)
    assert vcf.info.records > 0 and vcf.info.clusters > 0
    want = open(os.path.join(d, "ref.polish.vcf"), "rb").read()
    whole = vcf.header() + b"".join(api.variants_process(ctx, R, A, vcf, n, b, "polish", km) for n, b in recs)
    assert whole == want
    out = {}

    def work(i, n, b):
        out[i] = b"".join(api.variants_process(ctx, R, A, vcf, n, b, "polish", km, shard=s, nshards=4) for s in range(4))
    th = [threading.Thread(target=work, args=(i, n, b)) for i, (n, b) in enumerate(recs)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert vcf.header() + b"".join(out[i] for i in range(len(recs))) == want
    vcf.close(); R.close(); A.close(); ctx.close()


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "merfin")), reason="oracle/_ref not built (reference sources absent)")
def test_cli_100mbp_against_reference():
    """BASELINE configs[1] at full size: 100 Mbp FASTA + on-disk meryl databases through the CLI and through the
    unmodified reference binary (tools/bench_cli.py): -hist file and summary block, the 2.8 GB -dump file,
    -polish / -filter VCFs identical."""
    import json
    import sys
    mbp = os.environ.get("MFN_TEST_CLI_MBP", "100")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_cli.py"), "--mbp", mbp, "--variants", "40000", "--repeat", "1"],
                       capture_output=True, text=True, timeout=1500)
    assert p.returncode == 0, p.stderr[-3000:]
    r = json.loads(p.stdout.strip().splitlines()[-1])
    if os.environ.get("MFN_TEST_OUT"):                       # keep the timings of the run (profiles/)
        open(os.environ["MFN_TEST_OUT"], "w").write(json.dumps(r) + "\n")
    assert r["hist_file_identical"] and r["hist_summary_identical"], r
    assert r["dump_file_identical"] and r["dump_bytes"] > 20 * float(mbp) * 1e6, r
    assert r["polish_vcf_same_lines"] and r["filter_vcf_same_lines"] and r["polish_vcf_records"] > 1000, r
    print(json.dumps(r))


def test_cli_compressed_outputs(golden_dir, tmp_path):
    """-output names ending in .gz / .bz2 / .xz are written through the compressor, as the reference's
    compressedFileWriter does; a path with a single quote in it is quoted for the shell."""
    import bz2
    import lzma
    d = os.path.join(golden_dir, "g1_k21")
    base = ["-sequence", os.path.join(d, "asm.fa"), "-seqmers", os.path.join(d, "asm.meryl"),
            "-readmers", os.path.join(d, "reads.meryl"), "-peak", "8"]
    want = open(os.path.join(d, "ref.hist"), "rb").read()
    for name, opener in (("o.hist.gz", gzip.open), ("o.hist.bz2", bz2.open), ("o.hist.xz", lzma.open), ("it's.hist.gz", gzip.open)):
        _run(base + ["-hist", "-output", name], str(tmp_path))
        assert opener(tmp_path / name, "rb").read() == want, name
    _run(base + ["-dump", "-output", "o.dump.gz"], str(tmp_path))
    assert gzip.open(tmp_path / "o.dump.gz", "rb").read() == gzip.open(os.path.join(d, "ref.dump.gz"), "rb").read()


def test_cli_usage_errors(tmp_path):
    p = subprocess.run([MERFIN, "-hist"], capture_output=True, text=True)
    assert p.returncode == 1
    assert "No input sequences (-sequence) supplied." in p.stderr and "No read meryl database (-readmers) supplied." in p.stderr


def test_cli_reads_databases_made_by_meryl_set_operations(tmp_path, monkeypatch):
    """Read databases in the wild are `meryl union-sum` of per-run counts, often filtered with
    `greater-than`: written by meryl's streaming writer, with sparse prefixes.  Built here with the
    unmodified reference tools (oracle/_ref, when it travelled to the box); -hist through the CLI with
    the device decoder and with the host decoder must equal the reference's file and summary."""
    ref_merfin = os.path.join(ROOT, "oracle", "_ref", "merfin")
    ref_meryl = os.path.join(ROOT, "oracle", "_ref", "meryl")
    if not (os.path.exists(ref_merfin) and os.path.exists(ref_meryl)):
        pytest.skip("oracle/_ref not built (reference sources absent)")
    from merfin_b200 import synth
    d = str(tmp_path)
    k = 21

    def ref(cmd):
        p = subprocess.run(cmd, cwd=d, capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-2000:]
        return p.stderr

    truth = synth.make_assembly(1_500_000, 4, 11, n_run=50, lower_run=100, dup_len=20000)
    synth.write_fasta(os.path.join(d, "asm.fa"), truth, width=80)
    for i in range(3):
        synth.write_fasta(os.path.join(d, "r%d.fa" % i), synth.simulate_reads(truth, 5, 150, 0.01, 100 + i), prefix="r")
        ref([ref_meryl, "threads=4", "count", f"k={k}", "memory=2", "r%d.fa" % i, "output", "r%d.meryl" % i])
    ref([ref_meryl, "threads=4", "count", f"k={k}", "memory=2", "asm.fa", "output", "asm.meryl"])
    ref([ref_meryl, "threads=4", "union-sum", "r0.meryl", "r1.meryl", "r2.meryl", "output", "union.meryl"])
    ref([ref_meryl, "threads=4", "greater-than", "1", "union.meryl", "output", "solid.meryl"])
    for db in ("union.meryl", "solid.meryl"):
        base = ["-hist", "-sequence", "asm.fa", "-seqmers", "asm.meryl", "-readmers", db, "-peak", "14"]
        err = ref([ref_merfin] + base + ["-threads", "4", "-memory", "8", "-output", "ref.hist"])
        want, want_summary = open(os.path.join(d, "ref.hist"), "rb").read(), _summary(err)
        assert len(want) > 20 and "Merfin QV*" in want_summary
        for host_decode in ("0", "1"):
            monkeypatch.setenv("MFN_HOST_DECODE", host_decode)
            err = _run(base + ["-output", "ours.hist"], d)
            assert open(os.path.join(d, "ours.hist"), "rb").read() == want, (db, host_decode)
            assert _summary(err) == want_summary, (db, host_decode)
