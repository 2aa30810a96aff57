"""`bench.py --mode polish`: BASELINE.json configs[3].

polish   variants/s of `-polish` over a synthetic VCF on the bench.py assembly and tables (SURVEY 8(d) C4:
         5 000 000 records at 3 Gbp, 80 % SNV / 10 % insertions / 10 % deletions of 1-4 bp, mixed genotypes,
         20 % of the sites with a neighbour within 15 bp).  Ranks split the clusters of every record
         (mfn_variants_process shard/nshards), no collective on the data path; the per-rank VCF texts
         concatenate to the single-rank output (checked through a hash of the lines).
"""
import ctypes as C
import hashlib
import json
import os
import subprocess
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

GT_CHOICES = [b"1/1", b"0/1", b"1|0", b"1/2", b"0/0", b"./."]
GT_P = [0.55, 0.20, 0.10, 0.13, 0.01, 0.01]          # 0/0 + ./. <= 2 % (each silences its whole cluster, Q10)
BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def het_sites(asm_np, lens):
    """Positions where the two haplotypes of the bench.py assembly differ (record i and i + n/2 are a pair with SNVs only):
    {record: positions}.  A VCF record there that proposes the other haplotype's base creates k-mers the read table
    knows, so the selectors have real decisions to make (random edits are all rejected)."""
    n = len(lens) // 2
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    off[1:] = np.cumsum(lens)
    out = {}
    for i in range(n):
        if lens[i] != lens[i + n]:
            continue
        a, b = asm_np[int(off[i]):int(off[i + 1])], asm_np[int(off[i + n]):int(off[i + n + 1])]
        d = np.nonzero((a != b) & (a != 78) & (b != 78))[0]
        out[i] = d
        out[i + n] = d
    return out


def synth_vcf(path, asm_np, lens, names, n_records, seed=7, lo=None, hi=None, het=None, het_frac=0.25):
    """Write a sorted synthetic VCF for the records of the assembly.  lo/hi: only positions in [lo, hi) of every record
    (the reference arm's sample).  het: {record: positions} where the partner haplotype has another base; het_frac of the
    records are SNVs proposing that base.  Returns the number of records written."""
    rng = np.random.default_rng(seed)
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    off[1:] = np.cumsum(lens)
    spans = [max(0, (L if hi is None else min(L, hi)) - (0 if lo is None else lo)) for L in lens]
    span_total = sum(x for x in spans if x >= 200)
    written = 0
    with open(path, "wb") as f:
        f.write(b"##fileformat=VCFv4.2\n")
        for n, L in zip(names, lens):
            f.write(b"##contig=<ID=%s,length=%d>\n" % (n.encode(), L))
        f.write(b"##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n")
        f.write(b"#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tsample\n")
        for r, (name, L) in enumerate(zip(names, lens)):
            a, b = (0 if lo is None else lo), (L if hi is None else min(L, hi))
            span = b - a
            if span < 200:
                continue
            want = max(1, int(round(n_records * span / span_total)))
            nprim = max(1, int(want / 1.2))
            pos = np.sort(rng.integers(a + 50, b - 50, size=nprim))
            nb = rng.random(nprim) < 0.2                                     # a neighbour 1-15 bp downstream
            pos = np.sort(np.concatenate([pos, pos[nb] + rng.integers(1, 16, size=int(nb.sum()))]))
            pos = np.unique(pos)
            hetpos = np.zeros(0, dtype=np.int64)
            if het is not None and r in het and len(het[r]):
                cand = het[r][(het[r] >= a + 50) & (het[r] < b - 50)]
                take = min(len(cand), int(want * het_frac))
                if take:
                    hetpos = np.sort(rng.choice(cand, size=take, replace=False))
                    pos = np.unique(np.concatenate([pos[:max(1, len(pos) - take)], hetpos]))
            is_het = np.isin(pos, hetpos)
            partner = None
            if len(hetpos):
                pr = (r + len(lens) // 2) % len(lens)
                partner = asm_np[int(off[pr]):int(off[pr + 1])]
            m = len(pos)
            kind = rng.random(m)
            kind[is_het] = 0.0
            ln = rng.integers(1, 5, size=m)
            gt = rng.choice(len(GT_CHOICES), size=m, p=GT_P)
            qual = rng.integers(20, 61, size=m)
            alt1 = rng.integers(1, 4, size=m)
            alt2 = rng.integers(1, 3, size=m)
            ins = rng.integers(0, 4, size=(m, 4))
            seq = asm_np[int(off[r]):int(off[r + 1])]
            nm = name.encode()
            order = np.frombuffer(b"ACTG", dtype=np.uint8)                   # code (c >> 1) & 3: A0 C1 T2 G3
            c = seq[pos]
            ci = (c >> 1) & 3
            snv = kind < 0.8
            a1c = order[(ci + alt1) % 4]
            if partner is not None:
                a1c = np.where(is_het, partner[pos], a1c)
            a2c = order[(ci + alt1 + alt2) % 4]
            a2c = np.where(a2c == c, order[(ci + alt1 + 3) % 4], a2c)
            # REF / ALT as fixed-width byte strings: SNVs vectorised, the 20 % indels one by one
            ref = c.astype(np.uint8).view("S1").astype("S6")
            alt = a1c.astype(np.uint8).view("S1").astype("S16")
            two = (gt == 3)
            alt2s = a2c.astype(np.uint8).view("S1").astype("S8")
            for i in np.nonzero(~snv)[0]:
                p = int(pos[i])
                cb = bytes([c[i]])
                if kind[i] < 0.9:                                            # insertion after the anchor base
                    alt[i] = cb + bytes(BASES[ins[i, :int(ln[i])]])
                    alt2s[i] = cb + bytes(BASES[ins[i, ::-1][:max(1, int(ln[i]) - 1)]]) + b"A"
                else:                                                        # deletion of ln bases after the anchor
                    ref[i] = bytes(seq[p:p + 1 + int(ln[i])])
                    alt[i] = cb
                    alt2s[i] = cb + b"C"
            alt = np.where(two, np.char.add(np.char.add(alt, b","), alt2s), alt)
            gts = np.array(GT_CHOICES, dtype="S3")[gt]
            line = np.char.add(nm + b"\t", (pos + 1).astype("S12"))
            for piece in (b"\t.\t", ref, b"\t", alt, b"\t", qual.astype("S3"), b"\tPASS\t.\tGT\t", gts, b"\n"):
                line = np.char.add(line, piece)
            out = line.tolist()
            f.write(b"".join(out))
            written += m
    return written


def lines_digest(chunks):
    """Order-independent over records (contigs), order-dependent inside one: sha1 of the per-contig sha1s, sorted."""
    per = {}
    for text in chunks:
        for line in text.splitlines():
            if line.startswith(b"#") or not line:
                continue
            per.setdefault(line.split(b"\t", 1)[0], hashlib.sha1()).update(line + b"\n")
    h = hashlib.sha1()
    for k in sorted(per):
        h.update(k + per[k].digest())
    return h.hexdigest(), sum(1 for _ in per)


def reference_polish(workdir, threads, peak):
    """oracle/_ref/merfin -polish on the sample; returns seconds between '-- Generate variant mers' and 'Bye!'."""
    merfin = os.path.join(ROOT, "oracle", "_ref", "merfin")
    cmd = [merfin, "-polish", "-sequence", "sample.fa", "-seqmers", "asm.meryl", "-readmers", "reads.meryl", "-vcf", "sample.vcf",
           "-peak", str(peak), "-threads", str(threads), "-memory", "64", "-output", "ref"]
    p = subprocess.Popen(cmd, cwd=workdir, stderr=subprocess.PIPE, stdout=subprocess.DEVNULL, text=True)
    t0 = t1 = None
    for line in p.stderr:
        if line.startswith("-- Generate variant mers"):
            t0 = time.time()
        elif line.startswith("Bye!"):
            t1 = time.time()
    p.wait()
    if p.returncode != 0 or t0 is None or t1 is None:
        raise RuntimeError("reference merfin -polish failed")
    return t1 - t0


def polish(args, B, ctx, api, lib, torch, dist, rank, world, local, dev, asm_np, lens, off, R, A, config, setup_s):
    """B: the bench module (shared helpers).  R, A: resident tables (fused)."""
    K, PEAK = B.K, B.PEAK
    total = int(asm_np.size)
    names = [f"rec{i + 1}" for i in range(len(lens))]
    nvar = int(args.variants if args.variants else round(5_000_000 * total / 3.0e9))
    km = api.KModel(PEAK)
    work = tempfile.mkdtemp(prefix="mfn_polish_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    vpath = os.path.join(work, "calls.vcf")
    try:
        t0 = time.time()
        if rank == 0:
            nrec = synth_vcf(vpath, asm_np, lens, names, nvar, het=het_sites(asm_np, lens))
        if world > 1:
            # every rank needs the file: same box, /dev/shm is shared; broadcast the path and wait for the writer
            obj = [vpath if rank == 0 else None, nrec if rank == 0 else None]
            dist.broadcast_object_list(obj, src=0)
            vpath, nrec = obj
        gen_s = time.time() - t0
        threads = max(1, ((os.cpu_count() or 2) - 1) // world)
        if os.environ.get("MFN_POLISH_THREADS"):
            threads = int(os.environ["MFN_POLISH_THREADS"])
        os.environ.setdefault("MFN_QUIET", "1")
        os.environ.setdefault("MFN_VCF_THREADS", str(threads))
        base = asm_np.ctypes.data
        # Scorer lanes: the device part of a scorer call (upload, pack, scan, D2H) holds its ctx's lock; extra contexts
        # on the same device (own streams and staging, the SAME resident tables: they are read-only) let the calls of
        # different host threads overlap on the device.
        nlanes = max(1, int(os.environ.get("MFN_POLISH_LANES", "2")))
        lanes = [ctx] + [api.Context(local) for _ in range(nlanes - 1)]

        def load_vcf():
            v = api.Vcf(vpath)
            v.cluster(K)
            return v

        def process(v):
            """All records, this rank's cluster range of each, on `threads` host threads; returns the texts in record order.
            A record's range is cut further into pieces of ~50 000 clusters (shard = rank * S + s of world * S): records
            are long and few, and one record per task leaves the threads waiting for the biggest one."""
            tasks = []
            for i in range(len(lens)):
                nc = v.clusters_of(names[i])
                if nc == 0:
                    continue
                S = max(1, -(-nc // (world * 50000)))
                tasks += [(nc // (world * S), i, s, S) for s in range(S)]
            tasks.sort(reverse=True)
            parts = {}
            nxt = [0]
            lock = threading.Lock()
            err = []

            def worker():
                while True:
                    with lock:
                        j = nxt[0]
                        nxt[0] += 1
                    if j >= len(tasks):
                        return
                    _, i, s_, S = tasks[j]
                    try:
                        parts[(i, s_)] = api.variants_process(lanes[j % nlanes], R, A, v, names[i], base + int(off[i]), "polish", km,
                                                              shard=rank * S + s_, nshards=world * S, seq_len=int(lens[i]))
                    except Exception as e:                       # noqa: BLE001
                        err.append(e)
                        return
            th = [threading.Thread(target=worker) for _ in range(min(threads, max(1, len(tasks))))]
            for t in th:
                t.start()
            for t in th:
                t.join()
            if err:
                raise err[0]
            out = [b""] * len(lens)
            for (i, s_) in sorted(parts):
                out[i] += parts[(i, s_)]
            return out

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        # value: VCF loaded and clustered, tables resident; a step = every record scored and selected once
        vcf = load_vcf()
        info = vcf.info
        for _ in range(max(1, min(args.warmup, 2))):
            texts = process(vcf)
        steps = max(1, args.steps)
        barrier()
        sampler = B.ClockSampler(local) if rank == 0 else None
        if sampler:
            sampler.start()
        l0, h0 = sum(c.launches for c in lanes), sum(c.h2d_bytes for c in lanes)
        for c in lanes:
            c.time_passes(True)
        w0 = time.time()
        for _ in range(steps):
            texts = process(vcf)
        barrier()
        wall = time.time() - w0
        scan_ms = d2h_ms = 0.0
        ncalls = 0
        for c in lanes:
            a_, b_, n_ = c.pass_times()
            scan_ms += a_; d2h_ms += b_; ncalls += n_
            c.time_passes(False)
        clocks = sampler.stop() if sampler else None
        launches, path_bytes = sum(c.launches for c in lanes) - l0, (sum(c.h2d_bytes for c in lanes) - h0) // steps
        t = torch.tensor([wall], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall = float(t[0])
        value = info.records * steps / wall
        vcf.close()

        # e2e: the VCF is read, parsed and clustered from the file inside the timed region as well
        e_steps = max(1, min(steps, 2))
        barrier()
        w0 = time.time()
        for _ in range(e_steps):
            v2 = load_vcf()
            texts = process(v2)
            v2.close()
        barrier()
        t = torch.tensor([time.time() - w0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_wall = float(t[0])
        e2e_value = info.records * e_steps / e_wall

        # the ranks' texts concatenate (per record) to the single-rank output: compare digests with shard 0/1 of rank 0
        mine = texts
        if world > 1:
            gathered = [None] * world
            dist.all_gather_object(gathered, mine)
            merged = [b"".join(gathered[r][i] for r in range(world)) for i in range(len(lens))]
        else:
            merged = mine
        digest, ncontig = lines_digest(merged)
        nout = sum(t.count(b"\n") for t in merged)
        path_bytes_all = path_bytes
        if world > 1:
            tt = torch.tensor([path_bytes], device=dev, dtype=torch.int64)
            dist.all_reduce(tt)
            path_bytes_all = int(tt[0])

        if rank == 0:
            peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
            peak_gbs = float(json.load(open(peaks_path))["hbm_gbs"]) if os.path.exists(peaks_path) else 6650.0
            k_ms = (scan_ms / steps) if ncalls else None
            achieved = (path_bytes * 72.25 / (k_ms / 1e3) / 1e9) if k_ms else None
            out = {"metric": "variants/sec (-polish k=21)", "value": value, "unit": "variants/s", "n_gpus": world, "steps": steps,
                   "warmup": max(1, min(args.warmup, 2)), "ms_per_step": 1e3 * wall / steps, "higher_is_better": True,
                   "scaling": "strong", "vs_baseline": None, "dtype": "u64 (probes) + f64 (k* replay on the host)", "data": "synthetic",
                   "config": dict(config, workload=config["workload"].replace("-hist", "-polish") +
                                  f" + {info.records} VCF records ({info.clusters} clusters; 80% SNV, 10% ins, 10% del of 1-4 bp; 20% with a neighbour within 15 bp)",
                                  parallelism=f"clusters of every record split over {world} GPU(s) (mfn_variants_process shard/nshards), tables replicated, "
                                              f"{threads} host threads per rank, {nlanes} scorer context(s) per GPU, no collective on the data path",
                                  l2_policy="inputs larger than L2 (every step uploads fresh path strings; the fused table is 25 GB)"),
                   "e2e": {"value": e2e_value, "unit": "variants/s", "h2d_bytes_per_step": int(path_bytes_all),
                           "d2h_bytes_per_step": int(path_bytes_all * 9), "ms_per_step": 1e3 * e_wall / e_steps, "steps": e_steps,
                           "includes": "VCF read + parse + clustering from the file, path enumeration, H2D of the paths, scoring, D2H of the pairs, selection"},
                   "gpu_launches": int(launches),
                   "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                                "frac": (achieved / peak_gbs) if achieved else None, "traffic": None,
                                "kernel": "k_scan<3> (pair per position through the fused table) over the path strings",
                                "kernel_ms": k_ms, "d2h_ms": (d2h_ms / steps) if ncalls else None, "device_calls_per_step": ncalls // max(1, steps) if ncalls else 0,
                                "path_bases_per_step": int(path_bytes), "algorithmic_bytes_per_base": 72.25,
                                "note": "the step is bound by the host side (path enumeration and selection), not by this kernel: "
                                        f"device time {k_ms if k_ms is None else round(k_ms, 1)} ms of {round(1e3 * wall / steps, 1)} ms"},
                   "clocks": clocks,
                   "result": {"records_in": int(info.records), "clusters": int(info.clusters), "records_out": int(nout),
                              "contigs_out": int(ncontig), "lines_sha1": digest, "vcf_gen_s": round(gen_s, 1), "setup_s": round(setup_s, 1)}}
            if world == 1 and not args.no_cpu_baseline:
                try:
                    out["cpu_baseline"] = polish_cpu_baseline(args, B, ctx, api, asm_np, lens, names, nvar, total)
                except Exception as e:                      # noqa: BLE001
                    out["cpu_baseline"] = {"value": None, "unit": "variants/s", "cores": os.cpu_count() or 1, "kind": "reference",
                                           "sample": "failed: %s" % (str(e)[:200],)}
            print(json.dumps(out))
    finally:
        for c in locals().get("lanes", [])[1:]:
            c.close()
        if rank == 0:
            import shutil
            shutil.rmtree(work, ignore_errors=True)
    return 0


def polish_cpu_baseline(args, B, ctx, api, asm_np, lens, names, nvar, total, repeats=1):
    """The unmodified reference `merfin -polish` on a bounded sample (first sample_mbp/24 Mbp of every record, own DBs of
    the sample, the VCF records inside it), and the same sample through mfn_variants_process: same lines."""
    import shutil
    K, PEAK = B.K, B.PEAK
    workdir = tempfile.mkdtemp(prefix="mfn_bench_")
    try:
        recs, seq, R, A, stotal = B.build_sample(ctx, asm_np, lens, args.sample_mbp * 1e6, workdir)
        per = stotal // len(recs)
        nsample = synth_vcf(os.path.join(workdir, "sample.vcf"), asm_np, lens, names, int(round(nvar * stotal / total)), seed=11, lo=0, hi=per,
                            het=het_sites(asm_np, lens))
        have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "merfin"))
        cores = os.cpu_count() or 1
        if not have_ref:
            raise RuntimeError("oracle/_ref/merfin is missing")
        threads = min(cores, len(recs))
        times = [reference_polish(workdir, threads, PEAK) for _ in range(repeats)]
        want = open(os.path.join(workdir, "ref.polish.vcf"), "rb").read()
        # ours on the same sample
        R.fuse(A)
        v = api.Vcf(os.path.join(workdir, "sample.vcf"))
        v.cluster(K)
        km = api.KModel(PEAK)
        got = [api.variants_process(ctx, R, A, v, n, b, "polish", km) for n, b in recs]
        v.close()
        parity = lines_digest([want]) == lines_digest(got)
        for o in (seq, R, A):
            o.close()
        best = min(times)
        return {"value": nsample / best, "unit": "variants/s", "cores": threads, "kind": "reference",
                "sample": f"first {per} bases of each of the {len(recs)} records ({stotal} bases), own k={K} meryl DBs, {nsample} VCF records; "
                          f"phase between '-- Generate variant mers and score them.' and 'Bye!'; host has {cores} cores",
                "polish_seconds": times, "parity_vs_gpu": bool(parity)}
    finally:
        shutil.rmtree(workdir, ignore_errors=True)
