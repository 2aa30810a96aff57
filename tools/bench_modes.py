#!/usr/bin/env python
"""Secondary measurements next to bench.py (BASELINE.json configs[1], [3], [4]): the other report
modes of the same path and the k sweep, one JSON line each.  Not the driver's contract; the
numbers go into DESIGN.md and profiles/.

    python tools/bench_modes.py [--gbp 1.0] [--modes hist,dump,score,completeness] [--ks 15,21,31]
"""
import argparse
import ctypes as C
import json
import os
import shutil
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B                                        # workload generators of the main benchmark


def ev_ms(fn, n, warm=1):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sum(ts) / len(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gbp", type=float, default=1.0)
    ap.add_argument("--modes", default="hist,dump,score,completeness")
    ap.add_argument("--ks", default="15,21,31")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--db-fracs", default="0.25,1.0,2.0", help="dbsweep: read-database sizes as fractions of one GPU's HBM")
    args = ap.parse_args()
    import numpy as np
    import torch
    from merfin_b200 import api
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    peak_gbs = 6450.6
    pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pp):
        peak_gbs = float(json.load(open(pp))["hbm_gbs"])
    total = int(args.gbp * 1e9)
    asm_dev, lens = B.make_assembly(total, dev)
    total = int(asm_dev.numel())
    off = B.offsets_of(lens)
    asm_host = torch.empty(total, dtype=torch.uint8, pin_memory=True)
    asm_host.copy_(asm_dev)
    del asm_dev
    torch.cuda.empty_cache()
    asm_np = asm_host.numpy()
    ctx = api.Context(0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    lib = api.L.load()
    modes = args.modes.split(",")

    seq_h = C.c_void_p()
    api.L.check(lib.mfn_seq_create(ctx.h, C.c_void_p(asm_np.ctypes.data), off.ctypes.data_as(C.c_void_p), len(lens), C.byref(seq_h)))

    class _Seq:
        h = seq_h
        nseq = len(lens)
    seq = _Seq()

    def emit(d):
        d.update({"gbp": total / 1e9, "data": "synthetic", "device": torch.cuda.get_device_name(0)})
        print(json.dumps(d), flush=True)

    for k in ([int(x) for x in args.ks.split(",")] if set(modes) - {"dbsweep"} else []):
        A = api.MerylExactLookup.count_sequence(ctx, k, seq)
        R = api.MerylExactLookup.synth_reads(ctx, A, B.COV_PER_COPY, int(total * 0.67), 7)
        R.fuse(A)
        ctx.sync()
        ji = R.joint_info
        tinfo = {"k": k, "joint_GB": round(ji.device_bytes / 1e9, 2), "joint_kmers": int(ji.n_kmers), "slot_bits": int(ji.slot_bits)}
        if "hist" in modes:
            acc = api.Accum(ctx, len(lens))
            ctx.time_passes(True)
            ms = ev_ms(lambda: (acc.reset(), api.hist_resident(ctx, R, A, seq, None, acc, 0, 1)), args.steps)
            sc, pr, ns = ctx.pass_times()
            ctx.time_passes(False)
            per, pairs = acc.fetch()
            emit({"mode": "-hist resident", **tinfo, "ms": ms, "bases_per_s": total / ms * 1e3,
                  "roofline_frac": total * 64.25 / (ms / 1e3) / 1e9 / peak_gbs,
                  "pass_ms": {"scatter": sc / ns, "probe": pr / ns} if ns else None,
                  "kasm": int(per["kasm"].sum()), "kmissing": int(per["kmissing"].sum())})
            acc.close()
        if "dump" in modes and k == 21:
            # (the table restricted to the scanned sequences, as bench.py uses it: the range-partitioned -dump takes it)
            if not os.environ.get("MFN_MODES_NO_RESTRICT"):
                R.restrict(A, seq)
            n = total
            vr = torch.empty(n, dtype=torch.int32, device=dev)
            va = torch.empty(n, dtype=torch.int32, device=dev)
            ok = torch.empty(n, dtype=torch.uint8, device=dev)
            f = lambda: api.L.check(lib.mfn_dump_resident(ctx.h, R.h, A.h, seq.h, C.c_void_p(vr.data_ptr()), C.c_void_p(va.data_ptr()), C.c_void_p(ok.data_ptr())))
            ms = ev_ms(f, args.steps)
            emit({"mode": "-dump resident (pairs per position left on the device)", **tinfo, "ms": ms, "bases_per_s": total / ms * 1e3,
                  "roofline_frac": total * 72.25 / (ms / 1e3) / 1e9 / peak_gbs, "valid": int(ok.sum().item())})
            del vr, va, ok
            R.unrestrict()
        if "score" in modes and k == 21:
            # BASELINE configs[3]: variant-path scoring; paths = 2(k-1)+1 windows around an SNV, 2 paths per variant
            npaths, plen = 4_000_000, 2 * (k - 1) + 1 + k - 1
            rng = np.random.default_rng(3)
            starts = rng.integers(0, total - plen - 1, size=npaths // 2)
            paths = np.empty((npaths, plen), dtype=np.uint8)
            for i0 in range(0, npaths // 2, 1 << 18):
                st = starts[i0:i0 + (1 << 18)]
                blk = asm_np[st[:, None] + np.arange(plen)[None, :]]
                paths[2 * i0:2 * (i0 + len(st)):2] = blk
                alt = blk.copy()
                alt[:, plen // 2] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=len(st))]
                paths[2 * i0 + 1:2 * (i0 + len(st)):2] = alt
            poff = (np.arange(npaths + 1, dtype=np.uint64) * plen)
            flat = paths.reshape(-1)
            pv, pa, pk = np.zeros(flat.size, np.uint32), np.zeros(flat.size, np.uint32), np.zeros(flat.size, np.uint8)
            t0 = time.time()
            api.L.check(lib.mfn_score_paths(ctx.h, R.h, A.h, C.c_void_p(flat.ctypes.data), poff.ctypes.data_as(C.c_void_p), npaths,
                                            C.c_void_p(pv.ctypes.data), C.c_void_p(pa.ctypes.data), C.c_void_p(pk.ctypes.data)))
            dt = time.time() - t0
            windows = npaths * (plen - k + 1)
            emit({"mode": "-polish path scoring (mfn_score_paths, host buffers, wall clock)", **tinfo, "paths": npaths, "path_len": plen,
                  "windows": windows, "s": dt, "windows_per_s": windows / dt, "paths_per_s": npaths / dt, "valid_windows": int(pk.sum())})
        if "join" in modes and k == 21:
            # -completeness over the resident tables: every read k-mer against the assembly table (no database on disk)
            acc = api.Accum(ctx, 0)
            f = lambda: (acc.reset(), api.L.check(lib.mfn_completeness_resident(ctx.h, R.h, A.h, acc.h)))
            ms = ev_ms(f, args.steps)
            _, pairs = acc.fetch()
            tot, undr = C.c_double(), C.c_double()
            km_ = api.KModel(B.PEAK)
            api.L.check(lib.mfn_completeness_report(pairs.ctypes.data_as(C.c_void_p), len(pairs), C.byref(km_.c), C.byref(tot), C.byref(undr)))
            nk = int(R.nKmers())
            ri = R.info
            # bytes the join has to move: the read table's slots streamed once + one 32-byte sector of the assembly table per k-mer
            alg = ri.device_bytes + nk * 32
            emit({"mode": "-completeness join over resident tables (k_join_tables)", **tinfo, "read_kmers": nk, "asm_kmers": int(A.nKmers()),
                  "ms": ms, "kmers_per_s": nk / ms * 1e3, "algorithmic_GB": round(alg / 1e9, 2), "achieved_GBs": alg / ms / 1e6,
                  "roofline_frac": alg / ms / 1e6 / peak_gbs, "completeness": 1.0 - undr.value / tot.value if tot.value else None})
            acc.close()
        if "completeness" in modes and k == 21 and total <= 300_000_000:
            d = tempfile.mkdtemp(prefix="mfn_modes_")
            try:
                t0 = time.time()
                for t, name in ((A, "asm.meryl"), (R, "reads.meryl")):
                    km, vv = t.export()
                    api.meryl_write(os.path.join(d, name), k, km, vv)
                wsec = time.time() - t0
                # table load from disk: device decode (default) vs host decode
                loads = {}
                for label, envv in (("gpu_decode", None), ("host_decode", "1"), ("gpu_decode_again", None)):
                    if envv:
                        os.environ["MFN_HOST_DECODE"] = envv
                    else:
                        os.environ.pop("MFN_HOST_DECODE", None)
                    t0 = time.time()
                    T = api.MerylExactLookup.load(ctx, os.path.join(d, "reads.meryl"))
                    ctx.sync()
                    loads[label] = time.time() - t0
                    T.close()
                os.environ.pop("MFN_HOST_DECODE", None)
                emit({"mode": "mfn_table_load (reads.meryl from disk -> HBM table, wall clock)", **tinfo, "read_kmers": int(R.nKmers()),
                      "seconds": loads, "kmers_per_s": {k_: R.nKmers() / v for k_, v in loads.items()}})
                t0 = time.time()
                tot, undr, _pairs = api.completeness(ctx, os.path.join(d, "reads.meryl"), os.path.join(d, "asm.meryl"), api.KModel(B.PEAK))
                dt = time.time() - t0
                emit({"mode": "-completeness (on-disk meryl DBs -> stream -> probe, wall clock incl. DB decode)", **tinfo,
                      "read_kmers": int(R.nKmers()), "s": dt, "kmers_per_s": R.nKmers() / dt, "db_write_s": wsec,
                      "completeness": 1.0 - undr / tot if tot else None})
            finally:
                shutil.rmtree(d, ignore_errors=True)
        R.close(); A.close()

    if "dbsweep" in modes:
        # BASELINE configs[4]: read-database size sweep, k = 21, incl. a database of twice the HBM.  The database is the
        # synthetic one of bench.py with more error k-mers; it is STREAMED through the semi-join (mfn_tables_restrict_synth)
        # and never resident, so the table the scan probes stays the size of the assembly's k-mer set.
        k = 21
        hbm = torch.cuda.get_device_properties(0).total_memory
        A = api.MerylExactLookup.count_sequence(ctx, k, seq)
        per_kmer = 7.3                                   # bytes per k-mer of a resident read table (25.7 GB / 3.53 G k-mers at the benchmark)
        base_err = int(total * 0.67)
        ref = None
        for frac in [float(x) for x in args.db_fracs.split(",")]:
            n_kmers = int(frac * hbm / per_kmer)
            n_error = max(base_err, n_kmers - A.nKmers())
            t0 = time.time()
            S, st = api.MerylExactLookup.restrict_synth(ctx, A, seq, B.COV_PER_COPY, n_error, 7, chunk_kmers=1 << 28)
            ctx.sync()
            build_s = time.time() - t0
            acc = api.Accum(ctx, len(lens))
            ctx.time_passes(True)
            ms = ev_ms(lambda: (acc.reset(), api.hist_resident(ctx, S, A, seq, None, acc, 0, 1)), args.steps)
            sc, pr, ns = ctx.pass_times()
            ctx.time_passes(False)
            per, pairs = acc.fetch()
            row = {"mode": "-hist, read database streamed through the semi-join (never resident)", "k": k,
                   "db_fraction_of_hbm": frac, "db_kmers": int(st.db_kmers), "db_resident_table_GB_if_loaded": round(st.db_kmers * per_kmer / 1e9, 1),
                   "hbm_GB": round(hbm / 1e9, 1), "kept_error_kmers": int(st.kept_kmers), "shard_kmers": int(st.shard_kmers),
                   "restricted_table_GB": round(st.table_bytes / 1e9, 2), "stream_build_s": round(build_s, 2),
                   "stream_kmers_per_s": st.db_kmers / build_s, "ms": ms, "bases_per_s": total / ms * 1e3,
                   "roofline_frac": total * 64.25 / (ms / 1e3) / 1e9 / peak_gbs,
                   "pass_ms": {"scatter": sc / ns, "probe": pr / ns} if ns else None,
                   "kasm": int(per["kasm"].sum()), "kmissing": int(per["kmissing"].sum())}
            if frac <= 0.3:
                # small enough to load as well: the resident table of the same database gives the same integers
                R = api.MerylExactLookup.synth_reads(ctx, A, B.COV_PER_COPY, n_error, 7)
                R.fuse(A)
                acc.reset()
                api.hist_resident(ctx, R, A, seq, None, acc, 0, 1)
                per2, pairs2 = acc.fetch()
                row["equals_resident_table"] = bool(len(pairs2) == len(pairs) and (pairs2 == pairs).all() and (per2 == per).all())
                row["resident_read_table_GB"] = round(R.info.device_bytes / 1e9, 2)
                R.close()
            emit(row)
            acc.close(); S.close()
        A.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
