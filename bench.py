#!/usr/bin/env python
"""bench.py -- bases/sec of the -hist hot path (k=21) on a synthetic diploid assembly.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--gbp 3.0] [--impl ours|reference]

One "step" is one full -hist pass over the whole assembly: every window is canonicalised and
probed in the read and assembly tables and accumulated into the joint histogram.
  value  whole-job bases/s with the 2-bit packed assembly and both tables resident in HBM
         (CUDA events, max over ranks); N ranks shard the tiles, one NCCL all-reduce per step.
  e2e    the same metric through the host-buffer C-ABI call (mfn_hist_host): the assembly
         starts as ASCII in pinned host memory, H2D + pack + scan + all-reduce + D2H of the
         result every step.
  roofline / cpu_baseline: see DESIGN.md "Measurement".
`--impl reference` times the unmodified reference binary (oracle/_ref/merfin -hist) on the host
cores over a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K = 21
PEAK = 15.0
COV_PER_COPY = 15.0
BYTES_PER_BASE = 64.25          # 2 tables x one 32 B sector + 0.25 B of 2-bit sequence
CHROM_FRACTIONS = [250, 240, 200, 190, 180, 170, 160, 145, 140, 135, 130, 60]   # 12 haplotype pairs


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--gbp", type=float, default=float(os.environ.get("MFN_BENCH_GBP", "3.0")))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sample-mbp", type=float, default=24.0, help="size of the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="hist", choices=["hist", "polish"],
                    help="hist: the headline (BASELINE configs[2]); polish: variants/s of -polish over a synthetic VCF (configs[3], tools/bench_polish.py)")
    ap.add_argument("--variants", type=int, default=0, help="--mode polish: VCF records (default 5 000 000 per 3 Gbp)")
    ap.add_argument("--emulate-shards", type=int, default=0,
                    help="development aid, 1 GPU: time shard 0 of N exactly as rank 0 of an N-GPU run would (no collective); "
                         "the line is marked emulated and is not a bench value")
    ap.add_argument("--no-direct-check", action="store_true", help="skip the untimed direct-scan pass that re-derives the result")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------------------------
def record_lengths(total):
    half = total // 2
    s = sum(CHROM_FRACTIONS)
    lens = [max(1000, half * f // s) for f in CHROM_FRACTIONS]
    return lens + lens                         # hap A records then hap B records


def make_assembly(total, device, seed=1234):
    """Synthetic diploid assembly as ASCII bytes on `device`: hap A uniform random with a few
    duplicated segments and N gaps; hap B = hap A with 0.1 % SNVs.  Identical on every rank."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    lens = record_lengths(total)
    half = sum(lens[:12])
    def to_ascii(c):                              # 0,1,2,3 -> A,C,G,T without an int64 index tensor
        return torch.where(c == 3, torch.full_like(c, 84), 65 + 2 * c + 2 * (c == 2).to(torch.uint8))
    code = torch.randint(0, 4, (half,), dtype=torch.uint8, device=device, generator=g)
    # duplications inside hap A (assembly multiplicity 2 per haplotype)
    ndup = max(1, half // 50_000_000)
    for i in range(ndup):
        L = min(100_000, half // 100)
        src = int(torch.randint(0, half - L, (1,), generator=g, device=device))
        dst = int(torch.randint(0, half - L, (1,), generator=g, device=device))
        code[dst:dst + L] = code[src:src + L].clone()
    hapA = to_ascii(code)
    # hap B: SNVs at 0.1 %
    CH = 1 << 28
    hapB_code = code.clone()
    for o in range(0, half, CH):
        e = min(half, o + CH)
        r = torch.rand(e - o, device=device, generator=g)
        m = r < 0.001
        bump = (torch.randint(1, 4, (e - o,), dtype=torch.uint8, device=device, generator=g))
        hapB_code[o:e] = torch.where(m, (hapB_code[o:e] + bump) % 4, hapB_code[o:e])
        del r, m, bump
    hapB = to_ascii(hapB_code)
    del code, hapB_code
    asm = torch.cat([hapA, hapB])
    del hapA, hapB
    # N gaps: ~1 % of bases in runs of 1000 (at least one run)
    nruns = max(1, int(asm.numel() * 0.01 / 1000))
    starts = torch.randint(0, max(1, asm.numel() - 1000), (nruns,), generator=g, device=device)
    idx = (starts[:, None] + torch.arange(1000, device=device)[None, :]).reshape(-1)
    asm[idx] = ord("N")
    return asm, lens


def offsets_of(lens):
    import numpy as np
    off = np.zeros(len(lens) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens, dtype=np.uint64)
    return off


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed region: NVML from a thread every 2 ms (a timed region of
    a few milliseconds at N = 8 is over before nvidia-smi -lms prints its first line), nvidia-smi as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index, self.rows, self.p = index, [], None
        self.sm, self.mx, self.bits, self.stop_flag, self.thread, self.nvml = [], None, 0, False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber: go through the UUID of the torch device
            import torch
            uuid = str(torch.cuda.get_device_properties(index).uuid)
            h = None
            for i in range(pynvml.nvmlDeviceGetCount()):
                hi = pynvml.nvmlDeviceGetHandleByIndex(i)
                u = pynvml.nvmlDeviceGetUUID(hi)
                u = u.decode() if isinstance(u, bytes) else u
                if uuid in u:
                    h = hi
                    break
            self.nvml, self.h = pynvml, h if h is not None else pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:                                       # noqa: BLE001
            self.nvml = None

    def _poll(self):
        n = self.nvml
        while not self.stop_flag:
            try:
                self.sm.append(int(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                self.bits |= int(n.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:                                   # noqa: BLE001
                pass
            time.sleep(0.002)

    def start(self):
        if self.nvml:
            try:
                self.mx = int(self.nvml.nvmlDeviceGetMaxClockInfo(self.h, self.nvml.NVML_CLOCK_SM))
            except Exception:                                   # noqa: BLE001
                self.mx = None
            self.stop_flag = False
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.thread:
            self.stop_flag = True
            self.thread.join()
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx,
                    "reasons": sorted(v for k, v in self.REASONS.items() if self.bits & k), "samples": len(sm), "source": "nvml, 2 ms"}
        if self.p:
            self.p.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": "nvidia-smi -lms 100"}


def device_tensor_view(ptr, nwords):
    """torch int64 tensor aliasing a device buffer owned by the library (zero copy)."""
    import torch

    class _W:
        __cuda_array_interface__ = {"shape": (nwords,), "typestr": "<i8", "data": (ptr, False), "version": 2}
    return torch.as_tensor(_W(), device="cuda")


# ----------------------------------------------------------------------------------------------
# CPU baseline: the unmodified reference on a bounded sample
# ----------------------------------------------------------------------------------------------
def build_sample(ctx, asm_host, lens, sample_bases, workdir):
    """First sample_bases/len(lens) bases of every record -> FASTA + meryl DBs written with the
    library's own writer from tables built exactly like the full-size ones."""
    import numpy as np
    from merfin_b200 import api
    off = offsets_of(lens)
    per = max(K + 50, int(sample_bases) // len(lens))
    recs = []
    for i, L in enumerate(lens):
        a = int(off[i])
        recs.append((f"rec{i + 1}", asm_host[a:a + min(per, L)].tobytes()))
    with open(os.path.join(workdir, "sample.fa"), "wb") as f:
        for n, b in recs:
            f.write(b">" + n.encode() + b"\n" + b + b"\n")
    seq = api.SeqSet(ctx, recs)
    A = api.MerylExactLookup.count_sequence(ctx, K, seq)
    total = sum(len(b) for _, b in recs)
    R = api.MerylExactLookup.synth_reads(ctx, A, COV_PER_COPY, int(total * 0.67), 99)
    for t, name in ((A, "asm.meryl"), (R, "reads.meryl")):
        km, vv = t.export()
        api.meryl_write(os.path.join(workdir, name), K, km, vv)
    return recs, seq, R, A, total


def run_reference_hist(workdir, threads):
    """oracle/_ref/merfin -hist on the sample; returns (hist-phase seconds, load seconds)."""
    merfin = os.path.join(ROOT, "oracle", "_ref", "merfin")
    cmd = [merfin, "-hist", "-sequence", "sample.fa", "-seqmers", "asm.meryl", "-readmers", "reads.meryl",
           "-peak", str(PEAK), "-threads", str(threads), "-memory", "64", "-output", "ref.hist"]
    t0 = time.time()
    p = subprocess.Popen(cmd, cwd=workdir, stderr=subprocess.PIPE, stdout=subprocess.DEVNULL, text=True)
    t_gen = t_end = None
    for line in p.stderr:
        if line.startswith("-- Generate histogram"):
            t_gen = time.time()
        elif line.startswith("K-mers not found in reads"):
            t_end = time.time()
    p.wait()
    if p.returncode != 0 or t_gen is None or t_end is None:
        raise RuntimeError("reference merfin failed")
    return t_end - t_gen, t_gen - t0


def run_port_hist(recs, workdir):
    """Fallback when the reference binary is absent: the C oracle port, single thread."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as O
    oR, oA = O.Lookup(O.DB(os.path.join(workdir, "reads.meryl"))), O.Lookup(O.DB(os.path.join(workdir, "asm.meryl")))
    km = O.KModel(PEAK)
    h = O.Hist()
    t0 = time.time()
    for _, b in recs:
        h.add_record(oR, oA, km, b)
    dt = time.time() - t0
    open(os.path.join(workdir, "ref.hist"), "wb").write(h.report())
    return dt


def cpu_baseline(ctx, asm_host, lens, sample_mbp, repeats=1):
    from merfin_b200 import api
    workdir = tempfile.mkdtemp(prefix="mfn_bench_")
    try:
        recs, seq, R, A, total = build_sample(ctx, asm_host, lens, sample_mbp * 1e6, workdir)
        have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "merfin"))
        cores = os.cpu_count() or 1
        times = []
        if have_ref:
            threads = min(cores, len(recs))
            try:
                for _ in range(repeats):
                    t, load = run_reference_hist(workdir, threads)
                    times.append(t)
                kind, used = "reference", threads
            except (RuntimeError, OSError):
                have_ref, times = False, []             # e.g. a binary built for another libc: time the port instead
        if not have_ref:
            for _ in range(repeats):
                times.append(run_port_hist(recs, workdir))
            kind, used = "port", 1
        # parity on the same sample, same tables: GPU -hist file must equal the CPU one byte for byte
        res = api.hist(ctx, R, A, recs, api.KModel(PEAK))
        parity = res.file_text() == open(os.path.join(workdir, "ref.hist"), "rb").read()
        for o in (seq, R, A):
            o.close()
        best = min(times)
        return {"value": total / best, "unit": "bases/s", "cores": used, "kind": kind,
                "sample": f"first {total // len(recs)} bases of each of the {len(recs)} records ({total} bases), "
                          f"own k={K} read/assembly meryl DBs; hist phase only; host has {cores} cores",
                "hist_seconds": times, "parity_vs_gpu": bool(parity)}, total
    finally:
        shutil.rmtree(workdir, ignore_errors=True)


# ----------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return 0
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and args.impl == "ours":
        dist.init_process_group("nccl", device_id=dev)

    from merfin_b200 import api
    total = int(args.gbp * 1e9)
    asm_dev, lens = make_assembly(total, dev)
    total = int(asm_dev.numel())
    off = offsets_of(lens)
    asm_host = torch.empty(total, dtype=torch.uint8, pin_memory=True)
    asm_host.copy_(asm_dev)
    del asm_dev
    torch.cuda.empty_cache()
    asm_np = asm_host.numpy()

    # the host packing lane of mfn_hist_host uses the host cores: split them between the ranks of this box
    if "MFN_HOST_PACK_THREADS" not in os.environ and world > 1:
        os.environ["MFN_HOST_PACK_THREADS"] = str(max(1, ((os.cpu_count() or 2) - 2) // world))
    srank, sworld = rank, world                  # shard of the assembly this process scans
    if args.emulate_shards > 1 and world == 1:
        srank, sworld = 0, args.emulate_shards
    ctx = api.Context(local)
    # a non-default torch stream: our kernels, torch's events and NCCL are all ordered on it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    lib = api.L.load()
    config = {"workload": f"{total / 1e9:.2f} Gbp synthetic diploid assembly (24 records, 0.1% SNVs, 1% N) + "
                          f"synthetic {2 * COV_PER_COPY:.0f}x read table, k={K}, -hist",
              "k": K, "records": len(lens), "parallelism": f"tiles sharded over {world} GPU(s); read/assembly/fused tables replicated, the streamed table restricted to each rank's shard",
              "l2_policy": "inputs larger than L2 (24 GB of window entries and a 25 GB table stream through the 126 MB L2 every step)"}
    if sworld != world:
        config["emulated"] = f"shard 0 of {sworld} on one GPU, no collective: development figure, not a bench value"

    if args.impl == "reference" and args.mode == "polish":
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_polish as bench_modes
        names = [f"rec{i + 1}" for i in range(len(lens))]
        nvar = int(args.variants if args.variants else round(5_000_000 * total / 3.0e9))
        cb = bench_modes.polish_cpu_baseline(args, sys.modules[__name__], ctx, api, asm_np, lens, names, nvar, total,
                                             repeats=args.warmup + args.steps)
        times = cb["polish_seconds"][args.warmup:]
        nsample = cb["value"] * min(cb["polish_seconds"])
        v = nsample * len(times) / sum(times)
        cb["value"] = v
        config = dict(config, full_workload=config["workload"].replace("-hist", "-polish"),
                      workload=f"SAMPLE of the -polish workload: {cb['sample']}; each step = one reference run over the sample",
                      parallelism=f"{cb['cores']} host threads (reference -threads), no GPU", l2_policy="n/a (CPU)")
        print(json.dumps({"impl": "reference", "metric": "variants/sec (-polish k=21)", "value": v, "unit": "variants/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "cpu_baseline": cb,
                          "e2e": {"value": v, "unit": "variants/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0
    if args.impl == "reference":
        cb, stotal = cpu_baseline(ctx, asm_np, lens, args.sample_mbp, repeats=args.warmup + args.steps)
        times = cb["hist_seconds"][args.warmup:]
        v = stotal * len(times) / sum(times)
        cb["value"] = v
        # say what this arm ran: a bounded sample of the workload with its own (smaller) databases
        config = dict(config)
        config["full_workload"] = config["workload"]
        config["workload"] = (f"SAMPLE of the {total / 1e9:.2f} Gbp workload: {stotal / 1e6:.1f} Mbp (first {stotal // len(lens)} bases of each of the "
                              f"{len(lens)} records) + own k={K} read/assembly meryl DBs of the sample, -hist, reference merfin on {cb['cores']} host threads; "
                              f"each step = one reference run over the sample, hist phase only")
        config["parallelism"] = f"{cb['cores']} host threads (reference -threads), no GPU"
        config["l2_policy"] = "n/a (CPU)"
        print(json.dumps({"impl": "reference", "metric": "bases/sec (-hist k=21)", "value": v, "unit": "bases/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config,
                          "cpu_baseline": cb,
                          "e2e": {"value": v, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    # ---- tables (replicated on every rank) -------------------------------------------------
    t0 = time.time()
    bases_ptr = C.c_void_p(asm_np.ctypes.data)
    seq_h = C.c_void_p()
    api.L.check(lib.mfn_seq_create(ctx.h, bases_ptr, off.ctypes.data_as(C.c_void_p), len(lens), C.byref(seq_h)))

    class _Seq:                                   # minimal holder so api helpers can use the handle
        h = seq_h
        nseq = len(lens)
    seq = _Seq()
    A = api.MerylExactLookup.count_sequence(ctx, K, seq)
    R = api.MerylExactLookup.synth_reads(ctx, A, COV_PER_COPY, int(total * 0.67), 7)
    if not os.environ.get("MFN_NO_JOINT"):
        R.fuse(A)                                # one 128-byte line per window instead of two
    restricted = not os.environ.get("MFN_NO_JOINT") and os.environ.get("MFN_BENCH_RESTRICT", "1") != "0" and args.mode == "hist"
    if restricted:
        # the table the range-partitioned scan streams: only the k-mers of this rank's shard of the assembly
        R.restrict(A, seq, srank, sworld)
    ctx.sync()
    setup_s = time.time() - t0
    ai, ri = A.info, R.info
    ji = R.joint_info
    if ji.fused:
        config["joint_table"] = {"kmers": int(ji.n_kmers), "GB": round(ji.device_bytes / 1e9, 2),
                                 "slot_bits": int(ji.slot_bits), "spill_frac": round(ji.n_spill / max(1, ji.n_kmers), 4),
                                 "stash": int(ji.n_stash), "bigvalue": int(ji.n_bigvalue)}
    if restricted:
        hi = R.restricted_info
        config["restricted_table"] = {"kmers": int(hi.n_kmers), "GB": round(hi.device_bytes / 1e9, 2), "lines": int(hi.lines),
                                      "tag_bits": int(hi.tag_bits), "spill_frac": round(hi.n_spill / max(1, hi.n_kmers), 4),
                                      "stash": int(hi.n_stash), "scope": f"k-mers of shard {srank} of {sworld} (each rank its own)"}
    config["tables"] = {"asm_kmers": int(ai.n_kmers), "read_kmers": int(ri.n_kmers),
                        "asm_GB": round(ai.device_bytes / 1e9, 2), "read_GB": round(ri.device_bytes / 1e9, 2),
                        "read_level2_frac": round(ri.n_level2 / max(1, ri.n_kmers), 4),
                        "slot_bits": int(ri.slot_bits), "setup_s": round(setup_s, 1)}

    if args.mode == "polish":
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_polish as bench_modes
        rc = bench_modes.polish(args, sys.modules[__name__], ctx, api, lib, torch, dist, rank, world, local, dev, asm_np, lens, off,
                                R, A, config, setup_s)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return rc

    acc = api.Accum(ctx, len(lens))
    ptr, nwords = acc.device_buffer()
    acc_t = device_tensor_view(ptr, nwords)
    km = api.KModel(PEAK)

    def step_resident():
        acc.reset()
        api.hist_resident(ctx, R, A, seq, None, acc, srank, sworld)
        if world > 1:
            dist.all_reduce(acc_t)

    def step_e2e():
        acc.reset()
        api.L.check(lib.mfn_hist_host(ctx.h, R.h, A.h, bases_ptr, off.ctypes.data_as(C.c_void_p), len(lens),
                                      None, srank, sworld, acc.h))
        if world > 1:
            dist.all_reduce(acc_t)
        per, pairs = acc.fetch()                 # D2H of the reduced histogram + counters
        res = api.HistResult([f"rec{i}" for i in range(len(lens))], per, pairs, K, km)
        return res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, sampler=None, kernel_events=False):
        for _ in range(warmup):
            fn()
        barrier()
        if sampler:
            sampler.start()
        l0 = ctx.launches
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.time()
        e0.record()
        out = None
        for i in range(steps):
            evs[i][0].record()
            out = fn()
            evs[i][1].record()
        e1.record()
        barrier()
        wall = time.time() - w0
        clocks = sampler.stop() if sampler else None
        ms = e0.elapsed_time(e1)
        per_step = [a.elapsed_time(b) for a, b in evs]
        t = torch.tensor([ms, wall * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), per_step, ctx.launches - l0, clocks, out

    # kernel-only durations for the roofline (own events, no reset / collective): the whole pass
    # and, through the library's per-pass events, pass 1 (scatter) and pass 2 (probe) separately
    def kernel_only_ms(n):
        acc.reset()
        ts = []
        ctx.time_passes(True)
        for _ in range(n):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            api.hist_resident(ctx, R, A, seq, None, acc, srank, sworld)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        sc, pr, ns = ctx.pass_times()
        ctx.time_passes(False)
        return ts, (sc / ns if ns else None), (pr / ns if ns else None)

    sampler = ClockSampler(local) if rank == 0 else None
    ms, wall_ms, per_step, launches, clocks, _ = timed(step_resident, args.steps, max(3, args.warmup), sampler)
    value = total * args.steps / (ms / 1e3)
    kts, scatter_ms, probe_ms = kernel_only_ms(max(3, args.steps))
    k_ms = sum(kts) / len(kts)

    # Auxiliary weak-scaling view (N > 1 only): every rank scans ALL tiles against its replica of the
    # tables, i.e. N times the windows of the headline run, same collective.  The headline `value`
    # stays the strong-scaling one BASELINE.json asks for (one 3 Gbp assembly sharded over N GPUs).
    weak = None
    if world > 1 and not restricted:
        try:
            def step_weak():
                acc.reset()
                api.hist_resident(ctx, R, A, seq, None, acc, 0, 1)
                dist.all_reduce(acc_t)
            w_ms = timed(step_weak, args.steps, 3)[0]
            weak = {"value": world * total * args.steps / (w_ms / 1e3), "unit": "bases/s", "ms_per_step": w_ms / args.steps,
                    "workload": f"{world} x {total / 1e9:.2f} Gbp: every rank scans the whole assembly (tables replicated)"}
        except Exception as e:                        # never let the auxiliary number take the headline down
            weak = {"error": str(e)[:200]}

    e2e_steps = max(1, min(args.steps, 5))
    h2d0 = None
    def step_e2e_counted():
        nonlocal h2d0
        if h2d0 is None:
            h2d0 = ctx.h2d_bytes               # first timed step (warm-up steps call step_e2e directly)
        return step_e2e()
    for _ in range(3):
        step_e2e()
    e_ms, e_wall_ms, _, e_launches, _, res = timed(step_e2e_counted, e2e_steps, 0)
    h2d_per_step = (ctx.h2d_bytes - h2d0) // e2e_steps
    # e2e ends on the host (fetch + K* replay), so wall clock is the honest figure
    e2e_value = total * e2e_steps / (e_wall_ms / 1e3)

    # result checks: resident and host-buffer paths agree; counters consistent
    acc.reset()
    api.hist_resident(ctx, R, A, seq, None, acc, srank, sworld)
    if world > 1:
        dist.all_reduce(acc_t)
    per2, pairs2 = acc.fetch()
    same = (len(pairs2) == len(res.pairs) and bool((pairs2 == res.pairs).all()) and bool((per2 == res.per_seq).all()))
    # Full-size parity, untimed: the same shard once more with the DIRECT scan (k_scan: one complete lookup per
    # window in the full fused table, no range partitioning, no restricted table, no side lists) -- a different
    # kernel, a different table and a different accumulation order must give the same integers.
    direct_same, direct_ms = None, None
    if not args.no_direct_check:
        prev = os.environ.get("MFN_SWEEP")
        os.environ["MFN_SWEEP"] = "0"
        try:
            acc.reset()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            api.hist_resident(ctx, R, A, seq, None, acc, srank, sworld)
            b.record()
            if world > 1:
                dist.all_reduce(acc_t)
            per3, pairs3 = acc.fetch()
            direct_ms = a.elapsed_time(b)
            direct_same = (len(pairs3) == len(pairs2) and bool((pairs3 == pairs2).all()) and bool((per3 == per2).all()))
        finally:
            if prev is None:
                del os.environ["MFN_SWEEP"]
            else:
                os.environ["MFN_SWEEP"] = prev
    summary, _, _ = res.report()

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak_gbs, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak_gbs, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        bases_rank = total / sworld
        achieved = bases_rank * BYTES_PER_BASE / (k_ms / 1e3) / 1e9
        swept = scatter_ms is not None
        # DRAM bytes one launch of the pair has to move by construction (per rank): every window's 8-byte entry
        # written by pass 1 and read by pass 2, the table the probe pass streams once, the packed sequence.
        # The table does NOT shrink with N unless it is restricted to the rank's shard.
        table_bytes = (int(R.restricted_info.lines) if restricted else int(ji.lines) if ji.fused else 0) * 128
        windows_rank = int(res.report()[0].kasm) / sworld if sworld > 1 else int(res.report()[0].kasm)
        model = windows_rank * 16 + table_bytes + bases_rank * 0.375 if swept else None
        traffic, traffic_src = model, "model: windows x 16 B (entry written + read) + table lines x 128 B + 0.375 B/base of packed sequence"
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp) and sworld == 1:
            tj = json.load(open(tp))
            # the ncu figure only stands for the configuration it was captured on
            if tj.get("dram_bytes_per_base") and bool(tj.get("restricted_table")) == bool(restricted):
                traffic = tj["dram_bytes_per_base"] * bases_rank
                traffic_src = tj.get("source")
        kernel_name = ("k_sweep_scatter3<21,GEO,TILES> + k_sweep_probe<false> (range-partitioned scan of the fused table restricted to the rank's shard)" if swept and restricted
                       else "k_sweep_scatter3 / k_sweep_scatter<21,false> + k_sweep_probe<false> (range-partitioned scan of the full fused table)" if swept
                       else "k_scan<2> (joint table)" if ji.fused else "k_scan<0> (two tables)")
        out = {"metric": "bases/sec (-hist k=21)", "value": value, "unit": "bases/s", "n_gpus": world,
               "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps,
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64",
               "data": "synthetic", "config": config,
               "e2e": {"value": e2e_value, "unit": "bases/s", "h2d_bytes_per_step": int(h2d_per_step) * world,
                       "host_input_bytes_per_step": int(total),
                       "d2h_bytes_per_step": int(nwords * 8), "ms_per_step": e_wall_ms / e2e_steps, "steps": e2e_steps},
               "gpu_launches": int(launches), "gpu_launches_e2e": int(e_launches),
               "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                            "frac": achieved / peak_gbs, "traffic": traffic, "peak_source": peak_src,
                            "traffic_source": traffic_src, "kernel": kernel_name, "kernel_ms": k_ms,
                            "pass_ms": {"scatter": scatter_ms, "probe": probe_ms} if swept else None,
                            "bases_per_launch": bases_rank, "algorithmic_bytes_per_base": BYTES_PER_BASE,
                            "traffic_model": model,
                            # how close the pair runs to the time its OWN DRAM traffic needs at the copy peak (the contract
                            # `frac` above counts 64.25 algorithmic bytes per base, which the sweep does not move)
                            "frac_of_own_traffic": (traffic / 1e9 / peak_gbs) / (k_ms / 1e3) if traffic else None},
               "clocks": clocks, "weak_scaling": weak,
               "result": {"kasm": int(summary.kasm), "kmissing": int(summary.kmissing),
                          "qv_star": summary.qv_star, "resident_equals_e2e": bool(same),
                          "sweep_equals_direct": direct_same, "direct_scan_ms": direct_ms}}
        if world == 1 and not args.no_cpu_baseline:
            try:
                cb, _ = cpu_baseline(ctx, asm_np, lens, args.sample_mbp, repeats=2)
            except Exception as e:                      # the headline line must come out whatever the CPU leg does
                cb = {"value": None, "unit": "bases/s", "cores": os.cpu_count() or 1, "kind": "reference",
                      "sample": "failed: %s" % (str(e)[:200],)}
            out["cpu_baseline"] = cb
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
