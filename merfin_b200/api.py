"""Host-side mirror of the reference's interface for the k-mer multiplicity path.

Thin Python over the C ABI (merfin_b200.lib): names follow the reference
(merylExactLookup::load/value, processHistogram/reportHistogram, processDump,
computeCompleteness, varMer::score) so the parity tests read like tests of the
reference.  All compute happens in libmerfin_b200.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import lib as L


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """One CUDA device (mfn_ctx)."""

    def __init__(self, device: int = 0):
        self._lib = L.load()
        self.h = C.c_void_p()
        L.check(self._lib.mfn_ctx_create(C.byref(self.h), device))

    def set_stream(self, cuda_stream: int | None):
        L.check(self._lib.mfn_ctx_set_stream(self.h, C.c_void_p(cuda_stream or 0)))

    def sync(self):
        L.check(self._lib.mfn_ctx_sync(self.h))

    @property
    def launches(self) -> int:
        return int(self._lib.mfn_ctx_launch_count(self.h))

    @property
    def h2d_bytes(self) -> int:
        return int(self._lib.mfn_ctx_h2d_bytes(self.h))

    def set_sweep(self, mode: int):
        """-1 automatic, 0 never, 1 whenever possible: which -hist scan runs (mfn_ctx_set_sweep)."""
        L.check(self._lib.mfn_ctx_set_sweep(self.h, mode))

    def time_passes(self, enable: bool = True):
        """Bracket the passes of every range-partitioned -hist scan with CUDA events."""
        L.check(self._lib.mfn_ctx_time_passes(self.h, 1 if enable else 0))

    def pass_times(self):
        """(scatter ms, probe ms, scans) summed over the scans since the last call."""
        a, b, n = C.c_double(), C.c_double(), C.c_uint32()
        L.check(self._lib.mfn_ctx_pass_times(self.h, C.byref(a), C.byref(b), C.byref(n)))
        return a.value, b.value, n.value

    def close(self):
        if self.h:
            self._lib.mfn_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class KModel:
    """-peak / -prob parameters of merfinGlobal::getK (merfin-globals.C:66-98)."""

    def __init__(self, peak: float, probK=(), probP=()):
        self.K = np.ascontiguousarray(probK, dtype=np.uint32)
        self.P = np.ascontiguousarray(probP, dtype=np.float64)
        self.c = L.KModel(float(peak), len(self.K), self.K.ctypes.data_as(C.POINTER(C.c_uint32)),
                          self.P.ctypes.data_as(C.POINTER(C.c_double)))

    @staticmethod
    def load_Kmetric(peak: float, path: str | None) -> "KModel":
        """-prob CSV: one 'readK,prob' line per multiplicity; other lines are skipped
        (merfin-globals.C:21-62)."""
        Ks, Ps = [], []
        if path:
            for line in open(path):
                w = [x for x in line.rstrip("\n").split(",") if x != ""]
                if len(w) == 2:
                    Ks.append(int(w[0]))
                    Ps.append(float(w[1]))
        return KModel(peak, Ks, Ps)

    def getK(self, v_read: int, v_asm: int):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        L.load().mfn_getK(C.byref(self.c), v_read, v_asm, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value


class MerylExactLookup:
    """A k-mer -> value table resident in HBM (replaces merylExactLookup, kmers-lookup.H:79-380)."""

    def __init__(self, ctx: Context, handle):
        self.ctx, self.h = ctx, handle
        self._lib = L.load()

    @classmethod
    def load(cls, ctx: Context, meryl_dir: str, minV: int = 0, maxV: int = 0xFFFFFFFF):
        h = C.c_void_p()
        L.check(L.load().mfn_table_load(ctx.h, meryl_dir.encode(), minV, maxV, C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def from_kmers(cls, ctx: Context, k: int, kmers, values, minV: int = 0, maxV: int = 0xFFFFFFFF):
        km = np.ascontiguousarray(kmers, dtype=np.uint64)
        vv = np.ascontiguousarray(values, dtype=np.uint32)
        assert len(km) == len(vv)
        h = C.c_void_p()
        L.check(L.load().mfn_table_from_kmers(ctx.h, k, _ptr(km), _ptr(vv), len(km), minV, maxV, C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def count_sequence(cls, ctx: Context, k: int, seq: "SeqSet"):
        h = C.c_void_p()
        L.check(L.load().mfn_table_count_seq(ctx.h, k, seq.h, C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def synth_reads(cls, ctx: Context, asm: "MerylExactLookup", cov_per_copy: float, n_error: int, seed: int):
        h = C.c_void_p()
        L.check(L.load().mfn_table_synth_reads(ctx.h, None, asm.h, cov_per_copy, n_error, seed, C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def restrict_stream(cls, ctx: Context, meryl_dir: str, asm: "MerylExactLookup", seq, shard: int = 0, nshards: int = 1,
                        minV: int = 0, maxV: int = 0xFFFFFFFF):
        """A read-table handle whose database was streamed through the device into the restricted fused
        table of one shard of `seq` (mfn_tables_restrict_stream): serves -hist of that shard only.
        Returns (handle, StreamStats)."""
        h, st = C.c_void_p(), L.StreamStats()
        L.check(L.load().mfn_tables_restrict_stream(ctx.h, meryl_dir.encode(), minV, maxV, asm.h, seq.h, shard, nshards,
                                                    C.byref(h), C.byref(st)))
        return cls(ctx, h), st

    @classmethod
    def restrict_synth(cls, ctx: Context, asm: "MerylExactLookup", seq, cov_per_copy: float, n_error: int, seed: int,
                       chunk_kmers: int = 0, shard: int = 0, nshards: int = 1):
        """The synthetic read database of synth_reads (same draws), streamed: its n_error error k-mers are
        generated in chunks and semi-joined against the shard's k-mers (mfn_tables_restrict_synth)."""
        h, st = C.c_void_p(), L.StreamStats()
        L.check(L.load().mfn_tables_restrict_synth(ctx.h, asm.h, seq.h, shard, nshards, cov_per_copy, n_error, seed, chunk_kmers,
                                                   C.byref(h), C.byref(st)))
        return cls(ctx, h), st

    @property
    def info(self) -> L.TableInfo:
        i = L.TableInfo()
        L.check(self._lib.mfn_table_info(self.h, C.byref(i)))
        return i

    def nKmers(self) -> int:
        return int(self.info.n_kmers)

    def values(self, kmers) -> np.ndarray:
        km = np.ascontiguousarray(kmers, dtype=np.uint64)
        out = np.zeros(len(km), dtype=np.uint32)
        L.check(self._lib.mfn_table_lookup(self.ctx.h, self.h, _ptr(km), len(km), _ptr(out)))
        return out

    def value(self, kmer: int) -> int:
        return int(self.values([kmer])[0])

    def values_wide(self, kmers_lo, kmers_hi) -> np.ndarray:
        """merylExactLookup::value for a table of k = 33..64: k-mer i is kmers_hi[i] << 64 | kmers_lo[i]."""
        lo = np.ascontiguousarray(kmers_lo, dtype=np.uint64)
        hi = np.ascontiguousarray(kmers_hi, dtype=np.uint64)
        out = np.zeros(len(lo), dtype=np.uint32)
        L.check(self._lib.mfn_table_lookup_wide(self.ctx.h, self.h, _ptr(lo), _ptr(hi), len(lo), _ptr(out)))
        return out

    def fuse(self, asm: "MerylExactLookup"):
        """Attach the fused (self, asm) table: scans over this pair then cost one probe per window."""
        L.check(self._lib.mfn_tables_fuse(self.ctx.h, self.h, asm.h))

    def unfuse(self):
        L.check(self._lib.mfn_tables_unfuse(self.h))

    def restrict(self, asm: "MerylExactLookup", seq, shard: int = 0, nshards: int = 1):
        """Attach the fused table restricted to the k-mers of one shard of `seq` (what a -hist scan of
        that shard streams); lookups of other k-mers fall back to the per-database tables."""
        L.check(self._lib.mfn_tables_restrict(self.ctx.h, self.h, asm.h, seq.h, shard, nshards))

    def unrestrict(self):
        L.check(self._lib.mfn_tables_unrestrict(self.h))

    @property
    def restricted_info(self) -> L.JointInfo:
        i = L.JointInfo()
        L.check(self._lib.mfn_restricted_info(self.h, C.byref(i)))
        return i

    @property
    def joint_info(self) -> L.JointInfo:
        i = L.JointInfo()
        L.check(self._lib.mfn_joint_info(self.h, C.byref(i)))
        return i

    def joint_values(self, kmers):
        km = np.ascontiguousarray(kmers, dtype=np.uint64)
        vr = np.zeros(len(km), dtype=np.uint32)
        va = np.zeros(len(km), dtype=np.uint32)
        L.check(self._lib.mfn_joint_lookup(self.ctx.h, self.h, _ptr(km), len(km), _ptr(vr), _ptr(va)))
        return vr, va

    def export(self):
        n = self.nKmers()
        km = np.zeros(max(n, 1), dtype=np.uint64)
        vv = np.zeros(max(n, 1), dtype=np.uint32)
        got = C.c_uint64()
        L.check(self._lib.mfn_table_export(self.ctx.h, self.h, _ptr(km), _ptr(vv), n, C.byref(got)))
        return km[:got.value], vv[:got.value]

    def close(self):
        if self.h:
            self._lib.mfn_table_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def concat_records(records):
    """records: iterable of bytes / (ident, bytes) -> (uint8 array, uint64 offsets, idents)."""
    idents, chunks = [], []
    for i, r in enumerate(records):
        if isinstance(r, tuple):
            idents.append(r[0])
            r = r[1]
        else:
            idents.append(str(i))
        if isinstance(r, np.ndarray):
            r = r.tobytes()
        chunks.append(bytes(r))
    off = np.zeros(len(chunks) + 1, dtype=np.uint64)
    if chunks:
        off[1:] = np.cumsum([len(c) for c in chunks], dtype=np.uint64)
    bases = np.frombuffer(b"".join(chunks), dtype=np.uint8) if chunks else np.zeros(0, dtype=np.uint8)
    return bases, off, idents


class SeqSet:
    """Records packed 2 bits/base and resident in HBM (mfn_seq)."""

    def __init__(self, ctx: Context, records):
        self.ctx = ctx
        self._lib = L.load()
        bases, off, self.idents = concat_records(records)
        self.off = off
        self.h = C.c_void_p()
        L.check(self._lib.mfn_seq_create(ctx.h, _ptr(bases), _ptr(off), len(off) - 1, C.byref(self.h)))

    @property
    def nseq(self):
        return len(self.off) - 1

    @property
    def total_bases(self):
        return int(self.off[-1])

    def close(self):
        if self.h:
            self._lib.mfn_seq_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Accum:
    """Joint (v_read, v_asm) histogram + per-record counters on the device (mfn_accum)."""

    def __init__(self, ctx: Context, nseq: int):
        self.ctx, self.nseq = ctx, nseq
        self._lib = L.load()
        self.h = C.c_void_p()
        L.check(self._lib.mfn_accum_create(ctx.h, nseq, C.byref(self.h)))

    def reset(self):
        L.check(self._lib.mfn_accum_reset(self.ctx.h, self.h))

    def device_buffer(self):
        p, n = C.c_void_p(), C.c_uint64()
        L.check(self._lib.mfn_accum_device_buffer(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def fetch(self, dense=True, sparse=True):
        cap = 1 << 16
        while True:
            pairs = np.zeros(cap, dtype=PAIR_DTYPE)
            per = np.zeros(max(self.nseq, 1), dtype=COUNTS_DTYPE)
            n = C.c_uint64()
            st = self._lib.mfn_accum_fetch(self.ctx.h, self.h, int(dense), int(sparse), _ptr(per), _ptr(pairs),
                                           cap, C.byref(n))
            if st == L.MFN_ERR_CAPACITY and n.value > cap:
                cap = int(n.value)
                continue
            L.check(st)
            return per[:self.nseq], pairs[:n.value]

    def read(self):
        """The flat buffer (all-reduced or not) as a host array of uint64 words."""
        _, n = self.device_buffer()
        buf = np.zeros(n, dtype=np.uint64)
        L.check(self._lib.mfn_accum_read(self.ctx.h, self.h, _ptr(buf)))
        return buf

    def sparse(self):
        """This rank's pairs outside the dense tier."""
        n = C.c_uint64()
        st = self._lib.mfn_accum_sparse_fetch(self.ctx.h, self.h, None, 0, C.byref(n))
        if st == L.MFN_OK:
            return np.zeros(0, dtype=PAIR_DTYPE)
        if st != L.MFN_ERR_CAPACITY:
            L.check(st)
        pairs = np.zeros(int(n.value), dtype=PAIR_DTYPE)
        L.check(self._lib.mfn_accum_sparse_fetch(self.ctx.h, self.h, _ptr(pairs), len(pairs), C.byref(n)))
        return pairs[:n.value]

    def fetch_gathered(self, all_gather=None):
        """Multi-rank fetch after the all-reduce of the flat buffer: all_gather(array) -> list of every
        rank's array (called only when the reduced header says side-table pairs exist, and then on every
        rank, since every rank reads the same header)."""
        buf = self.read()
        sparse = np.zeros(0, dtype=PAIR_DTYPE)
        if buf[1]:
            mine = self.sparse()
            sparse = np.concatenate(all_gather(mine)) if all_gather else mine
        return decode_gathered(buf, self.nseq, sparse)

    def close(self):
        if self.h:
            self._lib.mfn_accum_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


PAIR_DTYPE = np.dtype([("v_read", np.uint32), ("v_asm", np.uint32), ("count", np.uint64)])
COUNTS_DTYPE = np.dtype([("kasm", np.uint64), ("kmissing", np.uint64)])


def decode_gathered(buf, nseq, sparse):
    """mfn_accum_decode_gathered: a (reduced) flat buffer + every rank's side-table pairs -> (per_seq, pairs)."""
    lib = L.load()
    sparse = np.ascontiguousarray(sparse, dtype=PAIR_DTYPE)
    cap = 1 << 16
    while True:
        pairs = np.zeros(cap, dtype=PAIR_DTYPE)
        per = np.zeros(max(nseq, 1), dtype=COUNTS_DTYPE)
        n = C.c_uint64()
        st = lib.mfn_accum_decode_gathered(_ptr(buf), nseq, _ptr(sparse) if len(sparse) else None, len(sparse),
                                           _ptr(per), _ptr(pairs), cap, C.byref(n))
        if st == L.MFN_ERR_CAPACITY and n.value > cap:
            cap = int(n.value)
            continue
        L.check(st)
        return per[:nseq], pairs[:n.value]


@dataclass
class HistResult:
    """What processHistogram/outputHistogram/reportHistogram produce (merfin-histogram.C)."""
    idents: list
    per_seq: np.ndarray          # COUNTS_DTYPE per record
    pairs: np.ndarray            # PAIR_DTYPE, the joint histogram
    k: int
    kmodel: KModel

    def report(self):
        lib = L.load()
        s = L.HistSummary()
        cap = 4096
        while True:
            undr = np.zeros(cap, dtype=np.uint64)
            over = np.zeros(cap, dtype=np.uint64)
            st = lib.mfn_hist_report(_ptr(self.pairs), len(self.pairs), C.byref(self.kmodel.c), self.k,
                                     _ptr(undr), _ptr(over), cap, C.byref(s))
            if st == L.MFN_ERR_CAPACITY:
                cap = max(s.n_undr, s.n_over) + 1
                continue
            L.check(st)
            return s, undr[:s.n_undr], over[:s.n_over]

    def file_text(self) -> bytes:
        """The -hist output file (merfin-histogram.C:151-163)."""
        lib = L.load()
        s, undr, over = self.report()
        need = C.c_uint64()
        lib.mfn_hist_format(_ptr(undr), len(undr), _ptr(over), len(over), None, 0, C.byref(need))
        buf = C.create_string_buffer(need.value + 1)
        L.check(lib.mfn_hist_format(_ptr(undr), len(undr), _ptr(over), len(over), buf, need.value + 1, C.byref(need)))
        return buf.raw[:need.value]

    def summary_text(self) -> str:
        """The block reportHistogram prints to stderr (merfin-histogram.C:167-172)."""
        s, _, _ = self.report()
        return ("K-mers not found in reads (missing) : %d\n"
                "K-mers overly represented in assembly: %.2f\n"
                "K-mers found in the assembly: %d\n"
                "Missing QV: %.2f\nMerfin QV*: %.2f\n") % (s.kmissing, s.koverCpy, s.kasm, s.qv_missing, s.qv_star)

    def per_seq_text(self) -> str:
        """Per-record lines outputHistogram prints (merfin-histogram.C:128-133), in input order."""
        import math
        out, cum = [], 0
        for ident, c in zip(self.idents, self.per_seq):
            kasm, kmiss = int(c["kasm"]), int(c["kmissing"])
            cum += kmiss
            base = (kmiss / kasm) if kasm else float("nan")
            try:
                qv = -10.0 * math.log10(1.0 - math.pow(1.0 - base, 1.0 / self.k))
            except (ValueError, ZeroDivisionError):
                qv = float("inf") if base == 0 else float("nan")
            out.append("%s\t%d\t%d\t%d\t%.2f" % (ident, kmiss, cum, kasm, qv))
        return "\n".join(out) + "\n"


def hist(ctx: Context, reads: MerylExactLookup, asm: MerylExactLookup, records, kmodel: KModel) -> HistResult:
    """merfin -hist for a batch of records given as HOST buffers (mfn_hist)."""
    lib = L.load()
    bases, off, idents = concat_records(records)
    nseq = len(off) - 1
    per = np.zeros(max(nseq, 1), dtype=COUNTS_DTYPE)
    cap = 1 << 16
    while True:
        pairs = np.zeros(cap, dtype=PAIR_DTYPE)
        n = C.c_uint64()
        st = lib.mfn_hist(ctx.h, reads.h, asm.h, _ptr(bases), _ptr(off), nseq, C.byref(kmodel.c), _ptr(per),
                          _ptr(pairs), cap, C.byref(n))
        if st == L.MFN_ERR_CAPACITY and n.value > cap:
            cap = int(n.value)
            continue
        L.check(st)
        break
    return HistResult(idents, per[:nseq], pairs[:n.value], int(reads.info.k), kmodel)


def hist_resident(ctx: Context, reads, asm, seq: SeqSet, kmodel: KModel | None, acc: Accum,
                  shard: int = 0, nshards: int = 1):
    L.check(L.load().mfn_hist_resident(ctx.h, reads.h, asm.h, seq.h, C.byref(kmodel.c) if kmodel else None,
                                       shard, nshards, acc.h))


def dump(ctx: Context, reads, asm, records):
    """merfin -dump probes: per position (v_read, v_asm, valid) for every record (mfn_dump)."""
    bases, off, idents = concat_records(records)
    n = int(off[-1])
    vr = np.zeros(max(n, 1), dtype=np.uint32)
    va = np.zeros(max(n, 1), dtype=np.uint32)
    ok = np.zeros(max(n, 1), dtype=np.uint8)
    L.check(L.load().mfn_dump(ctx.h, reads.h, asm.h, _ptr(bases), _ptr(off), len(off) - 1, _ptr(vr), _ptr(va), _ptr(ok)))
    return idents, off, vr[:n], va[:n], ok[:n]


def dump_text(idents, off, vr, va, ok, kmodel: KModel, k: int):
    """Text of the -dump file (merfin-dump.C:87-93) and the per-record stderr lines (:100-101)."""
    lib = L.load()
    out, per, cm, ca = [], [], 0, 0
    rk, ak, pr = C.c_double(), C.c_double(), C.c_double()
    cache = {}
    for i, ident in enumerate(idents):
        a, b = int(off[i]), int(off[i + 1])
        kasm = kmiss = 0
        for p in range(a, b):
            if not ok[p]:
                continue
            key = (int(vr[p]), int(va[p]))
            if key not in cache:
                lib.mfn_getK(C.byref(kmodel.c), key[0], key[1], C.byref(rk), C.byref(ak), C.byref(pr))
                cache[key] = (rk.value, ak.value, lib.mfn_getKmetric(rk.value, ak.value))
            r, s, m = cache[key]
            kasm += 1
            kmiss += r == 0
            if r != 0.0 or s != 0.0 or m != 0.0:
                out.append("%s\t%d\t%.2f\t%.2f\t%.2f\n" % (ident, p - a, r, s, m))
        cm += kmiss
        ca += kasm
        per.append("%s\t%d\t%d\t%d" % (ident, kmiss, cm, ca))
    return "".join(out).encode(), "\n".join(per) + "\n"


def score_paths(ctx: Context, reads, asm, paths):
    """Probe every window of a batch of variant path strings (mfn_score_paths)."""
    bases, off, _ = concat_records(paths)
    n = int(off[-1])
    vr = np.zeros(max(n, 1), dtype=np.uint32)
    va = np.zeros(max(n, 1), dtype=np.uint32)
    ok = np.zeros(max(n, 1), dtype=np.uint8)
    L.check(L.load().mfn_score_paths(ctx.h, reads.h, asm.h, _ptr(bases), _ptr(off), len(off) - 1,
                                     _ptr(vr), _ptr(va), _ptr(ok)))
    return off, vr[:n], va[:n], ok[:n]


class Vcf:
    """A loaded VCF (vcfFile::load, vcf.C:93-151) whose variants can be clustered (mergeChrPosGT,
    vcf.C:156-246) and scored record by record (processVariants, merfin-variants.C:131-310)."""

    def __init__(self, path: str):
        self.h = C.c_void_p()
        L.check(L.load().mfn_vcf_load(os.fsencode(path), C.byref(self.h)))

    def cluster(self, k: int, comb: int = 15, nosplit: bool = False):
        L.check(L.load().mfn_vcf_cluster(self.h, k, comb, int(nosplit)))

    @property
    def info(self) -> L.VcfInfo:
        i = L.VcfInfo()
        L.check(L.load().mfn_vcf_info(self.h, C.byref(i)))
        return i

    def header(self) -> bytes:
        lib = L.load()
        need = C.c_uint64()
        L.check(lib.mfn_vcf_header(self.h, None, 0, C.byref(need)))
        buf = C.create_string_buffer(need.value + 1)
        L.check(lib.mfn_vcf_header(self.h, buf, need.value, C.byref(need)))
        return buf.raw[:need.value]

    def clusters_of(self, ident: str) -> int:
        return int(L.load().mfn_vcf_clusters_of(self.h, ident.encode()))

    def close(self):
        if self.h:
            L.load().mfn_vcf_free(self.h)
            self.h = C.c_void_p()


def variants_process(ctx: Context, reads, asm, vcf: Vcf, ident: str, seq, mode: str, kmodel: KModel, comb: int = 15,
                     shard: int = 0, nshards: int = 1, debug: bool = False, seq_len: int | None = None):
    """processVariants for one sequence record -> the VCF text merfin appends for it (bytes), and the
    -debug records when asked.  `seq`: bytes, or an integer address of seq_len bytes."""
    lib = L.load()
    res, dbg = C.c_void_p(), C.c_void_p()
    rlen, dlen, vid = C.c_uint64(), C.c_uint64(), C.c_uint64()
    if isinstance(seq, (bytes, bytearray)):
        keep = bytes(seq)
        ptr, n = C.cast(C.c_char_p(keep), C.c_void_p), len(keep)
    else:
        ptr, n = C.c_void_p(int(seq)), int(seq_len)
    L.check(lib.mfn_variants_process(ctx.h, reads.h, asm.h, vcf.h, ident.encode(), ptr, n, L.MFN_MODE[mode], comb,
                                     C.byref(kmodel.c), shard, nshards, C.byref(res), C.byref(rlen),
                                     C.byref(dbg) if debug else None, C.byref(dlen) if debug else None, C.byref(vid)))
    out = C.string_at(res, rlen.value) if res else b""
    lib.mfn_text_free(res)
    d = None
    if debug:
        d = C.string_at(dbg, dlen.value) if dbg else b""
        lib.mfn_text_free(dbg)
    return (out, d) if debug else out


def completeness(ctx: Context, reads_dir: str, asm_dir: str, kmodel: KModel):
    """merfin -completeness (merfin-completeness.C:48-144): returns (total readK, undercopy, pairs)."""
    lib = L.load()
    cap = 1 << 16
    while True:
        pairs = np.zeros(cap, dtype=PAIR_DTYPE)
        n = C.c_uint64()
        st = lib.mfn_completeness(ctx.h, reads_dir.encode(), asm_dir.encode(), _ptr(pairs), cap, C.byref(n))
        if st == L.MFN_ERR_CAPACITY and n.value > cap:
            cap = int(n.value)
            continue
        L.check(st)
        break
    pairs = pairs[:n.value]
    t, u = C.c_double(), C.c_double()
    L.check(lib.mfn_completeness_report(_ptr(pairs), len(pairs), C.byref(kmodel.c), C.byref(t), C.byref(u)))
    return t.value, u.value, pairs


def completeness_resident(ctx: Context, reads: MerylExactLookup, asm: MerylExactLookup, kmodel: KModel):
    """computeCompleteness over two resident tables (raw values): (total readK, total undercopy, pairs)."""
    lib = L.load()
    acc = Accum(ctx, 0)
    L.check(lib.mfn_completeness_resident(ctx.h, reads.h, asm.h, acc.h))
    _, pairs = acc.fetch()
    acc.close()
    tot, undr = C.c_double(), C.c_double()
    L.check(lib.mfn_completeness_report(_ptr(pairs), len(pairs), C.byref(kmodel.c), C.byref(tot), C.byref(undr)))
    return tot.value, undr.value, pairs


def completeness_text(total: float, undr: float) -> str:
    return ("TOTAL readK:   %15.2f\nTOTAL undrcpy:    %15.5f\nCOMPLETENESS:             %0.5f\n"
            % (total, undr, 1.0 - undr / total))


def meryl_info(path: str) -> dict:
    """Header of a meryl database (mfn_meryl_info): k, layout and the v03 statistics."""
    info = L.MerylInfo()
    L.check(L.load().mfn_meryl_info(path.encode(), C.byref(info)))
    return {f: int(getattr(info, f)) for f, _ in L.MerylInfo._fields_}


def meryl_histogram(path: str):
    """(values, counts) of the stored value histogram (mfn_meryl_histogram)."""
    lib = L.load()
    n = C.c_uint64(0)
    st = lib.mfn_meryl_histogram(path.encode(), None, None, 0, C.byref(n))
    if st not in (0, L.MFN_ERR_CAPACITY):
        L.check(st)
    v = np.zeros(n.value, dtype=np.uint64)
    c = np.zeros(n.value, dtype=np.uint64)
    L.check(lib.mfn_meryl_histogram(path.encode(), _ptr(v), _ptr(c), n.value, C.byref(n)))
    return v, c


def meryl_read(path: str):
    """(kmers_lo, kmers_hi, values) of every k-mer of a database in stored order, decoded on the host
    (mfn_meryl_read); a k-mer is kmers_hi << 64 | kmers_lo (k <= 64)."""
    lib = L.load()
    n = C.c_uint64(0)
    L.check(lib.mfn_meryl_read(path.encode(), None, None, None, 0, C.byref(n)))
    lo = np.zeros(n.value, dtype=np.uint64)
    hi = np.zeros(n.value, dtype=np.uint64)
    v = np.zeros(n.value, dtype=np.uint32)
    L.check(lib.mfn_meryl_read(path.encode(), _ptr(lo), _ptr(hi), _ptr(v), n.value, C.byref(n)))
    return lo, hi, v


def meryl_write(path: str, k: int, kmers, values):
    km = np.ascontiguousarray(kmers, dtype=np.uint64)
    vv = np.ascontiguousarray(values, dtype=np.uint32)
    L.check(L.load().mfn_meryl_write(path.encode(), k, _ptr(km), _ptr(vv), len(km)))
