"""ctypes binding of the CPU oracle (oracle/mfn_oracle.c).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(ROOT, "oracle", "_build", "libmfn_oracle.so")


def _build():
    src = os.path.join(ROOT, "oracle", "mfn_oracle.c")
    if (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])


class _DB(C.Structure):
    _fields_ = [("k", C.c_uint32), ("prefixSize", C.c_uint32), ("suffixSize", C.c_uint32),
                ("numFilesBits", C.c_uint32), ("numBlocksBits", C.c_uint32), ("flags", C.c_uint32),
                ("version", C.c_uint32),
                ("numUnique", C.c_uint64), ("numDistinct", C.c_uint64), ("numTotal", C.c_uint64),
                ("histLen", C.c_uint64),
                ("histValue", C.POINTER(C.c_uint64)), ("histCount", C.POINTER(C.c_uint64)),
                ("n", C.c_uint64),
                ("lo", C.POINTER(C.c_uint64)), ("hi", C.POINTER(C.c_uint64)), ("val", C.POINTER(C.c_uint32))]


class _LK(C.Structure):
    _fields_ = [("k", C.c_uint32), ("minValue", C.c_uint32), ("maxValue", C.c_uint32),
                ("valueOffset", C.c_uint32), ("valueBits", C.c_uint32), ("n", C.c_uint64),
                ("lo", C.POINTER(C.c_uint64)), ("hi", C.POINTER(C.c_uint64)), ("val", C.POINTER(C.c_uint32))]


class _KM(C.Structure):
    _fields_ = [("peak", C.c_double), ("nprob", C.c_uint32),
                ("probK", C.POINTER(C.c_uint32)), ("probP", C.POINTER(C.c_double))]


class _HIST(C.Structure):
    _fields_ = [("kasm", C.c_uint64), ("kmissing", C.c_uint64), ("koverCpy", C.c_double),
                ("undrMax", C.c_uint32), ("overMax", C.c_uint32),
                ("undr", C.POINTER(C.c_uint64)), ("over", C.POINTER(C.c_uint64))]


_build()
L = C.CDLL(_SO)
L.o_db_load.restype = C.POINTER(_DB)
L.o_db_load.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
L.o_db_free.argtypes = [C.POINTER(_DB)]
L.o_lookup_build.restype = C.POINTER(_LK)
L.o_lookup_build.argtypes = [C.POINTER(_DB), C.c_uint32, C.c_uint32]
L.o_lookup_free.argtypes = [C.POINTER(_LK)]
L.o_lookup_value.restype = C.c_uint32
L.o_lookup_value.argtypes = [C.POINTER(_LK), C.c_uint64, C.c_uint64]
L.o_lookup_values.argtypes = [C.POINTER(_LK), C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
L.o_windows.restype = C.c_uint64
L.o_windows.argtypes = [C.c_char_p, C.c_uint64, C.c_uint32] + [C.c_void_p] * 5
L.o_pairs.restype = C.c_uint64
L.o_pairs.argtypes = [C.POINTER(_LK), C.POINTER(_LK), C.c_char_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
L.o_getK.argtypes = [C.POINTER(_KM), C.c_uint32, C.c_uint32] + [C.POINTER(C.c_double)] * 3
L.o_getKmetric.restype = C.c_double
L.o_getKmetric.argtypes = [C.c_double, C.c_double]
L.o_hist_new.restype = C.POINTER(_HIST)
L.o_hist_free.argtypes = [C.POINTER(_HIST)]
L.o_hist_record.argtypes = [C.POINTER(_LK), C.POINTER(_LK), C.POINTER(_KM), C.c_char_p, C.c_uint64,
                            C.POINTER(_HIST), C.POINTER(_HIST)]
L.o_hist_add_pair.argtypes = [C.POINTER(_HIST), C.POINTER(_KM), C.c_uint32, C.c_uint32, C.c_uint64]
L.o_hist_report.restype = C.c_uint64
L.o_hist_report.argtypes = [C.POINTER(_HIST), C.c_char_p, C.c_uint64]
L.o_histoQV.restype = C.c_double
L.o_histoQV.argtypes = [C.c_double, C.c_double, C.c_uint32]
L.o_dump_record.argtypes = [C.POINTER(_LK), C.POINTER(_LK), C.POINTER(_KM), C.c_char_p, C.c_uint64,
                            C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
L.o_dump_format.restype = C.c_uint64
L.o_dump_format.argtypes = [C.c_char_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_uint64]
L.o_completeness.argtypes = [C.POINTER(_DB), C.POINTER(_DB), C.POINTER(_KM), C.c_void_p, C.c_void_p,
                             C.POINTER(C.c_double), C.POINTER(C.c_double)]
L.o_score_path.restype = C.c_uint32
L.o_score_path.argtypes = [C.POINTER(_LK), C.POINTER(_LK), C.POINTER(_KM), C.c_char_p, C.c_uint64, C.c_int,
                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p,
                           C.POINTER(C.c_uint64), C.POINTER(C.c_double)]


def _np(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


class DB:
    """A meryl database decoded by the oracle (all k-mers in file order)."""

    def __init__(self, path):
        eb = C.create_string_buffer(512)
        self.p = L.o_db_load(path.encode(), eb, 512)
        if not self.p:
            raise IOError(eb.value.decode())
        d = self.p.contents
        self.k, self.n = d.k, d.n
        self.prefixSize, self.suffixSize = d.prefixSize, d.suffixSize
        self.numFilesBits, self.numBlocksBits, self.flags, self.version = d.numFilesBits, d.numBlocksBits, d.flags, d.version
        self.numUnique, self.numDistinct, self.numTotal = d.numUnique, d.numDistinct, d.numTotal
        self.lo = _np(d.lo, d.n, np.uint64)
        self.hi = _np(d.hi, d.n, np.uint64)
        self.val = _np(d.val, d.n, np.uint32)
        self.histValue = _np(d.histValue, d.histLen, np.uint64)
        self.histCount = _np(d.histCount, d.histLen, np.uint64)

    def __del__(self):
        if getattr(self, "p", None):
            L.o_db_free(self.p)
            self.p = None


class Lookup:
    def __init__(self, db: DB, minV=0, maxV=0xFFFFFFFF):
        self.db = db
        self.p = L.o_lookup_build(db.p, minV, maxV)
        self.k = db.k
        self.n = self.p.contents.n
        self.valueOffset = self.p.contents.valueOffset

    def value(self, lo, hi=0):
        return L.o_lookup_value(self.p, lo, hi)

    def values(self, lo, hi=None):
        lo = np.ascontiguousarray(lo, dtype=np.uint64)
        out = np.zeros(len(lo), dtype=np.uint32)
        hip = None
        if hi is not None:
            hi = np.ascontiguousarray(hi, dtype=np.uint64)
            hip = hi.ctypes.data
        L.o_lookup_values(self.p, lo.ctypes.data, hip, len(lo), out.ctypes.data)
        return out

    def __del__(self):
        if getattr(self, "p", None):
            L.o_lookup_free(self.p)
            self.p = None


class KModel:
    def __init__(self, peak, probK=(), probP=()):
        self.K = np.ascontiguousarray(probK, dtype=np.uint32)
        self.P = np.ascontiguousarray(probP, dtype=np.float64)
        self.c = _KM(peak, len(self.K), self.K.ctypes.data_as(C.POINTER(C.c_uint32)),
                     self.P.ctypes.data_as(C.POINTER(C.c_double)))

    @staticmethod
    def from_file(peak, path):
        """-prob CSV: lines with exactly two comma fields (merfin-globals.C:41-58)."""
        Ks, Ps = [], []
        for line in open(path):
            w = [x for x in line.rstrip("\n").split(",") if x != ""]
            if len(w) == 2:
                Ks.append(int(w[0])); Ps.append(float(w[1]))
        return KModel(peak, Ks, Ps)

    def getK(self, vr, va):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        L.o_getK(C.byref(self.c), vr, va, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value


def windows(bases: bytes, k: int):
    n = max(0, len(bases) - k + 1)
    arrs = [np.zeros(n, dtype=np.uint64) for _ in range(4)]
    valid = np.zeros(n, dtype=np.uint8)
    m = L.o_windows(bases, len(bases), k, *[a.ctypes.data for a in arrs], valid.ctypes.data)
    assert m == n
    return (*arrs, valid)


def pairs(R: Lookup, A: Lookup, bases: bytes):
    n = max(0, len(bases) - R.k + 1)
    vr = np.zeros(n, dtype=np.uint32); va = np.zeros(n, dtype=np.uint32); ok = np.zeros(n, dtype=np.uint8)
    m = L.o_pairs(R.p, A.p, bases, len(bases), vr.ctypes.data, va.ctypes.data, ok.ctypes.data)
    assert m == n
    return vr, va, ok


class Hist:
    def __init__(self):
        self.p = L.o_hist_new()

    def add_record(self, R, A, km, bases: bytes):
        rec = _HIST()
        L.o_hist_record(R.p, A.p, C.byref(km.c), bases, len(bases), self.p, C.byref(rec))
        return rec.kasm, rec.kmissing, rec.koverCpy

    def add_pair(self, km, vr, va, count=1):
        L.o_hist_add_pair(self.p, C.byref(km.c), int(vr), int(va), int(count))

    @property
    def kasm(self): return self.p.contents.kasm
    @property
    def kmissing(self): return self.p.contents.kmissing
    @property
    def koverCpy(self): return self.p.contents.koverCpy

    def report(self) -> bytes:
        n = L.o_hist_report(self.p, None, 0)
        buf = C.create_string_buffer(n + 1)
        L.o_hist_report(self.p, buf, n + 1)
        return buf.raw[:n]

    def summary(self, k) -> str:
        """The block reportHistogram prints to stderr (merfin-histogram.C:167-172)."""
        qv = L.o_histoQV(float(self.kmissing), float(self.kasm), k)
        qvs = L.o_histoQV(self.kmissing + self.koverCpy, float(self.kasm), k)
        return ("K-mers not found in reads (missing) : %d\n"
                "K-mers overly represented in assembly: %.2f\n"
                "K-mers found in the assembly: %d\n"
                "Missing QV: %.2f\nMerfin QV*: %.2f\n") % (self.kmissing, self.koverCpy, self.kasm, qv, qvs)

    def __del__(self):
        if getattr(self, "p", None):
            L.o_hist_free(self.p)
            self.p = None


def histoQV(kval, ktot, k):
    return L.o_histoQV(float(kval), float(ktot), k)


def dump_record(R, A, km, ident: str, bases: bytes):
    n = len(bases)
    rk = np.zeros(n + 1); ak = np.zeros(n + 1); ks = np.zeros(n + 1)
    kasm, kmiss = C.c_uint64(), C.c_uint64()
    L.o_dump_record(R.p, A.p, C.byref(km.c), bases, n, rk.ctypes.data, ak.ctypes.data, ks.ctypes.data,
                    C.byref(kasm), C.byref(kmiss))
    need = L.o_dump_format(ident.encode(), n, rk.ctypes.data, ak.ctypes.data, ks.ctypes.data, None, 0)
    buf = C.create_string_buffer(need + 1)
    L.o_dump_format(ident.encode(), n, rk.ctypes.data, ak.ctypes.data, ks.ctypes.data, buf, need + 1)
    return buf.raw[:need], kasm.value, kmiss.value


def completeness(S: DB, A: DB, km):
    t64 = np.zeros(64); u64 = np.zeros(64)
    t, u = C.c_double(), C.c_double()
    L.o_completeness(S.p, A.p, C.byref(km.c), t64.ctypes.data, u64.ctypes.data, C.byref(t), C.byref(u))
    return t64, u64, t.value, u.value


def score_path(R, A, km, path: bytes, filter_mode, alt_idx=(), alt_len=(), alt_gt=(), prob0=1.0):
    k = R.k
    n = max(0, len(path) - k + 1)
    ks = np.zeros(n + 1); dk = np.zeros(n + 1)
    ai = np.ascontiguousarray(alt_idx, dtype=np.uint32)
    al = np.ascontiguousarray(alt_len, dtype=np.uint32)
    ag = np.ascontiguousarray(alt_gt, dtype=np.int32)
    nw = C.c_uint64(); prob = C.c_double(prob0)
    numM = L.o_score_path(R.p, A.p, C.byref(km.c), path, len(path), int(filter_mode),
                          ai.ctypes.data, al.ctypes.data, ag.ctypes.data, len(ai),
                          ks.ctypes.data, dk.ctypes.data, C.byref(nw), C.byref(prob))
    return numM, ks[:nw.value], dk[:nw.value], prob.value


def read_fasta(path):
    """Minimal FASTA reader for fixtures: [(ident, bytes)]; ident = header up to whitespace."""
    recs, name, chunks = [], None, []
    for line in open(path, "rb"):
        line = line.rstrip(b"\r\n")
        if line.startswith(b">"):
            if name is not None:
                recs.append((name, b"".join(chunks)))
            name = line[1:].split()[0].decode() if len(line) > 1 else ""
            chunks = []
        else:
            chunks.append(line)
    if name is not None:
        recs.append((name, b"".join(chunks)))
    return recs


_ENC = {"A": 0, "C": 1, "T": 2, "G": 3}


def kmer_from_string(s: str) -> int:
    v = 0
    for c in s:
        v = (v << 2) | _ENC[c]
    return v
