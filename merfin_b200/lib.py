"""ctypes binding of libmerfin_b200.so (the C ABI in include/merfin_b200.h).

This is the same stub a maintainer of a Python caller would write;
it adds nothing: no CPU fallback, no torch types.  Loading fails loudly when the
extension has not been built.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MFN_LIB selects another build of the same library (e.g. the experimental `make exp` one) for A/B runs
SO_PATH = os.path.abspath(os.environ["MFN_LIB"]) if os.environ.get("MFN_LIB") else os.path.join(_HERE, "libmerfin_b200.so")

MFN_OK = 0
MFN_ERR_NO_DEVICE = -1
MFN_ERR_CUDA = -2
MFN_ERR_IO = -3
MFN_ERR_FORMAT = -4
MFN_ERR_UNSUPPORTED = -5
MFN_ERR_INVALID = -6
MFN_ERR_CAPACITY = -7
MFN_ERR_K_MISMATCH = -8
MFN_ERR_NOMEM = -9


class MfnError(RuntimeError):
    def __init__(self, status: int, detail: str):
        self.status = status
        super().__init__(f"merfin_b200 error {status}: {detail}")


class TableInfo(C.Structure):
    _fields_ = [("k", C.c_uint32), ("canonical", C.c_uint32), ("slot_bits", C.c_uint32),
                ("tag_bits", C.c_uint32), ("value_bits", C.c_uint32),
                ("min_value", C.c_uint32), ("max_value", C.c_uint32), ("value_offset", C.c_uint32),
                ("n_kmers", C.c_uint64), ("n_level1", C.c_uint64), ("n_level2", C.c_uint64),
                ("n_stash", C.c_uint64), ("n_bigvalue", C.c_uint64),
                ("buckets_level1", C.c_uint64), ("buckets_level2", C.c_uint64),
                ("device_bytes", C.c_uint64)]


class JointInfo(C.Structure):
    _fields_ = [("fused", C.c_uint32), ("slot_bits", C.c_uint32), ("tag_bits", C.c_uint32),
                ("read_bits", C.c_uint32), ("asm_bits", C.c_uint32),
                ("n_kmers", C.c_uint64), ("n_home", C.c_uint64), ("n_spill", C.c_uint64),
                ("n_stash", C.c_uint64), ("n_bigvalue", C.c_uint64),
                ("lines", C.c_uint64), ("device_bytes", C.c_uint64)]


class MerylInfo(C.Structure):
    _fields_ = [("k", C.c_uint32), ("version", C.c_uint32), ("prefix_bits", C.c_uint32), ("suffix_bits", C.c_uint32),
                ("files", C.c_uint32), ("blocks_per_file", C.c_uint32), ("multiset", C.c_uint32), ("hist_len", C.c_uint32),
                ("n_unique", C.c_uint64), ("n_distinct", C.c_uint64), ("n_total", C.c_uint64)]


class KModel(C.Structure):
    _fields_ = [("peak", C.c_double), ("nprob", C.c_uint32),
                ("probK", C.POINTER(C.c_uint32)), ("probP", C.POINTER(C.c_double))]


class SeqCounts(C.Structure):
    _fields_ = [("kasm", C.c_uint64), ("kmissing", C.c_uint64)]


class PairCount(C.Structure):
    _fields_ = [("v_read", C.c_uint32), ("v_asm", C.c_uint32), ("count", C.c_uint64)]


class HistSummary(C.Structure):
    _fields_ = [("kasm", C.c_uint64), ("kmissing", C.c_uint64), ("koverCpy", C.c_double),
                ("qv_missing", C.c_double), ("qv_star", C.c_double),
                ("n_undr", C.c_uint32), ("n_over", C.c_uint32)]


class StreamStats(C.Structure):
    _fields_ = [("db_kmers", C.c_uint64), ("kept_kmers", C.c_uint64), ("shard_kmers", C.c_uint64), ("table_bytes", C.c_uint64),
                ("seconds", C.c_double)]


class VcfInfo(C.Structure):
    _fields_ = [("headers", C.c_uint64), ("records", C.c_uint64), ("contigs", C.c_uint64), ("contig_ids", C.c_uint64),
                ("excluded", C.c_uint64), ("clusters", C.c_uint64)]


MFN_MODE = {"filter": 4, "polish": 5, "better": 6, "strict": 7, "loose": 8}

_P = C.c_void_p
_PP = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); mirrors include/merfin_b200.h one to one
SIGNATURES = {
    "mfn_strerror": (C.c_char_p, [C.c_int]),
    "mfn_last_error": (C.c_char_p, []),
    "mfn_version": (C.c_char_p, []),
    "mfn_ctx_create": (C.c_int, [_PP, C.c_int]),
    "mfn_ctx_destroy": (None, [_P]),
    "mfn_ctx_device": (C.c_int, [_P]),
    "mfn_ctx_set_stream": (C.c_int, [_P, _P]),
    "mfn_ctx_sync": (C.c_int, [_P]),
    "mfn_ctx_set_sweep": (C.c_int, [_P, C.c_int]),
    "mfn_host_alloc": (C.c_int, [_PP, C.c_size_t]),
    "mfn_host_free": (None, [_P]),
    "mfn_table_load": (C.c_int, [_P, C.c_char_p, C.c_uint32, C.c_uint32, _PP]),
    "mfn_table_from_kmers": (C.c_int, [_P, C.c_uint32, _P, _P, C.c_uint64, C.c_uint32, C.c_uint32, _PP]),
    "mfn_table_count_seq": (C.c_int, [_P, C.c_uint32, _P, _PP]),
    "mfn_table_synth_reads": (C.c_int, [_P, _P, _P, C.c_double, C.c_uint64, C.c_uint64, _PP]),
    "mfn_table_info": (C.c_int, [_P, C.POINTER(TableInfo)]),
    "mfn_table_free": (None, [_P]),
    "mfn_table_lookup": (C.c_int, [_P, _P, _P, C.c_uint64, _P]),
    "mfn_table_lookup_wide": (C.c_int, [_P, _P, _P, _P, C.c_uint64, _P]),
    "mfn_table_export": (C.c_int, [_P, _P, _P, _P, C.c_uint64, C.POINTER(C.c_uint64)]),
    "mfn_tables_fuse": (C.c_int, [_P, _P, _P]),
    "mfn_tables_unfuse": (C.c_int, [_P]),
    "mfn_joint_info": (C.c_int, [_P, C.POINTER(JointInfo)]),
    "mfn_tables_restrict": (C.c_int, [_P, _P, _P, _P, C.c_uint32, C.c_uint32]),
    "mfn_tables_unrestrict": (C.c_int, [_P]),
    "mfn_restricted_info": (C.c_int, [_P, C.POINTER(JointInfo)]),
    "mfn_tables_restrict_stream": (C.c_int, [_P, C.c_char_p, C.c_uint32, C.c_uint32, _P, _P, C.c_uint32, C.c_uint32, _PP, C.POINTER(StreamStats)]),
    "mfn_tables_restrict_synth": (C.c_int, [_P, _P, _P, C.c_uint32, C.c_uint32, C.c_double, C.c_uint64, C.c_uint64, C.c_uint64, _PP,
                                            C.POINTER(StreamStats)]),
    "mfn_joint_lookup": (C.c_int, [_P, _P, _P, C.c_uint64, _P, _P]),
    "mfn_meryl_write": (C.c_int, [C.c_char_p, C.c_uint32, _P, _P, C.c_uint64]),
    "mfn_meryl_info": (C.c_int, [C.c_char_p, C.POINTER(MerylInfo)]),
    "mfn_meryl_histogram": (C.c_int, [C.c_char_p, _P, _P, C.c_uint64, C.POINTER(C.c_uint64)]),
    "mfn_meryl_read": (C.c_int, [C.c_char_p, _P, _P, _P, C.c_uint64, C.POINTER(C.c_uint64)]),
    "mfn_seq_create": (C.c_int, [_P, _P, _P, C.c_uint32, _PP]),
    "mfn_seq_info": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]),
    "mfn_seq_free": (None, [_P]),
    "mfn_accum_create": (C.c_int, [_P, C.c_uint32, _PP]),
    "mfn_accum_reset": (C.c_int, [_P, _P]),
    "mfn_accum_free": (None, [_P]),
    "mfn_accum_device_buffer": (C.c_int, [_P, _PP, C.POINTER(C.c_uint64)]),
    "mfn_accum_fetch": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P, C.c_uint64, C.POINTER(C.c_uint64)]),
    "mfn_accum_read": (C.c_int, [_P, _P, _P]),
    "mfn_accum_sparse_fetch": (C.c_int, [_P, _P, _P, C.c_uint64, C.POINTER(C.c_uint64)]),
    "mfn_accum_decode_gathered": (C.c_int, [_P, C.c_uint32, _P, C.c_uint64, _P, _P, C.c_uint64, C.POINTER(C.c_uint64)]),
    "mfn_accum_add_from": (C.c_int, [_P, _P, _P]),
    "mfn_accum_layout": (C.c_int, [C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                   C.POINTER(C.c_uint64)]),
    "mfn_accum_decode": (C.c_int, [_P, C.c_uint32, _P, _P, C.c_uint64, C.POINTER(C.c_uint64)]),
    "mfn_shard_tiles": (C.c_int, [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64),
                                  C.POINTER(C.c_uint64)]),
    "mfn_hist_resident": (C.c_int, [_P, _P, _P, _P, C.POINTER(KModel), C.c_uint32, C.c_uint32, _P]),
    "mfn_hist_host": (C.c_int, [_P, _P, _P, _P, _P, C.c_uint32, C.POINTER(KModel), C.c_uint32, C.c_uint32, _P]),
    "mfn_hist": (C.c_int, [_P, _P, _P, _P, _P, C.c_uint32, C.POINTER(KModel), _P, _P, C.c_uint64,
                           C.POINTER(C.c_uint64)]),
    "mfn_dump_resident": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "mfn_dump": (C.c_int, [_P, _P, _P, _P, _P, C.c_uint32, _P, _P, _P]),
    "mfn_score_paths": (C.c_int, [_P, _P, _P, _P, _P, C.c_uint64, _P, _P, _P]),
    "mfn_vcf_load": (C.c_int, [C.c_char_p, _PP]),
    "mfn_vcf_cluster": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_int]),
    "mfn_vcf_info": (C.c_int, [_P, C.POINTER(VcfInfo)]),
    "mfn_vcf_header": (C.c_int, [_P, _P, C.c_uint64, C.POINTER(C.c_uint64)]),
    "mfn_vcf_clusters_of": (C.c_uint64, [_P, C.c_char_p]),
    "mfn_vcf_free": (None, [_P]),
    "mfn_variants_process": (C.c_int, [_P, _P, _P, _P, C.c_char_p, _P, C.c_uint64, C.c_int, C.c_uint32, C.POINTER(KModel),
                                       C.c_uint32, C.c_uint32, _PP, C.POINTER(C.c_uint64), _PP, C.POINTER(C.c_uint64),
                                       C.POINTER(C.c_uint64)]),
    "mfn_text_free": (None, [_P]),
    "mfn_completeness": (C.c_int, [_P, C.c_char_p, C.c_char_p, _P, C.c_uint64, C.POINTER(C.c_uint64)]),
    "mfn_completeness_resident": (C.c_int, [_P, _P, _P, _P]),
    "mfn_completeness_shard": (C.c_int, [_P, C.c_char_p, _P, C.c_uint32, C.c_uint32, _P]),
    "mfn_hist_report": (C.c_int, [_P, C.c_uint64, C.POINTER(KModel), C.c_uint32, _P, _P, C.c_uint32,
                                  C.POINTER(HistSummary)]),
    "mfn_hist_format": (C.c_int, [_P, C.c_uint32, _P, C.c_uint32, C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64)]),
    "mfn_getK": (None, [C.POINTER(KModel), C.c_uint32, C.c_uint32] + [C.POINTER(C.c_double)] * 3),
    "mfn_getKmetric": (C.c_double, [C.c_double, C.c_double]),
    "mfn_completeness_report": (C.c_int, [_P, C.c_uint64, C.POINTER(KModel), C.POINTER(C.c_double),
                                          C.POINTER(C.c_double)]),
    "mfn_ctx_launch_count": (C.c_uint64, [_P]),
    "mfn_ctx_h2d_bytes": (C.c_uint64, [_P]),
    "mfn_dump_write": (C.c_int, [C.c_int, C.c_char_p, C.c_uint64, _P, _P, _P, _P, C.c_int, _P]),
    "mfn_pack_host_range": (None, [_P, _P, C.c_uint32, C.c_uint64, C.c_uint64, _P, _P]),
    "mfn_pack_host_isa": (C.c_char_p, []),
    "mfn_ctx_time_passes": (C.c_int, [_P, C.c_int]),
    "mfn_ctx_pass_times": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint32)]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (built by `make -C merfin_b200/csrc` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} is missing: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()').  There is no CPU fallback.")
    lib = C.CDLL(SO_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != MFN_OK:
        lib = load()
        detail = lib.mfn_last_error().decode(errors="replace") or lib.mfn_strerror(status).decode()
        raise MfnError(status, detail)
