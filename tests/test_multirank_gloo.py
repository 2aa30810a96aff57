"""N>1 host-side path on CPU: two gloo ranks each accumulate their shard of the windows into a
flat buffer with the library's accumulator layout, sum-all-reduce it, and rank 0 decodes it with
mfn_accum_decode and replays K*.  (On GPUs the per-shard buffers come from mfn_hist_resident and
the all-reduce is NCCL; the layout, sharding and decode code is the same.)"""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, golden, out):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import oracle_lib as O
    from merfin_b200 import lib as L
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = L.load()
    d = os.path.join(golden, "g1_k21")
    recs = O.read_fasta(os.path.join(d, "asm.fa"))
    k = 21
    R, A = O.Lookup(O.DB(os.path.join(d, "reads.meryl"))), O.Lookup(O.DB(os.path.join(d, "asm.meryl")))
    nseq = len(recs)
    total = sum(len(b) for _, b in recs)
    hw, dr, da, nw = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint64()
    lib.mfn_accum_layout(nseq, C.byref(hw), C.byref(dr), C.byref(da), C.byref(nw))
    buf = np.zeros(nw.value, dtype=np.int64)
    gb, ge = C.c_uint64(), C.c_uint64()
    L.check(lib.mfn_shard_tiles(total, nseq, rank, world, C.byref(gb), C.byref(ge)))
    # this rank's windows: global coordinate = offset in the concatenation + record index
    g0 = 0
    for i, (_, b) in enumerate(recs):
        vr, va, ok = O.pairs(R, A, b)
        for p in range(len(ok)):
            g = g0 + p
            if not (gb.value <= g < ge.value) or not ok[p]:
                continue
            assert vr[p] < dr.value and va[p] < da.value
            buf[hw.value + int(va[p]) * dr.value + int(vr[p])] += 1
            buf[hw.value + dr.value * da.value + 2 * i] += 1
            buf[hw.value + dr.value * da.value + 2 * i + 1] += int(vr[p] == 0)
            buf[0] += 1
        g0 += len(b) + 1
    mine = int(buf[0])
    t = torch.from_numpy(buf)
    dist.all_reduce(t)                                  # the one collective of the path
    if rank == 0:
        from merfin_b200 import api
        per = np.zeros(nseq, dtype=api.COUNTS_DTYPE)
        pairs = np.zeros(1 << 16, dtype=api.PAIR_DTYPE)
        n = C.c_uint64()
        L.check(lib.mfn_accum_decode(buf.ctypes.data_as(C.c_void_p), nseq, per.ctypes.data_as(C.c_void_p),
                                     pairs.ctypes.data_as(C.c_void_p), len(pairs), C.byref(n)))
        res = api.HistResult([r[0] for r in recs], per, pairs[:n.value], k, api.KModel(8.0))
        ok = (res.file_text() == open(os.path.join(d, "ref.hist"), "rb").read()
              and res.summary_text() == open(os.path.join(d, "ref.hist.summary")).read()
              and res.per_seq_text() == open(os.path.join(d, "ref.hist.perseq")).read())
        open(out, "w").write(f"{int(ok)} {mine} {int(buf[0])}")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_allreduce_decode(golden_dir, tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, _free_port(), golden_dir, out), nprocs=2, join=True)
    ok, mine, total = open(out).read().split()
    assert ok == "1"
    assert 0 < int(mine) < int(total)              # both ranks contributed


def _worker_sparse(rank, world, port, out):
    """Pairs outside the dense tier (v_asm >= 64 or v_read >= 4096) are not in the flat buffer: each rank keeps
    its own list, the reduced header counts them over all ranks, and the decode wants exactly that many."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from merfin_b200 import api, lib as L
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = L.load()
    nseq = 3
    hw, dr, da, nw = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint64()
    lib.mfn_accum_layout(nseq, C.byref(hw), C.byref(dr), C.byref(da), C.byref(nw))
    buf = np.zeros(nw.value, dtype=np.int64)
    rng = np.random.default_rng(100 + rank)
    # dense pairs
    for _ in range(1000):
        vr, va = int(rng.integers(0, dr.value)), int(rng.integers(0, da.value))
        buf[hw.value + va * dr.value + vr] += 1
    # side-table pairs: some shared between the ranks (must be summed), some private
    mine = {(5000, 3): 10 + rank, (17, 64): 2, (4096 + rank, 100 + rank): 7}
    sparse = np.array([(vr, va, c) for (vr, va), c in sorted(mine.items())], dtype=api.PAIR_DTYPE)
    buf[1] = len(sparse)
    expect_dense = torch.from_numpy(buf.copy())
    t = torch.from_numpy(buf)
    dist.all_reduce(t)
    dist.all_reduce(expect_dense)
    gathered = [None] * world
    dist.all_gather_object(gathered, sparse)
    ok = True
    if rank == 0:
        ubuf = buf.view(np.uint64)
        per = np.zeros(nseq, dtype=api.COUNTS_DTYPE)
        pairs = np.zeros(1 << 16, dtype=api.PAIR_DTYPE)
        n = C.c_uint64()
        # the plain decode must refuse a buffer whose header counts side-table pairs ...
        st = lib.mfn_accum_decode(ubuf.ctypes.data_as(C.c_void_p), nseq, per.ctypes.data_as(C.c_void_p),
                                  pairs.ctypes.data_as(C.c_void_p), len(pairs), C.byref(n))
        ok = ok and st == L.MFN_ERR_INVALID
        # ... and so must a gather that misses a rank
        try:
            api.decode_gathered(ubuf, nseq, gathered[0])
            ok = False
        except L.MfnError as e:
            ok = ok and e.status == L.MFN_ERR_INVALID
        per, got = api.decode_gathered(ubuf, nseq, np.concatenate(gathered))
        want = {}
        for r in range(world):
            for (vr, va), c in {(5000, 3): 10 + r, (17, 64): 2, (4096 + r, 100 + r): 7}.items():
                want[(vr, va)] = want.get((vr, va), 0) + c
        d = expect_dense.numpy()
        for va in range(da.value):
            for vr in np.nonzero(d[hw.value + va * dr.value: hw.value + (va + 1) * dr.value])[0]:
                want[(int(vr), va)] = int(d[hw.value + va * dr.value + vr])
        have = {(int(p["v_read"]), int(p["v_asm"])): int(p["count"]) for p in got}
        ok = ok and have == want and len(got) == len(want)
        open(out, "w").write(str(int(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_side_table_pairs_are_gathered(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker_sparse, args=(2, _free_port(), out), nprocs=2, join=True)
    assert open(out).read() == "1"


def test_shard_ranges_partition_the_windows():
    from merfin_b200 import lib as L
    lib = L.load()
    for total, nseq, n in ((12000, 2, 2), (3_000_000_000, 24, 8), (5000, 3, 4), (10, 1, 3)):
        prev = 0
        for s in range(n):
            a, b = C.c_uint64(), C.c_uint64()
            L.check(lib.mfn_shard_tiles(total, nseq, s, n, C.byref(a), C.byref(b)))
            assert a.value == prev and b.value >= a.value
            prev = b.value
        assert prev == total + nseq
