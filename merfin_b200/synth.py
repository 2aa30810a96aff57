"""Seeded synthetic inputs: assemblies, reads, VCFs.

Small/medium sizes only (numpy on the host); the bench-scale tables are
synthesised on the GPU by the library itself (mfn_table_synth_reads).
"""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTacgtNn", b"TGCAtgcaNn"):
    _COMP[_a] = _b


def random_bases(n: int, rng: np.random.Generator) -> np.ndarray:
    """n uniform-random bases as an ASCII uint8 array."""
    return _ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]


def revcomp(b: np.ndarray) -> np.ndarray:
    return _COMP[b[::-1]]


def make_assembly(total: int, nrec: int, seed: int, n_run: int = 50, lower_run: int = 100,
                  dup_len: int = 0, names: str = "seq") -> list[tuple[str, np.ndarray]]:
    """Uniform-random assembly cut into nrec records.

    One N run and one lowercase run are planted in the last record; a segment
    of dup_len bases from record 0 is copied into the last record (so some
    k-mers have assembly multiplicity 2).
    """
    rng = np.random.default_rng(seed)
    bases = random_bases(total, rng)
    cuts = np.linspace(0, total, nrec + 1).astype(np.int64)
    recs = []
    for i in range(nrec):
        r = bases[cuts[i]:cuts[i + 1]].copy()
        recs.append((f"{names}{i + 1}", r))
    last = recs[-1][1]
    L = len(last)
    if dup_len and L > 4 * dup_len:
        last[L // 2: L // 2 + dup_len] = recs[0][1][10: 10 + dup_len]
    if n_run and L > 8 * n_run:
        last[L // 4: L // 4 + n_run] = ord("N")
    if lower_run and L > 8 * lower_run:
        s = slice(3 * L // 4, 3 * L // 4 + lower_run)
        last[s] = last[s] | 0x20
    return recs


def simulate_reads(recs, coverage: float, readlen: int, err: float, seed: int,
                   skip=None) -> list[np.ndarray]:
    """Reads sampled uniformly from the records (N -> random base), random strand,
    substitutions at rate err.  Regions in `skip` [(rec, bgn, end)] are never covered
    (their k-mers go missing from the read DB)."""
    rng = np.random.default_rng(seed)
    out = []
    for ri, (_, b) in enumerate(recs):
        src = b.copy()
        src &= 0xDF                                # upper-case
        nmask = ~np.isin(src, _ACGT)
        src[nmask] = random_bases(int(nmask.sum()), rng)
        L = len(src)
        if L < readlen:
            continue
        nreads = int(coverage * L / readlen)
        starts = rng.integers(0, L - readlen + 1, size=nreads)
        for s in starts:
            if skip and any(r == ri and s < e and s + readlen > a for r, a, e in skip):
                continue
            r = src[s:s + readlen].copy()
            ne = rng.binomial(readlen, err)
            if ne:
                pos = rng.integers(0, readlen, size=ne)
                r[pos] = _ACGT[(np.searchsorted(_ACGT, r[pos]) + rng.integers(1, 4, size=ne)) % 4]
            if rng.random() < 0.5:
                r = revcomp(r)
            out.append(r)
    return out


def write_fasta(path: str, recs, width: int = 0, prefix: str = "") -> None:
    with open(path, "wb") as f:
        for i, rec in enumerate(recs):
            if isinstance(rec, tuple):
                name, b = rec
            else:
                name, b = f"{prefix}{i}", rec
            f.write(b">" + name.encode() + b"\n")
            if width:
                for j in range(0, len(b), width):
                    f.write(b[j:j + width].tobytes() + b"\n")
            else:
                f.write(b.tobytes() + b"\n")


def make_vcf(recs, nvar: int, seed: int, gts=("1/1", "0/1", "1|0", "1/2"),
             near_frac: float = 0.2) -> str:
    """Sorted VCF (>=10 columns) of SNVs / small indels against recs."""
    rng = np.random.default_rng(seed)
    lines = ["##fileformat=VCFv4.2"]
    for name, b in recs:
        lines.append(f"##contig=<ID={name},length={len(b)}>")
    lines.append("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tSAMPLE")
    per = max(1, nvar // len(recs))
    for name, b in recs:
        L = len(b)
        pos = set()
        while len(pos) < per:
            p = int(rng.integers(50, L - 50))
            pos.add(p)
            if rng.random() < near_frac:
                pos.add(min(L - 50, p + int(rng.integers(1, 16))))
        last_end = -1
        for p in sorted(pos):
            if p <= last_end:
                continue
            kind = rng.random()
            ref_b = b[p:p + 5] & 0xDF
            if not np.isin(ref_b, _ACGT).all():
                continue
            gt = gts[int(rng.integers(0, len(gts)))]
            def other(c):
                return chr(_ACGT[(int(np.searchsorted(_ACGT, c)) + int(rng.integers(1, 4))) % 4])
            if kind < 0.8:
                ref = chr(ref_b[0]); alt = other(ref_b[0])
            elif kind < 0.9:
                ref = chr(ref_b[0]); alt = ref + random_bases(int(rng.integers(1, 5)), rng).tobytes().decode()
            else:
                n = int(rng.integers(1, 5)); ref = ref_b[:n + 1].tobytes().decode(); alt = ref[0]
            if gt == "1/2":
                alt2 = ref[0] + random_bases(int(rng.integers(1, 4)), rng).tobytes().decode()
                if alt2 == alt or alt2 == ref:
                    alt2 = alt + "A"
                alt = alt + "," + alt2
            last_end = p + len(ref) - 1
            lines.append(f"{name}\t{p + 1}\t.\t{ref}\t{alt}\t30\tPASS\t.\tGT\t{gt}")
    return "\n".join(lines) + "\n"


def plant_errors(truth_recs, n_per_rec: int, seed: int, false_per_rec: int = 0):
    """Assembly = truth with planted SNV / small indel errors; returns (asm_recs, vcf_text).

    The VCF proposes the corrections (REF = assembly allele, ALT = truth allele) with a mix of
    genotypes, some records paired with a close neighbour (clusters), a few `0/0` and `./.`
    records, and `false_per_rec` bogus SNVs per record that polishing should reject.
    """
    rng = np.random.default_rng(seed)
    asm_recs, lines = [], ["##fileformat=VCFv4.2"]
    body = []
    for name, t in truth_recs:
        t = t.copy()
        L = len(t)
        sites, p = [], 60
        while len(sites) < n_per_rec and p < L - 80:
            sites.append(p)
            if rng.random() < 0.25 and p + 20 < L - 80:      # a close neighbour -> same cluster
                sites.append(p + int(rng.integers(3, 13)))
            p = sites[-1] + int(rng.integers(45, max(46, 2 * L // max(1, n_per_rec))))
        out, edits, cur = [], [], 0                           # edits: (asm_pos, ref, alt)
        alen = 0
        for s in sites:
            if s <= cur or not np.isin(t[s:s + 8] & 0xDF, _ACGT).all():
                continue
            out.append(t[cur:s]); alen += s - cur
            kind = rng.random()
            tb = chr(t[s] & 0xDF)
            if kind < 0.6:                                     # SNV error in the assembly
                eb = chr(_ACGT[(int(np.searchsorted(_ACGT, t[s] & 0xDF)) + int(rng.integers(1, 4))) % 4])
                out.append(np.frombuffer(eb.encode(), dtype=np.uint8))
                edits.append((alen, eb, tb)); alen += 1; cur = s + 1
            elif kind < 0.8:                                   # assembly lost 1-3 bases after s
                n = int(rng.integers(1, 4))
                out.append(t[s:s + 1]); lost = (t[s + 1:s + 1 + n] & 0xDF).tobytes().decode()
                edits.append((alen, tb, tb + lost)); alen += 1; cur = s + 1 + n
            else:                                              # assembly gained 1-3 bases after s
                ins = random_bases(int(rng.integers(1, 4)), rng)
                out.append(t[s:s + 1]); out.append(ins)
                edits.append((alen, tb + ins.tobytes().decode(), tb)); alen += 1 + len(ins); cur = s + 1
        out.append(t[cur:])
        a = np.concatenate(out) if out else t
        asm_recs.append((name, a))
        lines.append(f"##contig=<ID={name},length={len(a)}>")
        recs = []
        for pos, ref, alt in edits:
            r = rng.random()
            gt = "1/1" if r < 0.55 else "0/1" if r < 0.7 else "1|0" if r < 0.8 else "1/2" if r < 0.9 else "0/0" if r < 0.95 else "./."
            alts = alt
            if gt == "1/2":
                alts = alt + "," + (alt + "A" if len(alt) < 4 else alt[:-1])
            recs.append((pos, f"{name}\t{pos + 1}\t.\t{ref}\t{alts}\t{int(rng.integers(10, 60))}\tPASS\t.\tGT:DP\t{gt}:20"))
        used = {pos for pos, _, _ in edits}
        for _ in range(false_per_rec):
            pos = int(rng.integers(30, len(a) - 30))
            if any(abs(pos - u) < 3 for u in used) or (a[pos] & 0xDF) not in _ACGT:
                continue
            used.add(pos)
            ref = chr(a[pos] & 0xDF)
            alt = chr(_ACGT[(int(np.searchsorted(_ACGT, a[pos] & 0xDF)) + int(rng.integers(1, 4))) % 4])
            recs.append((pos, f"{name}\t{pos + 1}\t.\t{ref}\t{alt}\t30\tPASS\t.\tGT\t1/1"))
        body += [line for _, line in sorted(recs)]
    lines.append("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tSAMPLE")
    return asm_recs, "\n".join(lines + body) + "\n"
