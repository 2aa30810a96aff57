#!/usr/bin/env python
"""Adversarial differential fuzz of the variant front-end (merfin_b200/csrc/host/variants.cpp, driven by
build/variant_harness with the CPU oracle as the scorer) against the UNMODIFIED reference binary
(oracle/_ref/merfin, one thread).  CPU only; test infrastructure.

Each case: an assembly with planted errors + the VCF of their corrections (synth.plant_errors), plus
random extra records crowded near record starts / ends and near each other (overlapping REF spans,
duplicate positions, multi-allelic, lower-case REF, odd genotypes), sorted by position; all five
selectors with varying k / -comb / -nosplit.

    python tests/fuzz_variants.py FIRST_SEED LAST_SEED [workdir]

Round 1: seeds 1-89 x 5 modes, no difference (tests/test_variants_host.py runs one seed of it).
"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from merfin_b200 import synth
MERFIN=ROOT+"/oracle/_ref/merfin"; MERYL=ROOT+"/oracle/_ref/meryl"; H=ROOT+"/build/variant_harness"
def run(seed, k=21, L=4000, nrec=2, nvar=60, peak=8.0, flags=(), work="/tmp/mfn_vfuzz"):
    rng=np.random.default_rng(seed)
    d=os.path.join(work, "s%d"%seed); os.makedirs(d, exist_ok=True)
    truth=synth.make_assembly(L*nrec, nrec, seed, n_run=10, lower_run=20, dup_len=300)
    asm, pvcf = synth.plant_errors(truth, nvar, seed+5, false_per_rec=5)
    planted={}
    for l in pvcf.splitlines():
        if l and not l.startswith("#"):
            f=l.split("\t"); planted.setdefault(f[0],[]).append((int(f[1])-1,l))
    synth.write_fasta(d+"/asm.fa", asm, width=60)
    reads=synth.simulate_reads(truth, 10, 100, 0.01, seed+1)
    synth.write_fasta(d+"/reads.fa", reads, prefix="r")
    lines=["##fileformat=VCFv4.2"]
    for name,b in asm: lines.append("##contig=<ID=%s,length=%d>"%(name,len(b)))
    lines.append("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tsample")
    gts=["1/1","0/1","1|0","1/2","0/0","./.","1","0|1","2/1","1/1:5"]
    for name,b in asm:
        n=len(b); pos=set(); rec=list(planted.get(name,[]))
        for _ in range(nvar//2):
            r=rng.random()
            if r<0.1: p=int(rng.integers(0,k+2))
            elif r<0.2: p=int(rng.integers(max(0,n-k-3),n-1))
            elif r<0.5 and pos: p=min(n-2,max(0,int(list(pos)[int(rng.integers(0,len(pos)))])+int(rng.integers(-5,8))))
            else: p=int(rng.integers(0,n-1))
            pos.add(p)
        for p in sorted(pos):
            rl=int(rng.choice([1,1,1,2,3,6])); rl=min(rl,n-p)
            ref=bytes(b[p:p+rl]).decode()
            if rng.random()<0.8: ref=ref.upper()
            def alt():
                t=rng.random()
                if t<0.5: return "ACGT"[int(rng.integers(0,4))]+ref[1:] if rl>0 else "A"
                if t<0.75: return ref[:1]
                return ref+ "".join("ACGT"[int(x)] for x in rng.integers(0,4,int(rng.integers(1,5))))
            gt=gts[int(rng.integers(0,len(gts)))]
            alts=alt()
            if "2" in gt or rng.random()<0.1: alts+=","+alt()
            fmt="GT:DP" if ":" in gt else "GT"
            rec.append((p,"%s\t%d\t.\t%s\t%s\t30\tPASS\t.\t%s\t%s"%(name,p+1,ref,alts,fmt,gt)))
        rec.sort(key=lambda x:x[0])
        lines += [l for _,l in rec]
    open(d+"/calls.vcf","w").write("\n".join(lines)+"\n")
    for fa,db in (("reads.fa","reads.meryl"),("asm.fa","asm.meryl")):
        subprocess.run([MERYL,"threads=1","count","k=%d"%k,"memory=1",fa,"output",db],cwd=d,capture_output=True,check=True)
    res={}
    for mode in ("polish","filter","better","strict","loose"):
        p=subprocess.run([MERFIN,"-"+mode,"-sequence","asm.fa","-seqmers","asm.meryl","-readmers","reads.meryl","-vcf","calls.vcf","-peak",str(peak),"-threads","1","-memory","4","-output","ref."+mode]+list(flags),cwd=d,capture_output=True,text=True,timeout=300)
        if p.returncode!=0: res[mode]="ref rc=%d %s"%(p.returncode,p.stderr[-200:].replace("\n"," | ")); continue
        want=open(d+"/ref.%s.%s.vcf"%(mode,"polish" if mode=="polish" else "filter")).read()
        hf=[f.replace("-comb","--comb").replace("-nosplit","--nosplit") for f in flags]
        q=subprocess.run([H,d+"/asm.fa",d+"/reads.meryl",d+"/asm.meryl",d+"/calls.vcf",mode,str(peak)]+hf,capture_output=True,text=True,timeout=300)
        if q.returncode!=0: res[mode]="harness rc=%d %s"%(q.returncode,q.stderr[-200:]); continue
        if q.stdout!=want:
            open(d+"/got.%s.vcf"%mode,"w").write(q.stdout)
            res[mode]="DIFF (%d vs %d lines)"%(len(q.stdout.splitlines()),len(want.splitlines()))
        else: res[mode]="ok %d"%len(want.splitlines())
    return res
if __name__=="__main__":
    work = sys.argv[3] if len(sys.argv) > 3 else "/tmp/mfn_vfuzz"
    for s in range(int(sys.argv[1]), int(sys.argv[2])):
        k=[21,15,31,12,24,32][s%6]; fl=[(),("-comb","2"),("-nosplit",),("-comb","4","-nosplit"),("-comb","30")][s%5]
        r=run(s,k=k,flags=fl,nvar=[60,150,30][s%3],L=[4000,9000][s%2],work=work)
        print(s,k,fl,r,flush=True)
