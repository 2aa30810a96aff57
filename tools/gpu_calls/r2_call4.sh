#!/bin/bash
# Round 2, GPU call 4: scatter3 + in-line spill lookup of the probe pass: parity, A/B against the previous scatter, table load, shard emulation, ncu.
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_call4
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $O/pytest_parity.log 2>&1; echo "pytest parity rc=$?" | tee -a $O/summary.txt
tail -5 $O/pytest_parity.log | tee -a $O/summary.txt
show() { python - "$1" "$2" <<'PY' | tee -a gpurun_out/r2_call4/summary.txt
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d['roofline']
    print(sys.argv[2], round(d['value']/1e9,2),'Gb/s', round(d['ms_per_step'],2),'ms', {k:round(v,2) for k,v in (r['pass_ms'] or {}).items()}, 'kernel_ms', round(r['kernel_ms'],2), 'e2e', round(d['e2e']['ms_per_step'],2), d['result'], 'own', r.get('frac_of_own_traffic'), (d['config'].get('restricted_table') or {}).get('lines'), (d['config'].get('restricted_table') or {}).get('spill_frac'))
except Exception as e: print(sys.argv[2], 'ERR',e)
PY
}
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_$name.json 2> $O/bench_$name.err; show $O/bench_$name.json $name; }
run default A=1
run oldscatter MFN_SCATTER=2
run load60 MFN_JOINT_LOAD_POW2=0.6
run norestrict MFN_BENCH_RESTRICT=0
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --emulate-shards 8 > $O/bench_emu8.json 2> $O/bench_emu8.err; show $O/bench_emu8.json emu8
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --emulate-shards 2 > $O/bench_emu2.json 2> $O/bench_emu2.err; show $O/bench_emu2.json emu2
MFN_LIB=merfin_b200/libmerfin_b200_prof.so timeout 600 python bench.py --gbp 1.0 --steps 1 --warmup 3 --no-cpu-baseline --no-direct-check > $O/bench_prof.json 2> $O/bench_prof.err
grep "scatter phases" $O/bench_prof.err | tail -2 | tee -a $O/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep_ -s 6 -c 3 -o $O/prof_sweep python bench.py --gbp 1.0 --steps 2 --warmup 1 --no-cpu-baseline --no-direct-check > $O/ncu_full.log 2>&1
ls -la $O >> $O/summary.txt
