#!/bin/bash
# Round 2, GPU call 6: all GPU tests (variant ABI + CLI shards), spread reservation counters (MFN_BIN_STRIDE) A/B, geometry choices
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_call6
mkdir -p $O
timeout 1200 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" | tee -a $O/summary.txt
tail -5 $O/pytest_gpu.log | tee -a $O/summary.txt
show() { python - "$1" "$2" <<'PY' | tee -a gpurun_out/r2_call6/summary.txt
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d['roofline']
    print(sys.argv[2], round(d['value']/1e9,2),'Gb/s', round(d['ms_per_step'],2),'ms', {k:round(v,2) for k,v in (r['pass_ms'] or {}).items()}, 'e2e', round(d['e2e']['ms_per_step'],2), d['result']['kasm'], d['result']['kmissing'], d['result']['sweep_equals_direct'], (d['config'].get('restricted_table') or {}).get('lines'), (d['config'].get('restricted_table') or {}).get('spill_frac'))
except Exception as e: print(sys.argv[2], 'ERR',e)
PY
}
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-direct-check > $O/bench_$name.json 2> $O/bench_$name.err; show $O/bench_$name.json $name; }
run hi_s1 MFN_JOINT_LOAD_POW2=0.75 MFN_BIN_STRIDE=1
run hi_s4 MFN_JOINT_LOAD_POW2=0.75 MFN_BIN_STRIDE=4
run hi_s16 MFN_JOINT_LOAD_POW2=0.75 MFN_BIN_STRIDE=16
run lo_s1 MFN_BIN_STRIDE=1
run lo_s16 MFN_BIN_STRIDE=16
run lo64_s1 MFN_SWEEP_SLICE_MB=64 MFN_BIN_STRIDE=1
run lo64_s16 MFN_SWEEP_SLICE_MB=64 MFN_BIN_STRIDE=16
ls -la $O >> $O/summary.txt
