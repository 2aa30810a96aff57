#!/bin/bash
# Round 2, GPU call 7: probe pass work-item size / L2 policy A/B at the new default geometry (64 MB slices); first run of bench.py --mode polish
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_call7
mkdir -p $O
show() { python - "$1" "$2" <<'PY' | tee -a gpurun_out/r2_call7/summary.txt
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d['roofline']
    print(sys.argv[2], round(d['value']/1e9,2),'Gb/s', round(d['ms_per_step'],2),'ms', {k:round(v,2) for k,v in (r['pass_ms'] or {}).items()}, 'e2e', round(d['e2e']['ms_per_step'],2), d['result']['kasm'], d['result']['kmissing'], d['result']['sweep_equals_direct'], (d['config'].get('restricted_table') or {}).get('lines'), (d['config'].get('restricted_table') or {}).get('spill_frac'))
except Exception as e: print(sys.argv[2], 'ERR',e)
PY
}
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-direct-check > $O/bench_$name.json 2> $O/bench_$name.err; show $O/bench_$name.json $name; }
run default A=1
for v in p512 p256 p1024el p512el; do run $v MFN_LIB=merfin_b200/libmerfin_b200_$v.so; done
run p256_32 MFN_LIB=merfin_b200/libmerfin_b200_p256.so MFN_SWEEP_SLICE_MB=32 MFN_BIN_STRIDE=16
run p512el_32 MFN_LIB=merfin_b200/libmerfin_b200_p512el.so MFN_SWEEP_SLICE_MB=32 MFN_BIN_STRIDE=16
timeout 1500 python bench.py --mode polish --steps 2 --warmup 1 > $O/bench_polish.json 2> $O/bench_polish.err; echo "polish rc=$?" | tee -a $O/summary.txt
tail -c 3000 $O/bench_polish.json | tee -a $O/summary.txt
tail -5 $O/bench_polish.err | tee -a $O/summary.txt
ls -la $O >> $O/summary.txt
