#!/bin/bash
# Round 2, GPU call 23: share of the first of two sub-sweeps in the host-buffer path; polish with the leaner replay
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_call23
mkdir -p $O
show() { python - "$1" "$2" <<'PY' | tee -a gpurun_out/r2_call23/summary.txt
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d['roofline']
    print(sys.argv[2], round(d['value']/1e9,2),'Gb/s', round(d['ms_per_step'],2),'ms', 'e2e', round(d['e2e']['ms_per_step'],2), round(d['e2e']['value']/1e9,1), d['result']['kasm'], d['result']['kmissing'], d['result']['resident_equals_e2e'])
except Exception as e: print(sys.argv[2], 'ERR',e)
PY
}
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-direct-check > $O/bench_$name.json 2> $O/bench_$name.err; show $O/bench_$name.json $name; }
run f50 MFN_SUBSWEEP_FIRST=0.5
run f60 MFN_SUBSWEEP_FIRST=0.6
run f67 MFN_SUBSWEEP_FIRST=0.67
run f75 MFN_SUBSWEEP_FIRST=0.75
run f50b MFN_SUBSWEEP_FIRST=0.5
timeout 1500 python bench.py --mode polish --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_polish.json 2> $O/bench_polish.err; echo "polish rc=$?" | tee -a $O/summary.txt
tail -1 $O/bench_polish.json | cut -c1-400 | tee -a $O/summary.txt
