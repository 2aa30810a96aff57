#!/bin/bash
# Round 2, GPU call 29: probe pass of the first sub-sweep on a second stream next to the second sub-sweep's scatters (host-buffer path)
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_call29
mkdir -p $O
show() { python - "$1" "$2" <<'PY' | tee -a gpurun_out/r2_call29/summary.txt
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], round(d['value']/1e9,2),'Gb/s', 'e2e', round(d['e2e']['ms_per_step'],2), round(d['e2e']['value']/1e9,1), d['result']['kasm'], d['result']['kmissing'], d['result']['resident_equals_e2e'])
except Exception as e: print(sys.argv[2], 'ERR',e)
PY
}
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-direct-check > $O/bench_$name.json 2> $O/bench_$name.err; show $O/bench_$name.json $name; }
run base A=1
run a1_f50 MFN_ASYNC_PROBE=1
run a1_f70 MFN_ASYNC_PROBE=1 MFN_SUBSWEEP_FIRST=0.7
run a2_f70 MFN_ASYNC_PROBE=2 MFN_SUBSWEEP_FIRST=0.7
run a3_f60 MFN_ASYNC_PROBE=3 MFN_SUBSWEEP_FIRST=0.6
