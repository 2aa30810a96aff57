#!/bin/bash
# Round 2, GPU call 12: all GPU tests, 8192-window scatter3 with 512 bins of 32 MB, sub-sweep count of the host-buffer path, CLI time-line
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_call12
mkdir -p $O
timeout 2400 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" | tee -a $O/summary.txt
tail -8 $O/pytest_gpu.log | cut -c1-300 | tee -a $O/summary.txt
show() { python - "$1" "$2" <<'PY' | tee -a gpurun_out/r2_call12/summary.txt
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d['roofline']
    print(sys.argv[2], round(d['value']/1e9,2),'Gb/s', round(d['ms_per_step'],2),'ms', {k:round(v,2) for k,v in (r['pass_ms'] or {}).items()}, 'e2e', round(d['e2e']['ms_per_step'],2), round(d['e2e']['value']/1e9,1), d['result']['kasm'], d['result']['kmissing'], d['result']['sweep_equals_direct'], (d['config'].get('restricted_table') or {}).get('lines'))
except Exception as e: print(sys.argv[2], 'ERR',e)
PY
}
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-direct-check > $O/bench_$name.json 2> $O/bench_$name.err; show $O/bench_$name.json $name; }
run default A=1
run s32 MFN_SWEEP_SLICE_MB=32
run s32_st16 MFN_SWEEP_SLICE_MB=32 MFN_BIN_STRIDE=16
run s32_t2 MFN_SWEEP_SLICE_MB=32 MFN_S3_TILES=2 MFN_BIN_STRIDE=16
run t4_64 MFN_S3_TILES=4
run sub1 MFN_SUBSWEEPS=1
run sub3 MFN_SUBSWEEPS=3
run sub4 MFN_SUBSWEEPS=4
run s32_sub3 MFN_SWEEP_SLICE_MB=32 MFN_SUBSWEEPS=3
timeout 900 python tools/bench_cli.py --mbp 100 --variants 200000 > $O/cli_100mbp.json 2> $O/cli_100mbp.err; echo "cli rc=$?" | tee -a $O/summary.txt
cat $O/cli_100mbp.json | tee -a $O/summary.txt
