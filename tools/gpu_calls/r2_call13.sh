#!/bin/bash
# Round 2, GPU call 13: -dump through the range-partitioned scan (parity + rate), headline with the 32 MB default, shards of 4 / 8 emulated
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_call13
mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -x -q -m gpu -k "dump or compressed or golden" > $O/pytest_sel.log 2>&1; echo "pytest selected rc=$?" | tee -a $O/summary.txt
tail -8 $O/pytest_sel.log | cut -c1-300 | tee -a $O/summary.txt
MFN_DUMP_SWEEP=0 timeout 900 python tools/bench_modes.py --gbp 1.0 --modes dump --ks 21 > $O/modes_dump_direct.jsonl 2> $O/modes_dump_direct.err
cut -c1-500 $O/modes_dump_direct.jsonl | tee -a $O/summary.txt
timeout 900 python tools/bench_modes.py --gbp 1.0 --modes dump --ks 21 > $O/modes_dump_sweep.jsonl 2> $O/modes_dump_sweep.err
cut -c1-500 $O/modes_dump_sweep.jsonl | tee -a $O/summary.txt
tail -3 $O/modes_dump_sweep.err | tee -a $O/summary.txt
show() { python - "$1" "$2" <<'PY' | tee -a gpurun_out/r2_call13/summary.txt
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d['roofline']
    print(sys.argv[2], round(d['value']/1e9,2),'Gb/s', round(d['ms_per_step'],2),'ms', {k:round(v,2) for k,v in (r['pass_ms'] or {}).items()}, 'e2e', round(d['e2e']['ms_per_step'],2), round(d['e2e']['value']/1e9,1), d['result'], 'own', r.get('frac_of_own_traffic'), (d['config'].get('restricted_table') or {}).get('lines'), d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('parity_vs_gpu'))
except Exception as e: print(sys.argv[2], 'ERR',e)
PY
}
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; show $O/bench_default.json default
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --emulate-shards 8 > $O/bench_emu8.json 2> $O/bench_emu8.err; show $O/bench_emu8.json emu8
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --emulate-shards 4 > $O/bench_emu4.json 2> $O/bench_emu4.err; show $O/bench_emu4.json emu4
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; tail -1 $O/bench_reference.json | cut -c1-900 | tee -a $O/summary.txt
