#!/bin/bash
# Round 2, GPU call 8: all GPU tests (CLI -devices, overflow recipe, new k_join), probe work items of 2048/4096 entries, other modes + k sweep
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_call8
mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" | tee -a $O/summary.txt
tail -15 $O/pytest_gpu.log | tee -a $O/summary.txt
show() { python - "$1" "$2" <<'PY' | tee -a gpurun_out/r2_call8/summary.txt
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d['roofline']
    print(sys.argv[2], round(d['value']/1e9,2),'Gb/s', round(d['ms_per_step'],2),'ms', {k:round(v,2) for k,v in (r['pass_ms'] or {}).items()}, 'e2e', round(d['e2e']['ms_per_step'],2), d['result']['kasm'], d['result']['kmissing'], d['result']['sweep_equals_direct'])
except Exception as e: print(sys.argv[2], 'ERR',e)
PY
}
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-direct-check > $O/bench_$name.json 2> $O/bench_$name.err; show $O/bench_$name.json $name; }
for v in p2048 p4096; do run $v MFN_LIB=merfin_b200/libmerfin_b200_$v.so; done
timeout 1200 python tools/bench_modes.py --gbp 1.0 --modes hist,dump,score --ks 15,21,31 > $O/modes_1gbp.jsonl 2> $O/modes_1gbp.err; echo "modes rc=$?" | tee -a $O/summary.txt
cut -c1-400 $O/modes_1gbp.jsonl | tee -a $O/summary.txt
timeout 900 python tools/bench_modes.py --gbp 0.3 --modes completeness --ks 21 > $O/modes_compl.jsonl 2> $O/modes_compl.err; echo "compl rc=$?" | tee -a $O/summary.txt
cut -c1-500 $O/modes_compl.jsonl | tee -a $O/summary.txt
tail -3 $O/modes_1gbp.err $O/modes_compl.err | tee -a $O/summary.txt
