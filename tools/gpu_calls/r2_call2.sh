#!/bin/bash
# Round 2, GPU call 2: correctness of the rewritten scatter / restricted table, then the scatter CTA shapes A/B.
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_call2
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $O/pytest_parity.log 2>&1; echo "pytest parity rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_parity.log | tee -a $O/summary.txt
for v in "" t2x512x2 t1x256x4 t2x256x3 t4x1024x1; do
  lib=merfin_b200/libmerfin_b200${v:+_$v}.so
  MFN_LIB=$lib timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_${v:-default}.json 2> $O/bench_${v:-default}.err
  echo "== ${v:-default} rc=$?" | tee -a $O/summary.txt
  python - "$O/bench_${v:-default}.json" <<'PY' | tee -a $O/summary.txt
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d['value']/1e9,2),'Gb/s', round(d['ms_per_step'],2),'ms', d['roofline']['pass_ms'], 'e2e', round(d['e2e']['ms_per_step'],2), d['result'], d['config'].get('restricted_table'))
except Exception as e: print('ERR',e)
PY
done
# the full fused table (no restriction): bins of 32 MB with a 33-bit low part
MFN_BENCH_RESTRICT=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_norestrict.json 2> $O/bench_norestrict.err
python - <<'PY' | tee -a $O/summary.txt
import json
d=json.loads(open('gpurun_out/r2_call2/bench_norestrict.json').read().strip().splitlines()[-1])
print('norestrict', round(d['value']/1e9,2),'Gb/s', round(d['ms_per_step'],2),'ms', d['roofline']['pass_ms'], d['result'])
PY
# 16 MB slices with the restricted table
MFN_SWEEP_SLICE_MB=16 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_slice16.json 2> $O/bench_slice16.err
python - <<'PY' | tee -a $O/summary.txt
import json
d=json.loads(open('gpurun_out/r2_call2/bench_slice16.json').read().strip().splitlines()[-1])
print('slice16', round(d['value']/1e9,2),'Gb/s', round(d['ms_per_step'],2),'ms', d['roofline']['pass_ms'], d['result'])
PY
# phase breakdown of the scatter pass
MFN_LIB=merfin_b200/libmerfin_b200_prof.so timeout 600 python bench.py --gbp 1.0 --steps 1 --warmup 3 --no-cpu-baseline > $O/bench_prof.json 2> $O/bench_prof.err
grep "scatter phases" $O/bench_prof.err | tail -2 | tee -a $O/summary.txt
