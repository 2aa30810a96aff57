#!/bin/bash
# Round 2, GPU call 3 (re-entry): GPU tests + smoke + headline of HEAD, restricted vs full table, scatter CTA shapes, phase profile, ncu.
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_call3
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $O/gpu.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" | tee -a $O/summary.txt
tail -5 $O/pytest_gpu.log | tee -a $O/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
show() { python - "$1" "$2" <<'PY' | tee -a gpurun_out/r2_call3/summary.txt
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], round(d['value']/1e9,2),'Gb/s', round(d['ms_per_step'],2),'ms', d['roofline']['pass_ms'], 'kernel_ms', round(d['roofline']['kernel_ms'],2), 'e2e', round(d['e2e']['ms_per_step'],2), d['result'], d['config'].get('restricted_table'), d['config'].get('joint_table'), d['config']['tables'])
except Exception as e: print(sys.argv[2], 'ERR',e)
PY
}
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" | tee -a $O/summary.txt
show $O/bench_default.json default
MFN_BENCH_RESTRICT=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_norestrict.json 2> $O/bench_norestrict.err
show $O/bench_norestrict.json norestrict
for v in t1x256x4 t2x256x3 t4x1024x1; do
  MFN_LIB=merfin_b200/libmerfin_b200_$v.so timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_$v.json 2> $O/bench_$v.err
  show $O/bench_$v.json $v
done
MFN_SWEEP_SLICE_MB=16 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/bench_slice16.json 2> $O/bench_slice16.err
show $O/bench_slice16.json slice16
MFN_LIB=merfin_b200/libmerfin_b200_prof.so timeout 600 python bench.py --gbp 1.0 --steps 1 --warmup 3 --no-cpu-baseline > $O/bench_prof.json 2> $O/bench_prof.err
grep "scatter phases" $O/bench_prof.err | tail -2 | tee -a $O/summary.txt
# ncu: launch list, then full capture of the two sweep kernels
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --gbp 0.3 --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_sweep_ -s 6 -c 3 -o $O/prof_sweep python bench.py --gbp 1.0 --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_full.log 2>&1
ls -la $O | tee -a $O/summary.txt
