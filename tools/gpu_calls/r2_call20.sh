#!/bin/bash
# Round 2, GPU call 20: -dump through the range-partitioned scan, second try (queued slow path): parity + per-pass times
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_call20
mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "dump_through or dump or golden" > $O/pytest_sel.log 2>&1; echo "pytest selected rc=$?" | tee -a $O/summary.txt
tail -8 $O/pytest_sel.log | cut -c1-300 | tee -a $O/summary.txt
MFN_DUMP_DEBUG=1 timeout 900 python tools/bench_modes.py --gbp 1.0 --modes dump --ks 21 > $O/modes_dump_sweep.jsonl 2> $O/modes_dump_sweep.err
cut -c1-500 $O/modes_dump_sweep.jsonl | tee -a $O/summary.txt
grep "dump sweep" $O/modes_dump_sweep.err | tail -3 | tee -a $O/summary.txt
MFN_DUMP_DEBUG=1 timeout 900 python tools/bench_modes.py --gbp 3.0 --modes dump --ks 21 > $O/modes_dump_sweep3.jsonl 2> $O/modes_dump_sweep3.err
cut -c1-500 $O/modes_dump_sweep3.jsonl | tee -a $O/summary.txt
grep "dump sweep" $O/modes_dump_sweep3.err | tail -4 | tee -a $O/summary.txt
tail -3 $O/modes_dump_sweep3.err | tee -a $O/summary.txt
