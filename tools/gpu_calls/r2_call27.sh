#!/bin/bash
# Round 2, GPU call 27: -completeness join over resident tables: parity on the goldens, rate at 3 Gbp (3.5 G read k-mers)
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_call27
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "completeness" > $O/pytest_sel.log 2>&1; echo "pytest selected rc=$?" | tee -a $O/summary.txt
tail -4 $O/pytest_sel.log | cut -c1-300 | tee -a $O/summary.txt
timeout 900 python tools/bench_modes.py --gbp 3.0 --modes join --ks 21 > $O/modes_join.jsonl 2> $O/modes_join.err; echo "join rc=$?" | tee -a $O/summary.txt
cut -c1-700 $O/modes_join.jsonl | tee -a $O/summary.txt
tail -3 $O/modes_join.err | tee -a $O/summary.txt
