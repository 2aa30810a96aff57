#!/bin/bash
# Round 2, GPU call 32: last full check of the final tree -- all GPU tests (incl. the k = 33..64 ones), smoke, the headline
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_call32
mkdir -p $O
timeout 460 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" | tee -a $O/summary.txt
tail -4 $O/pytest_gpu.log | cut -c1-300 | tee -a $O/summary.txt
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt; tail -1 $O/smoke.log | tee -a $O/summary.txt
timeout 200 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" | tee -a $O/summary.txt
tail -1 $O/bench_default.json | cut -c1-300 | tee -a $O/summary.txt
