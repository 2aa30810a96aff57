#!/bin/bash
# Round 2, GPU call 9: all GPU tests (streamed read database, 100 Mbp CLI against the reference), -completeness time-line, read-database size sweep
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_call9
mkdir -p $O
MFN_TEST_OUT=$PWD/$O/cli_100mbp.json timeout 2400 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" | tee -a $O/summary.txt
tail -25 $O/pytest_gpu.log | cut -c1-300 | tee -a $O/summary.txt
cat $O/cli_100mbp.json | tee -a $O/summary.txt
MFN_LOAD_DEBUG=1 timeout 900 python tools/bench_modes.py --gbp 0.3 --modes completeness --ks 21 > $O/modes_compl.jsonl 2> $O/modes_compl.err; echo "compl rc=$?" | tee -a $O/summary.txt
cut -c1-500 $O/modes_compl.jsonl | tee -a $O/summary.txt
grep "meryl load" $O/modes_compl.err | tail -8 | tee -a $O/summary.txt
timeout 1500 python tools/bench_modes.py --gbp 1.0 --modes dbsweep > $O/modes_dbsweep.jsonl 2> $O/modes_dbsweep.err; echo "dbsweep rc=$?" | tee -a $O/summary.txt
cut -c1-900 $O/modes_dbsweep.jsonl | tee -a $O/summary.txt
tail -3 $O/modes_dbsweep.err | tee -a $O/summary.txt
