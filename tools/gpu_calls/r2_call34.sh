#!/bin/bash
# Round 2, GPU call 34: CLI start-up overlap (FASTA read under the context creation, both databases loaded at once):
# every CLI test, then the 100 Mbp comparison with the reference (-hist, -polish, -filter)
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_call34
mkdir -p $O
timeout 100 python -m pytest tests/test_gpu_cli.py tests/test_gpu_wide.py -x -q -m gpu -k "cli and not 100mbp" > $O/pytest_cli.log 2>&1; echo "pytest cli rc=$?" | tee -a $O/summary.txt
tail -4 $O/pytest_cli.log | cut -c1-300 | tee -a $O/summary.txt
MFN_TIMELINE=1 timeout 110 python tools/bench_cli.py --no-dump > $O/cli.json 2> $O/cli.err; echo "cli rc=$?" | tee -a $O/summary.txt
tail -1 $O/cli.json | cut -c1-3000 | tee -a $O/summary.txt; tail -3 $O/cli.err | tee -a $O/summary.txt
