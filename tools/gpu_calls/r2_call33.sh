#!/bin/bash
# Round 2, GPU call 33: the CLI at 100 Mbp against the reference binary with the final tree (-hist, -polish, -filter; no -dump)
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_call33
mkdir -p $O
MFN_TIMELINE=1 timeout 200 python tools/bench_cli.py --no-dump > $O/cli.json 2> $O/cli.err; echo "cli rc=$?" | tee -a $O/summary.txt
tail -1 $O/cli.json | cut -c1-3000 | tee -a $O/summary.txt; tail -3 $O/cli.err | tee -a $O/summary.txt
