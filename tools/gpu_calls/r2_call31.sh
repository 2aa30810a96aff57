#!/bin/bash
# Round 2, GPU call 31: mfn_dump / mfn_score_paths with their per-call sequence arrays from the stream-ordered pool
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_call31
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_cli.py tests/test_gpu_parity.py tests/test_gpu_wide.py -q -m gpu -x \
  -k "cli_variant or variants_abi or dump or score_paths" > $O/pytest_subset.log 2>&1; echo "pytest subset rc=$?" | tee -a $O/summary.txt
tail -5 $O/pytest_subset.log | cut -c1-300 | tee -a $O/summary.txt
timeout 420 python bench.py --mode polish --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_polish.json 2> $O/bench_polish.err; echo "polish rc=$?" | tee -a $O/summary.txt
tail -1 $O/bench_polish.json | cut -c1-400 | tee -a $O/summary.txt; tail -2 $O/bench_polish.err | tee -a $O/summary.txt
