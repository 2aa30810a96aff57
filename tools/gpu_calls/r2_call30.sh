#!/bin/bash
# Round 2, GPU call 30: the k = 33..64 path (tests/test_gpu_wide.py), -polish with the flat-array host side,
# the main older parity tests over the code the k > 32 dispatch touched, smoke
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_call30
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_wide.py -q -m gpu > $O/pytest_wide.log 2>&1; echo "pytest wide rc=$?" | tee -a $O/summary.txt
tail -25 $O/pytest_wide.log | cut -c1-400 | tee -a $O/summary.txt
timeout 420 python bench.py --mode polish --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_polish.json 2> $O/bench_polish.err; echo "polish rc=$?" | tee -a $O/summary.txt
tail -1 $O/bench_polish.json | cut -c1-600 | tee -a $O/summary.txt; tail -2 $O/bench_polish.err | tee -a $O/summary.txt
timeout 300 python -m pytest tests/test_gpu_cli.py tests/test_gpu_parity.py -q -m gpu -x \
  -k "cli_variant or variants_abi or cli_hist_dump or table_load_and_value or hist_matches or dump_matches or completeness_matches or score_paths or count_sequence or k_mismatch or completeness_over_resident" \
  > $O/pytest_subset.log 2>&1; echo "pytest subset rc=$?" | tee -a $O/summary.txt
tail -5 $O/pytest_subset.log | cut -c1-300 | tee -a $O/summary.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt; tail -1 $O/smoke.log | tee -a $O/summary.txt
