#!/bin/bash
# Round 2, GPU call 25: scatter3 -hist parity at 110 Mbp; compute-sanitizer memcheck / racecheck over the new kernels
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_call25
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "scatter3_hist or dump_through or streamed" > $O/pytest_sel.log 2>&1; echo "pytest selected rc=$?" | tee -a $O/summary.txt
tail -6 $O/pytest_sel.log | cut -c1-300 | tee -a $O/summary.txt
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "scatter3_hist and True" > $O/memcheck_hist.log 2>&1; echo "memcheck hist rc=$?" | tee -a $O/summary.txt
grep -E "ERROR SUMMARY|Invalid|passed|failed" $O/memcheck_hist.log | tail -5 | tee -a $O/summary.txt
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "dump_through and True" > $O/memcheck_dump.log 2>&1; echo "memcheck dump rc=$?" | tee -a $O/summary.txt
grep -E "ERROR SUMMARY|Invalid|passed|failed" $O/memcheck_dump.log | tail -5 | tee -a $O/summary.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "streamed_read_database or streamed_synthetic" > $O/memcheck_stream.log 2>&1; echo "memcheck stream rc=$?" | tee -a $O/summary.txt
grep -E "ERROR SUMMARY|Invalid|passed|failed" $O/memcheck_stream.log | tail -5 | tee -a $O/summary.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "scatter3_hist and True" > $O/racecheck_hist.log 2>&1; echo "racecheck hist rc=$?" | tee -a $O/summary.txt
grep -E "RACECHECK SUMMARY|hazard|passed|failed" $O/racecheck_hist.log | tail -5 | tee -a $O/summary.txt
