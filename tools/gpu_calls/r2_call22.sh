#!/bin/bash
# Round 2, GPU call 22: everything once more on the final code -- all GPU tests, smoke, headline + reference arm, ncu launch list and full capture, CLI
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_call22
mkdir -p $O
timeout 2400 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?" | tee -a $O/summary.txt
tail -6 $O/pytest_gpu.log | cut -c1-300 | tee -a $O/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt; tail -1 $O/smoke.log | tee -a $O/summary.txt
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?" | tee -a $O/summary.txt
tail -1 $O/bench_default.json | cut -c1-4000 | tee -a $O/summary.txt
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; tail -1 $O/bench_reference.json | cut -c1-600 | tee -a $O/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --gbp 0.3 --steps 2 --warmup 1 --no-cpu-baseline --no-direct-check > $O/ncu_launch.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_sweep_ -s 9 -c 3 -o $O/prof_sweep_3gbp python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-direct-check > $O/ncu_full.log 2>&1
tail -2 $O/ncu_full.log | cut -c1-200 | tee -a $O/summary.txt
timeout 900 python tools/bench_cli.py --mbp 100 --variants 200000 > $O/cli_100mbp.json 2> $O/cli_100mbp.err; echo "cli rc=$?" | tee -a $O/summary.txt
cat $O/cli_100mbp.json | tee -a $O/summary.txt
