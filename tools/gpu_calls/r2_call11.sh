#!/bin/bash
# Round 2, GPU call 11: fixed/new GPU tests, CLI time-line at 100 Mbp, polish with pieces, ncu launch list + full capture at 3 Gbp
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_call11
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -x -q -m gpu -k "streamed or variant or devices or overflow" > $O/pytest_sel.log 2>&1; echo "pytest selected rc=$?" | tee -a $O/summary.txt
tail -8 $O/pytest_sel.log | cut -c1-300 | tee -a $O/summary.txt
timeout 900 python tools/bench_cli.py --mbp 100 --variants 200000 > $O/cli_100mbp.json 2> $O/cli_100mbp.err; echo "cli rc=$?" | tee -a $O/summary.txt
cat $O/cli_100mbp.json | tee -a $O/summary.txt
timeout 1500 python bench.py --mode polish --steps 2 --warmup 1 --no-cpu-baseline > $O/bench_polish.json 2> $O/bench_polish.err; echo "polish rc=$?" | tee -a $O/summary.txt
tail -1 $O/bench_polish.json | cut -c1-600 | tee -a $O/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --gbp 0.3 --steps 2 --warmup 1 --no-cpu-baseline --no-direct-check > $O/ncu_launch.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:k_sweep_ -s 9 -c 3 -o $O/prof_sweep_3gbp python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-direct-check > $O/ncu_full.log 2>&1
tail -3 $O/ncu_full.log | cut -c1-300 | tee -a $O/summary.txt
ls -la $O >> $O/summary.txt
