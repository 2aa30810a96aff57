#!/bin/bash
# Round 2, GPU call 16: ncu of the -dump passes at 1 Gbp
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_call16
mkdir -p $O
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_sweep_gather|k_sweep_probe_dump|k_sweep_scatter3" -s 3 -c 3 -o $O/prof_dump python tools/bench_modes.py --gbp 1.0 --modes dump --ks 21 --steps 1 > $O/ncu.log 2>&1
tail -5 $O/ncu.log | cut -c1-300 | tee -a $O/summary.txt
