#!/bin/bash
# Round 2, GPU call 26 (2 GPUs): bench.py at N = 2 with the final code (32 MB slices), CLI -devices tests on two real GPUs
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_call26
mkdir -p $O
P=29817
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?" | tee -a $O/summary.txt
tail -1 $O/bench_n2.json | cut -c1-3000 | tee -a $O/summary.txt
timeout 900 python -m pytest tests/test_gpu_cli.py -x -q -m gpu -k "devices" > $O/pytest_devices.log 2>&1; echo "pytest devices rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_devices.log | tee -a $O/summary.txt
