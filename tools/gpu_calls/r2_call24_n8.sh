#!/bin/bash
# Round 2, GPU call 24 (8 GPUs): bench.py --mode polish at N = 8 and N = 4
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_call24
mkdir -p $O
P=29717
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --mode polish --steps 2 --warmup 1 > $O/bench_polish_n8.json 2> $O/bench_polish_n8.err; echo "polish n8 rc=$?" | tee -a $O/summary.txt
tail -1 $O/bench_polish_n8.json | cut -c1-3000 | tee -a $O/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus 4 --mode polish --steps 2 --warmup 1 > $O/bench_polish_n4.json 2> $O/bench_polish_n4.err; echo "polish n4 rc=$?" | tee -a $O/summary.txt
tail -1 $O/bench_polish_n4.json | cut -c1-3000 | tee -a $O/summary.txt
tail -3 $O/*.err | cut -c1-300 >> $O/summary.txt
