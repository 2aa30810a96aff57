#!/bin/bash
# Round 2, GPU call 21 (8 GPUs): bench.py at N = 8 and N = 4 through torchrun
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_call21
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus.txt 2>&1
P=29617
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --steps 5 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench n8 rc=$?" | tee -a $O/summary.txt
tail -1 $O/bench_n8.json | cut -c1-2500 | tee -a $O/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus 4 --steps 5 --warmup 3 > $O/bench_n4.json 2> $O/bench_n4.err; echo "bench n4 rc=$?" | tee -a $O/summary.txt
tail -1 $O/bench_n4.json | cut -c1-2500 | tee -a $O/summary.txt
tail -3 $O/*.err | cut -c1-300 >> $O/summary.txt
