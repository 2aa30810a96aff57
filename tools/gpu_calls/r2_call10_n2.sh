#!/bin/bash
# Round 2, GPU call 10 (2 GPUs): CLI -devices 0,1, bench.py at N = 2 (hist and polish) through torchrun
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_call10
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_cli.py -x -q -m gpu -k "devices" > $O/pytest_devices.log 2>&1; echo "pytest devices rc=$?" | tee -a $O/summary.txt
tail -5 $O/pytest_devices.log | tee -a $O/summary.txt
P=29517
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?" | tee -a $O/summary.txt
tail -1 $O/bench_n2.json | cut -c1-1500 | tee -a $O/summary.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus 2 --mode polish --steps 2 --warmup 1 > $O/bench_polish_n2.json 2> $O/bench_polish_n2.err; echo "polish n2 rc=$?" | tee -a $O/summary.txt
tail -1 $O/bench_polish_n2.json | cut -c1-1500 | tee -a $O/summary.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench n1 rc=$?" | tee -a $O/summary.txt
tail -1 $O/bench_n1.json | cut -c1-1500 | tee -a $O/summary.txt
tail -3 $O/*.err | cut -c1-300 >> $O/summary.txt
