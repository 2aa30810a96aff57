#!/bin/bash
# First GPU call of the next round: everything that changed on the host after round 1's GPU budget ran
# out, measured in one go.  Run as
#   gpurun --timeout 1500 -- 'bash tools/r2_first_gpu_call.sh'
# Results land in gpurun_out/r2_first/.
cd "$(dirname "$0")/../.." || exit 1
O=gpurun_out/r2_first
mkdir -p $O
python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
MFN_TEST_EXTRA_K=1 python -m pytest tests/test_gpu_parity.py -q -k "test_sweep_equals_direct_scan" > $O/pytest_extra_k.log 2>&1; echo "extra k (32, 11) rc=$?" | tee -a $O/summary.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
# headline, with the ingest lanes' own report (new host packer: expect a higher host share than 0.69)
MFN_INGEST_DEBUG=1 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; tail -1 $O/bench.json | tee -a $O/summary.txt
grep "\[ingest\]" $O/bench.err | tail -3 | tee -a $O/summary.txt
# three sub-sweeps instead of two: a shorter exposed probe at the end of the e2e step?
MFN_SUBSWEEPS=3 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_sub3.json 2>/dev/null; tail -1 $O/bench_sub3.json | tee -a $O/summary.txt
# the CLI against the reference binary at BASELINE configs[1] (new -dump formatter / batching, FASTA loader, VCF reader)
python tools/bench_cli.py --mbp 100 > $O/cli_100mbp.json 2> $O/cli.err; tail -1 $O/cli_100mbp.json | tee -a $O/summary.txt
# the experimental scatter shape (16384-window chunks, one 1024-thread CTA per SM): correctness through the
# sweep tests, then the headline A/B (pass_ms in the JSON line: scatter / probe)
make -s -C merfin_b200/csrc exp -j8 > $O/make_exp.log 2>&1
MFN_LIB=merfin_b200/libmerfin_b200_exp.so python -m pytest tests/test_gpu_parity.py -q -k "sweep or midsize or host_packing" > $O/pytest_exp.log 2>&1; echo "exp lib sweep tests rc=$?" | tee -a $O/summary.txt
MFN_LIB=merfin_b200/libmerfin_b200_exp.so python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_exp.json 2>/dev/null; tail -1 $O/bench_exp.json | tee -a $O/summary.txt
