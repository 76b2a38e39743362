#!/bin/bash
# first GPU call of round 2: parity with the new carry / full-fuzz tests, bench line with roofline_twitter, sanitizer
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > $O/gpu.txt 2>&1; nproc >> $O/gpu.txt; lscpu | grep -E 'Model name|Socket|NUMA' >> $O/gpu.txt
( time timeout 900 python -m pytest tests -m gpu -q --timeout 600 ) > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -5 $O/pytest_gpu.log
timeout 400 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; cut -c1-400 $O/bench_n1.json
timeout 300 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; cut -c1-300 $O/bench_reference.json
bash tools/sanitize.sh
