#!/bin/bash
# streaming stage 2 on the GPU for the first time: parity first, then numbers
set -u
O=gpurun_out
mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -x ) > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -15 $O/pytest_gpu.log
timeout 400 python bench.py --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; cut -c1-300 $O/bench_n1.json; tail -3 $O/bench_n1.err
SJ_B200_STAGE2=legacy timeout 400 python bench.py --no-cpu --twitter-mib 0 > $O/bench_n1_legacy.json 2> $O/bench_n1_legacy.err; cut -c1-200 $O/bench_n1_legacy.json
timeout 300 python tools/config_bench.py 256 > $O/config_stream.md 2>&1; cat $O/config_stream.md | cut -d'|' -f2,7,10,11
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/launches_bench_ndjson128MiB.csv \
    python bench.py --steps 2 --warmup 1 --batch-mib 128 --inflight 1 --no-cpu --twitter-mib 0 > $O/bench_under_ncu.log 2>&1
python tools/summarize_launches.py $O/launches_bench_ndjson128MiB.csv 2>&1 | tail -25
