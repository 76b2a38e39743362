#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -x ) > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
timeout 600 python bench.py --no-cpu --twitter-mib 0 --stream-gib 0 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -3 $O/bench_n1.err; cut -c1-130 $O/bench_n1.json
timeout 300 python tools/config_bench.py 256 > $O/config_stream.md 2>&1; cat $O/config_stream.md | cut -d'|' -f2,7,10,11
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/launches_bench_ndjson128MiB.csv \
    python bench.py --steps 2 --warmup 1 --batch-mib 128 --inflight 1 --no-cpu --twitter-mib 0 --stream-gib 0 > $O/bench_under_ncu.log 2>&1
python tools/summarize_launches.py $O/launches_bench_ndjson128MiB.csv 2>&1 | tail -14
timeout 600 ncu --set full --clock-control none --import-source on -k regex:s2s_emit -s 2 -c 1 -o $O/k2r -f \
   python bench.py --steps 1 --warmup 1 --batch-mib 128 --inflight 1 --no-cpu --twitter-mib 0 --stream-gib 0 > $O/ncu_k2r.log 2>&1
ls -la $O/*.ncu-rep
