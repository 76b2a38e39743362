#!/bin/bash
set -u
O=gpurun_out
V=$PWD/build_variants
mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -x ) > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -4 $O/pytest_gpu.log
timeout 400 python bench.py --no-cpu --twitter-mib 0 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; cut -c1-130 $O/bench_n1.json; tail -3 $O/bench_n1.err
timeout 300 python tools/config_bench.py 256 > $O/config_stream.md 2>&1; cat $O/config_stream.md | cut -d'|' -f2,7,10,11
for v in $V/*.so; do
  n=$(basename $v .so)
  ( SJ_B200_LIB=$v timeout 300 python -m pytest tests/test_gpu_stage2.py -m gpu -q --timeout 300 -x ) > $O/pytest_gpu_$n.log 2>&1
  echo "$n: $(tail -1 $O/pytest_gpu_$n.log)"
  SJ_B200_LIB=$v timeout 200 python bench.py --no-cpu --twitter-mib 0 > $O/bench_n1_$n.json 2> $O/bench_n1_$n.err
  cut -c1-130 $O/bench_n1_$n.json
  SJ_B200_LIB=$v timeout 200 python tools/config_bench.py 256 twitter,twitterescaped,canada > $O/config_$n.md 2>&1
  tail -n +3 $O/config_$n.md | cut -d'|' -f2,7,10,11
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/launches_bench_ndjson128MiB.csv \
    python bench.py --steps 2 --warmup 1 --batch-mib 128 --inflight 1 --no-cpu --twitter-mib 0 > $O/bench_under_ncu.log 2>&1
python tools/summarize_launches.py $O/launches_bench_ndjson128MiB.csv 2>&1 | tail -14
