#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -x ) > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -5 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python tools/fixture_bench.py 200 > $O/fixture_bench.md 2> $O/fixture_bench.err; cat $O/fixture_bench.md; tail -3 $O/fixture_bench.err
SJ_B200_STAGE2=legacy timeout 300 python tools/fixture_bench.py 100 > $O/fixture_bench_legacy.md 2>&1; cut -d'|' -f2,4,5,6 $O/fixture_bench_legacy.md
timeout 500 python bench.py --no-cpu --twitter-mib 0 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; cut -c1-130 $O/bench_n1.json; tail -3 $O/bench_n1.err
timeout 300 python tools/config_bench.py 256 > $O/config_stream.md 2>&1; cat $O/config_stream.md | cut -d'|' -f2,7,10,11
