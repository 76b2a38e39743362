#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -x ) > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -5 $O/pytest_gpu.log
timeout 300 python tools/config_bench.py 256 > $O/config_stream.md 2>&1; cat $O/config_stream.md | cut -d'|' -f2,7,10,11
timeout 300 python tools/fixture_bench.py 100 2>&1 | cut -d'|' -f2,4,5,9 | tail -15
bash tools/sanitize.sh
