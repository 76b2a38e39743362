#!/bin/bash
set -u
O=gpurun_out
V=$PWD/build_variants
mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -x ) > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -3 $O/pytest_gpu.log
timeout 300 python tools/config_bench.py 256 > $O/config_stream.md 2>&1; cat $O/config_stream.md | cut -d'|' -f2,7,10,11
for v in $V/*.so; do
  n=$(basename $v .so)
  ( SJ_B200_LIB=$v timeout 300 python -m pytest tests/test_gpu_stage2.py -m gpu -q --timeout 300 -x ) > $O/pytest_gpu_$n.log 2>&1
  echo "$n: $(tail -1 $O/pytest_gpu_$n.log)"
  SJ_B200_LIB=$v timeout 200 python tools/config_bench.py 256 > $O/config_$n.md 2>&1
  tail -n +3 $O/config_$n.md | cut -d'|' -f2,7,10,11
done
