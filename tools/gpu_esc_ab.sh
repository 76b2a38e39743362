#!/bin/bash
# A/B of escape-path variants (build_variants/*.so) against the in-tree library: parity first, then the BASELINE inputs
set -u
O=gpurun_out
mkdir -p $O
V=${1:-escnew}
( SJ_B200_LIB=$PWD/build_variants/$V.so timeout 900 python -m pytest tests/test_gpu_stage2.py -m gpu -q -x --timeout 600 ) > $O/pytest_$V.log 2>&1
tail -3 $O/pytest_$V.log
for lib in "" $PWD/build_variants/escnew.so $PWD/build_variants/escnofast.so; do
  echo "== ${lib:-in-tree}"
  SJ_B200_LIB=$lib timeout 300 python tools/config_bench.py 256 twitterescaped,twitter,gsoc-2018,parking-citations 2>&1 | cut -d'|' -f2,10,11 | tail -4
done
SJ_B200_LIB=$PWD/build_variants/$V.so bash tools/gpu_profile.sh twitterescaped 64 > $O/profile_$V.log 2>&1
ls $O/*.ncu-rep
