#!/bin/bash
# One gpurun call (a call costs box time, so everything a change needs is done in one):
#   GPU parity suite + smoke(), the bench line and the CPU arm, the launch list of the bench batch, the other
#   BASELINE configs -- and the same suite / timings for every variant library found in build_variants/
#   (compile-time switches built with `nvcc -D...`, selected through SJ_B200_LIB), so an A/B decision and its
#   parity check come from the same box.
# Everything lands in gpurun_out/ (scratch; summaries are copied to profiles/ by hand).
#   usage: gpurun --timeout 600 -- 'bash tools/gpu_checks.sh [full]'      (full: + launch list of the configs)
set -u
MODE=${1:-default}
O=gpurun_out
V=$PWD/build_variants
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv,noheader > $O/gpu.txt 2>&1
nproc >> $O/gpu.txt

# essentials first: the box time left may be short
( time timeout 700 python -m pytest tests -m gpu -q --timeout 300 ) > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log
tail -4 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
tail -1 $O/smoke.log
timeout 400 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
echo "bench rc=$?"
cut -c1-300 $O/bench_n1.json
timeout 200 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
cut -c1-200 $O/bench_reference.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/launches_bench_ndjson128MiB.csv \
    python bench.py --steps 2 --warmup 1 --batch-mib 128 --inflight 1 --no-cpu > $O/bench_under_ncu.log 2>&1
timeout 200 python tools/config_bench.py 256 > $O/config_default.md 2>&1
tail -n +3 $O/config_default.md | cut -d'|' -f2,7,10,11

# variants: parity first, then their numbers
for v in $V/*.so; do
  [ -e "$v" ] || continue
  n=$(basename $v .so)
  ( SJ_B200_LIB=$v timeout 300 python -m pytest tests -m gpu -q --timeout 300 ) > $O/pytest_gpu_$n.log 2>&1
  echo "$n: $(tail -1 $O/pytest_gpu_$n.log)"
  SJ_B200_LIB=$v timeout 200 python bench.py --no-cpu > $O/bench_n1_$n.json 2> $O/bench_n1_$n.err
  cut -c1-200 $O/bench_n1_$n.json
  SJ_B200_LIB=$v timeout 200 python tools/config_bench.py 256 > $O/config_$n.md 2>&1
  tail -n +3 $O/config_$n.md | cut -d'|' -f2,7,10,11
done

if [ "$MODE" = full ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_configs_64MiB.csv \
      python tools/config_bench.py 64 twitterescaped,canada,twitter > $O/configs_under_ncu.log 2>&1
fi
ls -la $O | tail -24
