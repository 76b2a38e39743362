#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -x ) > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -5 $O/pytest_gpu.log
timeout 600 python bench.py --no-cpu --twitter-mib 0 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -3 $O/bench_n1.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_n1.json') if l.startswith('{')][-1])
for k in ('value','ms_per_step','e2e','e2e_nocopy','stream','roofline_parse'): print(k, d.get(k))
PY
timeout 300 python tools/config_bench.py 256 twitter,twitterescaped,canada > $O/config_stream.md 2>&1; cat $O/config_stream.md | cut -d'|' -f2,7,10,11
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_configs_64MiB.csv \
      python tools/config_bench.py 64 canada,twitterescaped > $O/configs_under_ncu.log 2>&1
python tools/summarize_launches.py $O/launches_configs_64MiB.csv 2>&1 | tail -22
