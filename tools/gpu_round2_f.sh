#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
( time timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -x ) > $O/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $O/pytest_gpu.log; tail -6 $O/pytest_gpu.log
timeout 500 python bench.py --no-cpu > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; cut -c1-130 $O/bench_n1.json; tail -3 $O/bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json'))
for k in ('value','ms_per_step','e2e','e2e_nocopy','roofline_parse','parse_count_where'): print(k, d.get(k))
print(d['config'].get('numa_node_bound'))
PY
