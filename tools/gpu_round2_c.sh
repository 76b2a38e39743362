#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:s2s_emit -s 2 -c 1 -o $O/k2r -f \
   python bench.py --steps 1 --warmup 1 --batch-mib 128 --inflight 1 --no-cpu --twitter-mib 0 > $O/ncu_k2r.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:s2s_count -s 2 -c 1 -o $O/k2p -f \
   python bench.py --steps 1 --warmup 1 --batch-mib 128 --inflight 1 --no-cpu --twitter-mib 0 > $O/ncu_k2p.log 2>&1
tail -3 $O/ncu_k2r.log
ls -la $O/*.ncu-rep
