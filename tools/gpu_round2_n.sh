#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:s2s_emit -s 3 -c 1 -o $O/k2r_twesc -f \
   python tools/config_bench.py 64 twitterescaped > $O/ncu_k2r_twesc.log 2>&1
tail -2 $O/ncu_k2r_twesc.log
ls -la $O/k2r_twesc.ncu-rep
