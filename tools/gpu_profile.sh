#!/bin/bash
# One gpurun call: `ncu --set full` captures (with source correlation) of the two stage-2 kernels that dominate
# the bench step, on the bench's own 128 MiB NDJSON batch.  Reports land in gpurun_out/ (scratch); read them here with
#   ncu -i gpurun_out/k2c.ncu-rep --page raw --csv      /      --page source --csv
#   usage: gpurun --timeout 600 -- 'bash tools/gpu_profile.sh'
set -u
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 2 --warmup 1 --batch-mib 128 --inflight 1 --no-cpu"
timeout 280 ncu --set full --import-source on --clock-control none -k regex:s2_emit_kernel -s 3 -c 1 -o $O/k2c -f $B > $O/ncu_k2c.log 2>&1
timeout 280 ncu --set full --import-source on --clock-control none -k regex:s2_classify_measure_kernel -s 3 -c 1 -o $O/k2a -f $B > $O/ncu_k2a.log 2>&1
ls -la $O/*.ncu-rep
