#!/bin/bash
# One gpurun call: `ncu --set full` captures (with source correlation) of the two streaming stage-2 kernels on one of the
# BASELINE inputs (default twitterescaped, 64 MiB).  Reports land in gpurun_out/ (scratch); read them here with
#   python tools/ncu_source_lines.py gpurun_out/k2r_<input>.ncu-rep 60
#   ncu -i gpurun_out/k2r_<input>.ncu-rep --page raw --csv
#   usage: gpurun --timeout 600 -- 'bash tools/gpu_profile.sh [input] [MiB]'
set -u
O=gpurun_out
IN=${1:-twitterescaped}
MIB=${2:-64}
mkdir -p $O
B="python tools/config_bench.py $MIB $IN"
timeout 280 ncu --set full --import-source on --clock-control none -k regex:s2s_emit_kernel -s 3 -c 1 -o $O/k2r_$IN -f $B > $O/ncu_k2r_$IN.log 2>&1
timeout 280 ncu --set full --import-source on --clock-control none -k regex:s2s_count_kernel -s 3 -c 1 -o $O/k2p_$IN -f $B > $O/ncu_k2p_$IN.log 2>&1
ls -la $O/*.ncu-rep
