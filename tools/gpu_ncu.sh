#!/bin/bash
# ncu --set full capture of one kernel of the fused step: tools/gpu_ncu.sh <kernel regex> <out name> [extra env]
K=$1; O=$2; shift; shift
mkdir -p gpurun_out
timeout 600 env "$@" ncu --set full --clock-control none --import-source on -k regex:$K -s 6 -c 2 -f -o gpurun_out/$O \
   python bench.py --steps 6 --warmup 3 --no-graph > gpurun_out/$O.log 2>&1
echo "ncu $K rc=$?"; tail -3 gpurun_out/$O.log | cut -c1-200
