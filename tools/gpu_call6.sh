#!/bin/bash
mkdir -p gpurun_out
for v in 0 1; do
EXB_SPARSE_V2=$v timeout 300 python tools/mp_timeline.py --steps 40 > gpurun_out/r2_timeline_v2is${v}_n1.log 2>&1
echo "v2=$v rc=$?"; grep -E "^rank 0|phases" gpurun_out/r2_timeline_v2is${v}_n1.log
done
