#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
for v in 0 1; do
EXB_SPARSE_V2=$v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/mp_timeline.py --steps 40 > gpurun_out/r2_timeline_v2is${v}_n$N.log 2>&1
echo "v2=$v rc=$?"; grep -E "^rank 0|phases" gpurun_out/r2_timeline_v2is${v}_n$N.log
done
