#!/bin/bash
mkdir -p gpurun_out
for mode in bwd 1 0; do
EXB_GEMM_CHAIN=$mode timeout 400 python bench.py --steps 200 --warmup 20 > gpurun_out/r2_q_$mode.log 2>&1; echo "mode=$mode rc=$?"
grep '^{' gpurun_out/r2_q_$mode.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['final_loss'])"
done
