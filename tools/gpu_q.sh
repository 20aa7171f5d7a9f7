#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -5
timeout 300 python -m pytest tests/test_gpu_fused.py -x -q 2>&1 | tail -3
for mode in "bwd 8" "0 8" "0 0" "bwd 0"; do set -- $mode
EXB_GEMM_CHAIN=$1 EXB_GEMM_MC=$2 timeout 400 python bench.py --steps 200 --warmup 20 > gpurun_out/r2_q_$1_$2.log 2>&1; echo "chain=$1 mc=$2 rc=$?"
grep '^{' gpurun_out/r2_q_$1_$2.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['final_loss'])"
done
EXB_GEMM_CHAIN=0 EXB_GEMM_MC=8 timeout 300 python tools/mp_timeline.py --steps 40 2>&1 | grep -E "^rank 0"
